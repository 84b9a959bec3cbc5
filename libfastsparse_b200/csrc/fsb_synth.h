// fsb_synth.h -- counter-based synthetic COO generator, identical on host and device.
//
// Entry j of the COO is a pure function of (seed, j) built from the splitmix64
// finaliser; integer-only index maths so both sides agree bit for bit.
//   rows: uniform over [0, nrow)
//   cols: dist 0 -> uniform over [0, ncol)
//         dist 1 -> power law with exponent ~1 ("octave-uniform": the bit length of
//                   the rank is uniform, the rank is uniform inside its octave, so
//                   P(rank = k) ~ 1 / (k log2 ncol)), then the rank is scattered by a
//                   fixed affine permutation of the column ids so hot columns are
//                   not neighbours.
//   vals: uniform in [0, 1) with 53 random mantissa bits.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FSB_HD __host__ __device__ __forceinline__
#else
#define FSB_HD static inline
#endif

FSB_HD uint64_t fsb_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// unbiased-enough map of a 64-bit hash to [0, n): high word of the 128-bit product,
// computed from 32-bit halves so host and device share one code path
FSB_HD uint32_t fsb_scale_to(uint64_t h, uint32_t n) {
  const uint64_t hi = h >> 32, lo = h & 0xffffffffull;
  return (uint32_t)((hi * n + ((lo * n) >> 32)) >> 32);
}

FSB_HD int fsb_synth_row(uint64_t seed, uint64_t j, int nrow) {
  return (int)fsb_scale_to(fsb_mix64(seed ^ (3 * j)), (uint32_t)nrow);
}

FSB_HD int fsb_synth_col(uint64_t seed, uint64_t j, int ncol, int dist) {
  const uint64_t h = fsb_mix64(seed ^ (3 * j + 1));
  if (dist == 0) return (int)fsb_scale_to(h, (uint32_t)ncol);
  int nbits = 0;                                   // octaves covering [1, ncol]
  while ((1ll << nbits) <= (long long)ncol) ++nbits;
  const uint32_t oct = fsb_scale_to(h, (uint32_t)nbits);
  const uint64_t h2 = fsb_mix64(h);
  uint64_t rank = (1ull << oct) + fsb_scale_to(h2, (uint32_t)(1u << oct));   // in [2^oct, 2^(oct+1))
  rank = (rank - 1) % (uint64_t)ncol;              // 0-based, folded into range
  // affine permutation of the column ids: multiplier odd and not a multiple of 5;
  // made coprime with ncol by stepping past common factors
  uint64_t a = 2654435761ull % (uint64_t)ncol;
  if (a == 0) a = 1;
  for (;;) {
    uint64_t x = a, y = (uint64_t)ncol;
    while (y) { uint64_t t = x % y; x = y; y = t; }
    if (x == 1) break;
    ++a;
  }
  return (int)((rank * a + 12345ull) % (uint64_t)ncol);
}

FSB_HD double fsb_synth_val(uint64_t seed, uint64_t j) {
  return (double)(fsb_mix64(seed ^ (3 * j + 2)) >> 11) * (1.0 / 9007199254740992.0);
}

// standard normal N(0,1), element i a pure function of (seed, i): Box-Muller on two 53-bit
// uniforms.  (The reference's randn, bench_a_mul_b.c:43-60, is the polar method on erand48 --
// a sequential stream; a counter-based one lets every GPU thread, and the host-side check,
// regenerate any element.)  Host and device agree to the last ulps of log / cos.
#if defined(__CUDACC__) || defined(FSB_SYNTH_WITH_MATH)
#include <math.h>
FSB_HD double fsb_synth_normal(uint64_t seed, uint64_t i) {
  const double u1 = (double)((fsb_mix64(seed ^ (2 * i)) >> 11) + 1) * (1.0 / 9007199254740992.0);       // (0, 1]
  const double u2 = (double)(fsb_mix64(seed ^ (2 * i + 1)) >> 11) * (1.0 / 9007199254740992.0);         // [0, 1)
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
}
#endif

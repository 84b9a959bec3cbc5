// kernels_csr_stream.cu -- CSR SpMV / narrow SpMM (R = 1, 2, 4), merge-path "stream" kernel.
//
// Same products as bcsr_A_mul_B / _B2 / _B4 (csr.h:149-202) and csr_A_mul_B (csr.h:425-438),
// and -- through the cached transpose -- At_mul_B / sdm_At_mul_B (sparse.h:68-75,
// dsparse.h:54-62).  With one or a few right-hand sides the product is bound by STREAMING the
// matrix (4 or 12 bytes per entry); the dense operand is small and L2-resident.  Row-wise
// kernels cannot stream well: short rows leave lanes idle, long rows (10^5 entries in the
// transpose of a power-law matrix) serialise a CTA.  Here the work is split by ENTRIES, not
// rows (merge-path, Merrill & Garland): CTA b takes items [bT, (b+1)T) of the merged list
// (row ends, entries); its two end points are found by binary search in row_ptr.
//   1. the CTA's run of cols/vals comes into shared memory -- R = 1: by two TMA bulk copies (csr_stream_tma_kernel,
//      the default), otherwise by coalesced per-thread loads -- all 256 threads gather x[col] and leave the products
//      x[col]*val in shared memory: gather parallelism is independent of rows;
//   2. one thread per row that ENDS in the tile sums its segment in stored order and stores
//      Y[row];
//   3. the piece of the row that continues into the next tile goes to a carry slot; a tiny
//      fix-up kernel adds the carries in tile order.  Deterministic, no atomics, perfectly
//      balanced for any row-length distribution, empty rows included.
#include <stdint.h>

#include <algorithm>
#include <type_traits>

#include "fsb_device.cuh"
#include "fsb_internal.h"

using namespace fsbdev;

namespace {

constexpr int kThreads = 256;
// merged items (row ends + entries) per CTA: 16 KB of products in shared memory for every width
#ifndef FSB_STREAM_TILE
#define FSB_STREAM_TILE 2048   // 1024 measured 10 % slower at C3 (1.21 vs 1.10 ms); 4096 exceeds the 48 KB of static shared memory
#endif
template <int RT> struct Tile { static constexpr int n = FSB_STREAM_TILE / RT; };
// lanes per row in the shared-memory row reduction when rows are short (<= 24 entries on average): C3 double SpMV
// 0.966 ms with 1 lane, 0.949 with 2, 1.000 with 4 (profiles/r2g_c3_{base,sgl2,sgl4}.jsonl)
#ifndef FSB_STREAM_SHORT_GL
#define FSB_STREAM_SHORT_GL 2
#endif

// merge-path split: first i such that row_end[i] > d - i - 1, i.e. rows [0,i) are complete
// once d items of the merged (row ends, entries) sequence are consumed
__device__ __forceinline__ int merge_search(const int* __restrict__ row_ptr, int nrow, long long nnz, long long d) {
  long long lo = d > nnz ? d - nnz : 0;
  long long hi = d < nrow ? d : nrow;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)__ldg(row_ptr + mid + 1) <= d - mid - 1) lo = mid + 1; else hi = mid;
  }
  return (int)lo;
}

// the tile boundaries depend only on the matrix: computed once per (matrix, tile size) and cached
__global__ void merge_splits_kernel(int nrow, long long nnz, const int* __restrict__ row_ptr, int tile, int ntiles, int* __restrict__ split) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > ntiles) return;
  const long long total = (long long)nrow + nnz;
  const long long d = min((long long)b * tile, total);
  split[b] = merge_search(row_ptr, nrow, nnz, d);
}

// Phase 2 and 3 of a tile: rows ending in the tile are summed from the products in shared memory (sp[t * RT + q] for
// entry t of the tile), the piece of the row that continues into the next tile becomes the tile's carry.
template <int RT>
__device__ __forceinline__ void reduce_tile(const double* __restrict__ s_p, const int* __restrict__ s_end, long long j0, int i0, int i1,
                                            int ndone, int nn, int nrow, double* __restrict__ Y, int* __restrict__ carry_row,
                                            double* __restrict__ carry_val) {
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  // rows ending in this tile: GL lanes per row chosen from the tile's mean row length
  // (1 lane for short rows -- a shuffle tree per 20-entry row costs more than the sum itself --
  // 8 lanes up to ~128 entries, a full warp beyond); lanes stride over the segment and are
  // combined by a fixed shuffle tree, so the result does not depend on scheduling
  auto reduce_rows = [&](auto gl_tag) {
    constexpr int GL = decltype(gl_tag)::value;
    const int grp = tid / GL, gl = tid % GL;
    for (int k = grp; k < ndone; k += kThreads / GL) {
      const int s = (k == 0) ? 0 : (int)(s_end[k - 1] - j0);
      const int e = (int)(s_end[k] - j0);
      double acc[RT];
#pragma unroll
      for (int q = 0; q < RT; ++q) acc[q] = 0.0;
      for (int t = s + gl; t < e; t += GL)
#pragma unroll
        for (int q = 0; q < RT; ++q) acc[q] += s_p[t * RT + q];
      if (GL > 1) {
        const unsigned mask = GL == 32 ? 0xffffffffu : (((1u << GL) - 1u) << (lane - gl));
#pragma unroll
        for (int off = GL / 2; off > 0; off >>= 1)
#pragma unroll
          for (int q = 0; q < RT; ++q) acc[q] += __shfl_xor_sync(mask, acc[q], off, GL);
      }
      if (gl == 0) {
#pragma unroll
        for (int q = 0; q < RT; ++q) Y[(long long)(i0 + k) * RT + q] = acc[q];
      }
    }
  };
  if (nn <= 24 * ndone) reduce_rows(std::integral_constant<int, FSB_STREAM_SHORT_GL>{});
  else if (nn <= 128 * ndone) reduce_rows(std::integral_constant<int, 8>{});
  else reduce_rows(std::integral_constant<int, 32>{});
  // the row that continues past this tile: its piece here becomes a carry
  if (tid < 32) {
    const int s = (ndone == 0) ? 0 : (int)(s_end[ndone - 1] - j0);
    double acc[RT];
#pragma unroll
    for (int q = 0; q < RT; ++q) acc[q] = 0.0;
    const bool has = (i1 < nrow) && (s < nn);
    if (has) {
      for (int t = s + tid; t < nn; t += 32)
#pragma unroll
        for (int q = 0; q < RT; ++q) acc[q] += s_p[t * RT + q];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int q = 0; q < RT; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], off);
    if (tid == 0) {
      carry_row[blockIdx.x] = has ? i1 : -1;
#pragma unroll
      for (int q = 0; q < RT; ++q) carry_val[(long long)blockIdx.x * RT + q] = acc[q];
    }
  }
}

// MINB = minimum resident CTAs per SM promised to ptxas.  Without one it budgets 32 registers and serialises the PER
// independent loads.  Two builds: 6 CTAs/SM (42 registers) is 11 % faster when the dense operand is small and mostly
// hits in cache (C3 double SpMV, x = 8 MB: 0.97 vs 1.08 ms), 4 CTAs/SM (56 registers, all loads of a thread in flight)
// is 13 % faster when it is large (the transpose, x = 80 MB: 1.31 vs 1.48 ms).  The launcher picks by operand size.
// POL (knob "stream_policy", off by default): L2 cache-policy words on the loads -- the matrix stream (cols / vals,
// touched once) evict_first, the gathered dense operand evict_last, like the staged kernel (fsb_device.cuh).
template <int RT, bool VALS, int MINB, int POL>
__global__ void __launch_bounds__(kThreads, MINB)
csr_stream_kernel(int nrow, long long nnz, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                  const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                  const int* __restrict__ split, int* __restrict__ carry_row, double* __restrict__ carry_val) {
  constexpr int kTile = Tile<RT>::n;
  __shared__ int s_end[kTile + 1];
  __shared__ double s_p[kTile * RT];
  const int tid = threadIdx.x;
  const long long total = (long long)nrow + nnz;
  const long long d0 = (long long)blockIdx.x * kTile;
  const long long d1 = min(d0 + kTile, total);
  const int i0 = __ldg(split + blockIdx.x), i1 = __ldg(split + blockIdx.x + 1);
  const long long j0 = d0 - i0, j1 = d1 - i1;
  const int ndone = i1 - i0;            // rows whose end falls in this tile
  const int nn = (int)(j1 - j0);        // entries in this tile
  const int nend = ndone + (i1 < nrow ? 1 : 0);
  for (int k = tid; k < nend; k += kThreads) s_end[k] = __ldg(row_ptr + i0 + 1 + k);
  {
    // all of this thread's entries are loaded before any gather is issued, and all gathers
    // before any product is stored: PER independent loads in flight per thread in each phase
    constexpr int PER = kTile / kThreads;
    int c[PER];
    double v[PER];
    double xv[PER][RT];
    unsigned long long pol_stream = 0, pol_keep = 0;
    if (POL == 1) { pol_stream = make_l2_policy(2); pol_keep = make_l2_policy(1); }
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int t = tid + p * kThreads;
      c[p] = 0;
      v[p] = 1.0;
      if (t < nn) {
        c[p] = POL == 1 ? ld_stream_s32_pol(cols + j0 + t, pol_stream) : ld_stream_s32(cols + j0 + t);
        if (VALS) v[p] = POL == 1 ? ld_stream_f64_pol(vals + j0 + t, pol_stream) : ld_stream_f64(vals + j0 + t);
      }
    }
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int t = tid + p * kThreads;
#pragma unroll
      for (int k = 0; k < RT; ++k) {
        xv[p][k] = 0.0;
        if (t < nn) {
          if (POL == 1) XLoad<1>::ldp(&xv[p][k], X + (long long)c[p] * RT + k, pol_keep);
          else xv[p][k] = __ldg(X + (long long)c[p] * RT + k);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int t = tid + p * kThreads;
      if (t < nn) {
#pragma unroll
        for (int k = 0; k < RT; ++k) s_p[t * RT + k] = VALS ? xv[p][k] * v[p] : xv[p][k];
      }
    }
  }
  __syncthreads();
  reduce_tile<RT>(s_p, s_end, j0, i0, i1, ndone, nn, nrow, Y, carry_row, carry_val);
}

// ---- R = 1, TMA-fed form (default when the arrays sit on 16-byte boundaries).  The tile's run of column indices and
// values comes into shared memory by two bulk copies (cp.async.bulk, SASS UBLKCP, completion on one mbarrier, L2
// evict_first) instead of per-thread LDGs: the matrix stream no longer passes through L1TEX -- whose tag stage the
// 8-byte gathers saturate (one wavefront per gather, profiles/r2_gather_ceiling.md) -- nor through registers (40 at six
// CTAs per SM with all eight gathers of a thread in flight).  A run starts at an arbitrary entry: each copy starts at the
// enclosing 16-byte boundary and the entries are read at a shift (0..3 indices, 0..1 values).  The row ends share the
// index buffer (a tile holds kTile merged items: nn entries + ndone row ends), the products overwrite the values in
// place, so the footprint stays at 24.7 KB per CTA.
template <bool VALS, int MINB, bool TEX>
__global__ void __launch_bounds__(kThreads, MINB)
csr_stream_tma_kernel(int nrow, long long nnz, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                      const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                      const int* __restrict__ split, int* __restrict__ carry_row, double* __restrict__ carry_val,
                      cudaTextureObject_t xtex, int xtex_off) {
  constexpr int kTile = Tile<1>::n;
  __shared__ __align__(16) int s_ci[kTile + 16];      // [index run from the 16-byte boundary | row ends]
  __shared__ __align__(16) double s_pv[kTile + 4];    // values in, products out (same slots)
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x;
  const long long total = (long long)nrow + nnz;
  const long long d0 = (long long)blockIdx.x * kTile;
  const long long d1 = min(d0 + kTile, total);
  const int i0 = __ldg(split + blockIdx.x), i1 = __ldg(split + blockIdx.x + 1);
  const long long j0 = d0 - i0, j1 = d1 - i1;
  const int ndone = i1 - i0;
  const int nn = (int)(j1 - j0);
  const int nend = ndone + (i1 < nrow ? 1 : 0);
  const int shc = (int)(j0 & 3), shv = (int)(j0 & 1);
  const int ncw = (shc + nn + 3) & ~3;                // index slots the bulk copy fills
  int* s_end = s_ci + ncw;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    if (nn > 0) {
      const unsigned long long spol = make_l2_policy(2);
      const unsigned bc = (unsigned)ncw * 4u;
      const unsigned bv = VALS ? (unsigned)((shv + nn + 1) & ~1) * 8u : 0u;
      mbar_expect_tx(&s_bar, bc + bv);
      tma_load_1d(s_ci, cols + (j0 - shc), bc, &s_bar, spol);
      if (VALS) tma_load_1d(s_pv, vals + (j0 - shv), bv, &s_bar, spol);
    }
  }
  for (int k = tid; k < nend; k += kThreads) s_end[k] = __ldg(row_ptr + i0 + 1 + k);
  __syncthreads();                                    // barrier initialised, row ends in place
  if (nn > 0) mbar_wait(&s_bar, 0);
  {
    constexpr int PER = kTile / kThreads;
    double xv[PER];
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int t = tid + p * kThreads;
      xv[p] = 0.0;
      if (t < nn) {
        if (TEX) {   // the gather through the texture pipe: its data stage is not the LSU's, which the shared-memory traffic shares
          const int2 w = tex1Dfetch<int2>(xtex, s_ci[shc + t] + xtex_off);
          xv[p] = __hiloint2double(w.y, w.x);
        } else {
          xv[p] = __ldg(X + s_ci[shc + t]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int t = tid + p * kThreads;
      if (t < nn) s_pv[shv + t] = VALS ? xv[p] * s_pv[shv + t] : xv[p];
    }
  }
  __syncthreads();
  reduce_tile<1>(s_pv + shv, s_end, j0, i0, i1, ndone, nn, nrow, Y, carry_row, carry_val);
}

// add the carried pieces to their rows, run by run, in tile order (deterministic)
template <int RT>
__global__ void csr_stream_fixup_kernel(int ntiles, const int* __restrict__ carry_row, const double* __restrict__ carry_val,
                                        double* __restrict__ Y) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ntiles) return;
  const int row = carry_row[b];
  if (row < 0 || (b > 0 && carry_row[b - 1] == row)) return;
  double acc[RT];
#pragma unroll
  for (int q = 0; q < RT; ++q) acc[q] = 0.0;
  for (int t = b; t < ntiles && carry_row[t] == row; ++t)
#pragma unroll
    for (int q = 0; q < RT; ++q) acc[q] += carry_val[(long long)t * RT + q];
#pragma unroll
  for (int q = 0; q < RT; ++q) Y[(long long)row * RT + q] += acc[q];
}

template <int RT, bool VALS>
int launch(fsb_matrix* A, double* dY, const double* dX, cudaStream_t st) {
  constexpr int kTile = Tile<RT>::n;
  const long long total = (long long)A->nrow + A->nnz;
  const int ntiles = (int)((total + kTile - 1) / kTile);
  if (ntiles == 0) return FSB_OK;
  double* scratch = nullptr;
  FSB_TRY(fsb_matrix_carry(A, (size_t)ntiles * (RT * sizeof(double) + sizeof(int)) + 16, &scratch));
  double* carry_val = scratch;
  int* carry_row = reinterpret_cast<int*>(scratch + (size_t)ntiles * RT);
  if (A->split_tile != kTile) {   // tile boundaries: once per matrix and tile size
    if (A->split) cudaFree(A->split);
    A->split = nullptr;
    A->split_tile = 0;
    FSB_CUDA(cudaMalloc(&A->split, ((size_t)ntiles + 1) * sizeof(int)));
    merge_splits_kernel<<<(ntiles + 256) / 256, 256, 0, st>>>(A->nrow, A->nnz, A->row_ptr, kTile, ntiles, A->split);
    FSB_KERNEL_CHECK();
    A->split_tile = kTile;
  }
  // R = 1: the TMA-fed form (knob "stream_tma", 0 = the per-thread-load form below)
  if (RT == 1 && fsb_knob("stream_tma", 1) && !fsb_knob("stream_policy", 0) && ((uintptr_t)A->cols & 15) == 0 &&
      (!VALS || ((uintptr_t)A->vals & 15) == 0)) {
    // built for 6 CTAs per SM (40 registers) whatever the operand size.  The build for 8 (32 registers) is 1.6x SLOWER:
    // eight 24.7 KB CTAs make the driver carve 228 KB of the SM's 256 KB out as shared memory, and a gather that misses L1
    // holds a line there until its data returns -- with ~28 KB of L1 left the gathers in flight, hence the gather rate,
    // collapse (profiles/r2w_l1_carveout.md).  Knob "stream_carveout": explicit carve-out in percent (-1: driver's choice).
    const int minb = fsb_knob("stream_tma_minb", 6);
    const int co = fsb_knob("stream_carveout", -1);
    // gathers through a linear texture over x (knob "stream_tex", default on): the texture pipe's data stage is not the
    // LSU's, which this kernel's shared-memory traffic saturates together with the gathers (LSU data pipe 83 % busy):
    // C3 double SpMV 0.909 -> 0.800 ms, A'x 0.920 -> 0.814 ms, same bits (profiles/r2z_tex_gathers.md)
    int xtex_off = 0;
    const cudaTextureObject_t xtex = fsb_knob("stream_tex", 1) ? fsb_linear_texture(dX, (size_t)A->ncol, 8, st, &xtex_off) : 0;
#define FSB_TMA_LAUNCH2(MINB_, TEX_)                                                                                         \
  do {                                                                                                                       \
    if (co >= 0) cudaFuncSetAttribute(csr_stream_tma_kernel<VALS, MINB_, TEX_>, cudaFuncAttributePreferredSharedMemoryCarveout, co); \
    csr_stream_tma_kernel<VALS, MINB_, TEX_><<<ntiles, kThreads, 0, st>>>(A->nrow, A->nnz, A->row_ptr, A->cols, A->vals, dX, dY, A->split, \
                                                                          carry_row, carry_val, xtex, xtex_off);                       \
  } while (0)
#define FSB_TMA_LAUNCH(MINB_) do { if (xtex) FSB_TMA_LAUNCH2(MINB_, true); else FSB_TMA_LAUNCH2(MINB_, false); } while (0)
    if (minb >= 8) FSB_TMA_LAUNCH(8); else if (minb <= 4) FSB_TMA_LAUNCH(4); else FSB_TMA_LAUNCH(6);
#undef FSB_TMA_LAUNCH2
#undef FSB_TMA_LAUNCH
    FSB_KERNEL_CHECK();
    csr_stream_fixup_kernel<RT><<<(ntiles + 255) / 256, 256, 0, st>>>(ntiles, carry_row, carry_val, dY);
    FSB_KERNEL_CHECK();
    return FSB_OK;
  }
  // bytes of the dense operand a wave of CTAs gathers from: the whole operand, or one block of it for an x-blocked transpose
  const double x_live = A->x_live_bytes > 0 ? (double)A->x_live_bytes : (double)A->ncol * RT * 8.0;
  const bool small_x = x_live <= 34e6;
  // L2 policy words on the loads: measured neutral to slightly negative (C3 SpMV 0.965 -> 1.000 ms, transposed SpMV
  // 1.318 -> 1.320 ms, DRAM bytes unchanged: profiles/r2a_ncu_spmv_policy.md) -- an operand that does not fit the
  // ~60 MB a gather operand gets of B200's L2 is not kept by a hint; the x-blocked transpose fixes that case instead
  const bool pol = fsb_knob("stream_policy", 0) != 0;
#define FSB_STREAM_LAUNCH(MINB_, POL_)                                                                               \
  csr_stream_kernel<RT, VALS, MINB_, POL_><<<ntiles, kThreads, 0, st>>>(A->nrow, A->nnz, A->row_ptr, A->cols, A->vals, dX, dY, \
                                                                        A->split, carry_row, carry_val)
  if (small_x) { if (pol) FSB_STREAM_LAUNCH(6, 1); else FSB_STREAM_LAUNCH(6, 0); }
  else         { if (pol) FSB_STREAM_LAUNCH(4, 1); else FSB_STREAM_LAUNCH(4, 0); }
#undef FSB_STREAM_LAUNCH
  FSB_KERNEL_CHECK();
  csr_stream_fixup_kernel<RT><<<(ntiles + 255) / 256, 256, 0, st>>>(ntiles, carry_row, carry_val, dY);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

}  // namespace

// R = 2 and 4 are implemented (and tested) but the staged row kernel is faster there
// (profiles/r1c_small_R.md): the automatic choice uses the stream kernel for R = 1 only.
bool fsb_csr_stream_supports(int R) { return R == 1 || R == 2 || R == 4; }
bool fsb_csr_stream_preferred(int R) { return R == 1; }

int fsb_launch_csr_stream(fsb_matrix* A, double* dY, const double* dX, int R, cudaStream_t st) {
  if (A->nrow == 0) return FSB_OK;
  switch (R) {
    case 1: return A->has_vals ? launch<1, true>(A, dY, dX, st) : launch<1, false>(A, dY, dX, st);
    case 2: return A->has_vals ? launch<2, true>(A, dY, dX, st) : launch<2, false>(A, dY, dX, st);
    case 4: return A->has_vals ? launch<4, true>(A, dY, dX, st) : launch<4, false>(A, dY, dX, st);
  }
  return fsb_set_error(FSB_EINVAL, "stream kernel: R must be 1, 2 or 4 (got %d)", R);
}

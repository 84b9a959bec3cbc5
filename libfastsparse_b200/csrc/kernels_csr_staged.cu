// kernels_csr_staged.cu -- CSR SpMM / SpMV, "staged row block" kernel (sm_100a).
//
// Same products as kernels_csr.cu (bcsr_A_mul_B* csr.h:149-302, csr_A_mul_B* csr.h:425-465).
// The first kernel (one team per row) is latency bound on B200: ncu shows DRAM at ~53 %
// with long-scoreboard stalls, because every row is a chain of three dependent global
// loads (row_ptr -> cols -> X rows) and each team has little in flight
// (profiles/r1a_ncu_c2_spmm_details.md).  This kernel breaks the chain:
//
//   * a CTA owns RB consecutive rows; it first brings their row_ptr slice and the whole
//     contiguous run of column indices (and values) into SHARED MEMORY -- the index run by one
//     TMA bulk copy (cp.async.bulk + mbarrier; binary matrices), values by fully coalesced
//     streaming loads -- one exposed latency per RB rows instead of per row;
//   * then every sub-group of G lanes walks its rows with the indices already on chip:
//     U independent gathers of the dense operand are issued back to back (each lane one
//     vector LDG of VEC doubles; VEC = 4 is a single 256-bit LDG.E.256), so a warp keeps
//     (32/G)*U X rows in flight;
//   * a sub-group sums a row's terms strictly in stored order, i.e. in the reference's
//     own order: the binary product is bit-identical to the serial reference;
//   * rows too long for the staging buffer are processed by the whole CTA (fixed
//     chunking + shared-memory reduction, still deterministic);
//   * two builds of the same body (lean: 32 registers, full occupancy; deep: all U gathers
//     of a batch in flight) and one or two column passes are chosen per handle by timing
//     (kernels_csr.cu); with two lanes per row it is also the binary SpMV kernel.
#include <stdint.h>

#include <algorithm>

#include "fsb_device.cuh"
#include "fsb_internal.h"

using namespace fsbdev;

namespace {

constexpr int kThreads = 256;
constexpr int kLongRow = 2048;   // rows at least this long are split across the CTA

// ---- two builds of the same body, chosen per matrix at run time (fsb_launch_csr_spmm times both once):
//   LEAN  __launch_bounds__(256): ptxas budgets 32 registers (8 CTAs = 64 warps per SM) and, to fit,
//         interleaves the gathers of a batch with their adds -- about three gathers in flight per lane;
//   DEEP  __launch_bounds__(256, 4): 54 registers, 4 CTAs per SM, all U gathers of a batch issued
//         back to back (an empty volatile asm after the loads pins that order; SASS checked).
// Same box, C2 (profiles/r1f_sweep_gather_issue.md): one pass 5.63 ms either way, two column passes
// 5.25 (lean) vs 5.10 ms (deep); power-law columns (C4) 2.98 (lean) vs 3.21 ms (deep): hot X rows
// hit in L1/L2 and occupancy wins over depth there.  Neither saturates a unit (DRAM 62-81 %,
// L2 51 %): the product sits on the random-gather throughput of L2 + HBM.
// Round 2: the gather probe (profiles/r2_gather_ceiling.md) runs fastest with FEWER loads in flight per lane and more
// resident warps once the operand misses L2; same-box sweep of the deep build (profiles/r2f_sweep_*): U=8 / 4 CTAs per SM
// (54 registers) 4.86 ms at C2 and 3.17 ms at C4, U=6 / 5 CTAs (46 registers) 4.73 and 3.02 ms, U=4 / 6 CTAs 4.94 and
// 2.99 ms, U=4 / 8 CTAs 4.84 and 3.54 ms.  Binary matrices take U=6 / 5; matrices with values keep U=8 / 4.
#ifndef FSB_STAGED_U
#define FSB_STAGED_U 6        // gathers per batch and lane, binary matrices
#endif
#ifndef FSB_STAGED_U_VALS
#define FSB_STAGED_U_VALS 8   // same, matrices with values
#endif
#ifndef FSB_STAGED_DEEP_MINB
#define FSB_STAGED_DEEP_MINB 5
#endif
#ifndef FSB_STAGED_DEEP_MINB_VALS      // matrices with values carry U more doubles per batch
#define FSB_STAGED_DEEP_MINB_VALS 4
#endif
// FSB_STAGED_TMA 1 (default): the contiguous run of column indices of a CTA's rows is brought into shared
// memory by one TMA bulk copy (cp.async.bulk, SASS UBLKCP, completion on an mbarrier) instead of
// per-thread coalesced loads + shared stores.  The run starts at an arbitrary entry, so the copy starts
// at the enclosing 16-byte boundary and the indices are read at a shift of 0..3 entries.  Same box, C2:
// 5.15 -> 5.05 ms (deep, two passes), 5.44 -> 5.09 ms (lean, two passes); C4 3.08 -> 2.95 ms; R = 8
// 1.08 -> 1.03 ms (profiles/r1k_sweep_tma_staging.md).  Matrices with values keep the per-thread
// staging: with a second bulk copy the deep build spills and loses 4 %.
#ifndef FSB_STAGED_TMA
#define FSB_STAGED_TMA 1
#endif
// experiment: matrices with values bring their indices by TMA too (the values stay per-thread loads);
// measured equal to plain per-thread staging (6.46 vs 6.46-6.50 ms at C2 with values), so it stays off
#ifndef FSB_STAGED_TMA_VALS
#define FSB_STAGED_TMA_VALS 0
#endif

// Empty volatile asm that takes a loaded row piece in and out: volatile asms keep their order, so
// placing these after the U gather asms keeps "all U loads, then the adds" in the emitted code.
template <int VEC> __device__ __forceinline__ void pin_after_loads(double (&x)[VEC]);
template <> __device__ __forceinline__ void pin_after_loads<1>(double (&x)[1]) { asm volatile("" : "+d"(x[0])); }
template <> __device__ __forceinline__ void pin_after_loads<2>(double (&x)[2]) { asm volatile("" : "+d"(x[0]), "+d"(x[1])); }
template <> __device__ __forceinline__ void pin_after_loads<4>(double (&x)[4]) {
  asm volatile("" : "+d"(x[0]), "+d"(x[1]), "+d"(x[2]), "+d"(x[3]));
}

// gather of VEC doubles at element offset `off` (doubles) through a linear texture of 8-byte (VEC = 1) or 16-byte texels
template <int VEC> __device__ __forceinline__ void tex_gather(double (&v)[VEC], cudaTextureObject_t t, int off);
template <> __device__ __forceinline__ void tex_gather<1>(double (&v)[1], cudaTextureObject_t t, int off) {
  const int2 w = tex1Dfetch<int2>(t, off);
  v[0] = __hiloint2double(w.y, w.x);
}
template <> __device__ __forceinline__ void tex_gather<2>(double (&v)[2], cudaTextureObject_t t, int off) {
  const int4 w = tex1Dfetch<int4>(t, off >> 1);
  v[0] = __hiloint2double(w.y, w.x); v[1] = __hiloint2double(w.w, w.z);
}
template <> __device__ __forceinline__ void tex_gather<4>(double (&v)[4], cudaTextureObject_t t, int off) {
  const int4 a = tex1Dfetch<int4>(t, off >> 1), b = tex1Dfetch<int4>(t, (off >> 1) + 1);
  v[0] = __hiloint2double(a.y, a.x); v[1] = __hiloint2double(a.w, a.z);
  v[2] = __hiloint2double(b.y, b.x); v[3] = __hiloint2double(b.w, b.z);
}

// One row, summed strictly in stored order, gathers issued in batches of U.
template <int G, int VEC, bool VALS, bool FROM_SMEM, bool DEEP, int MODE>
__device__ __forceinline__ void walk_row(const int* __restrict__ ci, const double* __restrict__ vi, int s, int e,
                                         double (&acc)[VEC], const double* __restrict__ xbase, int ldx, bool col_ok,
                                         unsigned long long xpol, cudaTextureObject_t xtex, int xoff) {
  constexpr int U = VALS ? FSB_STAGED_U_VALS : FSB_STAGED_U;
  constexpr bool TEX = MODE == 1;
  for (int i = s; i < e; i += U) {
    double xr[U][VEC];
    double vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = i + u;
      if (idx < e && col_ok) {
        const int c = FROM_SMEM ? ci[idx] : __ldg(ci + idx);
        if (VALS) vv[u] = FROM_SMEM ? vi[idx] : __ldg(vi + idx);
        if (TEX) tex_gather<VEC>(xr[u], xtex, c * ldx + xoff);
        else XLoad<VEC>::ldp(xr[u], xbase + (long long)c * ldx, xpol);
      } else {
        if (VALS) vv[u] = 0.0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) xr[u][v] = 0.0;
      }
    }
    if (DEEP) {
#pragma unroll
      for (int u = 0; u < U; ++u) pin_after_loads<VEC>(xr[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = VALS ? fma(xr[u][v], vv[u], acc[v]) : acc[v] + xr[u][v];
  }
}

// optional fused epilogue: acc += lambda * Z[row, cols]  (the "+ lambda P" of the CG operator, cg.h:19-21)
template <int VEC>
__device__ __forceinline__ void add_scaled_row(double (&acc)[VEC], const double* __restrict__ Z, double lambda, long long off) {
  if (Z == nullptr) return;
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = fma(lambda, __ldg(Z + off + v), acc[v]);
}

template <int G, int VEC, bool VALS, bool DEEP, int MODE>
__device__ __forceinline__ void staged_body(int nrow, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                                            const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                                            int R, int col0, int ncols, int RB, int CAP, int l2mode,
                                            const double* __restrict__ Z, double lambda, int ldx, int xcol0, cudaTextureObject_t xtex, int xtex_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [row_ptr: RB+1 ints, padded to 16 B] [vals: CAP doubles (VALS)] [cols: CAP ints];
  // the long-row reduction buffer (kThreads*VEC doubles) aliases the vals/cols region
  int* s_rp = reinterpret_cast<int*>(smem_raw);
  unsigned char* body = smem_raw + (((size_t)(RB + 1) * 4 + 15) & ~(size_t)15);
  double* s_vals = reinterpret_cast<double*>(body);
  int* s_cols = reinterpret_cast<int*>(body + (VALS ? (size_t)(CAP + 16) * 8 : 0));
  double* s_red = reinterpret_cast<double*>(body);

  constexpr int NT = kThreads / G;   // sub-groups per CTA
  const int tid = threadIdx.x;
  const int team = tid / G, l = tid & (G - 1);
  const int r0 = blockIdx.x * RB;
  const int nr = min(RB, nrow - r0);
  const bool col_ok = l * VEC < ncols;
  // the dense operand may be a column slab stored on its own: row stride ldx, first column xcol0
  const double* xbase = X + xcol0 + l * VEC;
  const int xoff = xtex_off + xcol0 + l * VEC;      // the same offset in doubles (from the texture's first texel), for the texture form of the gather
  // l2mode 1: dense operand evict_last, matrix stream evict_first (keep X resident in L2)
  const unsigned long long xpol = make_l2_policy(l2mode);   // 1 = evict_last (default), 0 = none, 3..6 = fractional experiments
  const unsigned long long spol = make_l2_policy(l2mode ? 2 : 0);

  constexpr bool kTma = FSB_STAGED_TMA && (!VALS || FSB_STAGED_TMA_VALS);
  __shared__ __align__(8) unsigned long long s_bar;
  if (kTma && tid == 0) mbar_init(&s_bar, 1);
  for (int i = tid; i <= nr; i += kThreads) s_rp[i] = __ldg(row_ptr + r0 + i);
  __syncthreads();
  const int base = s_rp[0];
  const int total = s_rp[nr] - base;

  if (total <= CAP) {
    const int* ci = s_cols;
    const double* vi = s_vals;
    if constexpr (kTma) {
      // one bulk copy from the enclosing 16-byte boundary (the buffer has 8 entries of slack for that)
      const int shc = base & 3;
      if (total > 0) {
        if (tid == 0) {
          const unsigned bytes = (unsigned)((shc + total + 3) & ~3) * 4u;
          mbar_expect_tx(&s_bar, bytes);
          tma_load_1d(s_cols, cols + (base - shc), bytes, &s_bar, spol);
        }
        if (VALS) {   // values by per-thread streaming loads while the bulk copy is in flight
          for (int i = tid; i < total; i += kThreads) s_vals[i] = ld_stream_f64_pol(vals + base + i, spol);
          __syncthreads();
        }
        mbar_wait(&s_bar, 0);
      }
      ci = s_cols + shc;
    } else {
      for (int i = tid; i < total; i += kThreads) {
        s_cols[i] = ld_stream_s32_pol(cols + base + i, spol);
        if (VALS) s_vals[i] = ld_stream_f64_pol(vals + base + i, spol);
      }
      __syncthreads();
    }
    for (int r = team; r < nr; r += NT) {
      double acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
      walk_row<G, VEC, VALS, true, DEEP, MODE>(ci, vi, s_rp[r] - base, s_rp[r + 1] - base, acc, xbase, ldx, col_ok, xpol, xtex, xoff);
      if (col_ok) {
        const long long off = (long long)(r0 + r) * R + col0 + l * VEC;
        add_scaled_row<VEC>(acc, Z, lambda, off);
        YStore<VEC>::st(Y + off, acc);
      }
    }
    return;
  }

  // overflow: at least one long row in this block.  Short rows go to sub-groups as usual
  // (indices read from global memory); long rows are split across all sub-groups.
  __syncthreads();   // s_rp is read below while s_red (aliased with vals/cols, not s_rp) is written
  for (int r = 0; r < nr; ++r) {
    const int s = s_rp[r], e = s_rp[r + 1];
    if (e - s >= kLongRow) {
      const int chunk = (e - s + NT - 1) / NT;
      const int cs = min(e, s + team * chunk), ce = min(e, cs + chunk);
      double acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
      walk_row<G, VEC, VALS, false, DEEP, MODE>(cols, vals, cs, ce, acc, xbase, ldx, col_ok, xpol, xtex, xoff);
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_red[(team * G + l) * VEC + v] = acc[v];
      __syncthreads();
      if (team == 0 && col_ok) {
        double tot[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) tot[v] = 0.0;
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int v = 0; v < VEC; ++v) tot[v] += s_red[(t * G + l) * VEC + v];
        const long long off = (long long)(r0 + r) * R + col0 + l * VEC;
        add_scaled_row<VEC>(tot, Z, lambda, off);
        YStore<VEC>::st(Y + off, tot);
      }
      __syncthreads();
    } else if (r % NT == team) {
      double acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
      walk_row<G, VEC, VALS, false, DEEP, MODE>(cols, vals, s, e, acc, xbase, ldx, col_ok, xpol, xtex, xoff);
      if (col_ok) {
        const long long off = (long long)(r0 + r) * R + col0 + l * VEC;
        add_scaled_row<VEC>(acc, Z, lambda, off);
        YStore<VEC>::st(Y + off, acc);
      }
    }
  }
}

template <int G, int VEC, bool VALS, int MODE>
__global__ void __launch_bounds__(kThreads)
csr_spmm_staged_kernel(int nrow, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                       const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                       int R, int col0, int ncols, int RB, int CAP, int l2mode, const double* __restrict__ Z, double lambda,
                       int ldx, int xcol0, cudaTextureObject_t xtex, int xtex_off) {
  staged_body<G, VEC, VALS, false, MODE>(nrow, row_ptr, cols, vals, X, Y, R, col0, ncols, RB, CAP, l2mode, Z, lambda, ldx, xcol0, xtex, xtex_off);
}

template <int G, int VEC, bool VALS, int MODE>
__global__ void __launch_bounds__(kThreads, VALS ? FSB_STAGED_DEEP_MINB_VALS : FSB_STAGED_DEEP_MINB)
csr_spmm_staged_deep_kernel(int nrow, const int* __restrict__ row_ptr, const int* __restrict__ cols,
                            const double* __restrict__ vals, const double* __restrict__ X, double* __restrict__ Y,
                            int R, int col0, int ncols, int RB, int CAP, int l2mode, const double* __restrict__ Z, double lambda,
                            int ldx, int xcol0, cudaTextureObject_t xtex, int xtex_off) {
  staged_body<G, VEC, VALS, true, MODE>(nrow, row_ptr, cols, vals, X, Y, R, col0, ncols, RB, CAP, l2mode, Z, lambda, ldx, xcol0, xtex, xtex_off);
}

thread_local int g_rb = 0, g_cap_mult = 0, g_l2mode = 1;
thread_local bool g_use_tex = false;   // set per launch by fsb_launch_csr_spmm_staged

template <int G, int VEC, bool VALS>
int launch(const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, int RB, int CAP, cudaStream_t st,
           const double* dZ, double lambda, bool deep, int ldx, int xcol0) {
  // the TMA staging copies from the 16-byte boundary below a row's first entry: the index array itself must sit on one
  // (every array the library allocates does)
  if (FSB_STAGED_TMA && (!VALS || FSB_STAGED_TMA_VALS) && ((uintptr_t)A->cols & 15))
    return fsb_set_error(FSB_EINVAL, "staged SpMM: column index array is not 16-byte aligned");
  // gathers through a linear texture over the dense operand (knob "staged_tex": 1 = on, 0 = off; default: narrow
  // operands, see fsb_launch_csr_spmm_staged).  No L2 policy word exists for texture fetches, so operands that rely on
  // evict_last (the half-resident 256 MB operand of C2) keep the LDG form.
  cudaTextureObject_t xtex = 0;
  int xtex_off = 0;   // first texel of the operand in the texture (which starts on the 512-byte boundary below it), in doubles
  if (g_use_tex) {
    const size_t xd = (size_t)A->ncol * (size_t)ldx;            // doubles in the operand (or its repacked slab)
    const double* xorigin = dX;                                   // texel 0; gathers add c * ldx + xcol0 + l * VEC doubles
    if (xd + 32 < ((size_t)1 << 31) && (VEC == 1 || (ldx % 2 == 0 && xcol0 % 2 == 0)))
      xtex = fsb_linear_texture(xorigin, VEC == 1 ? xd : xd / 2, VEC == 1 ? 8 : 16, st, &xtex_off);
    if (VEC != 1) xtex_off *= 2;
  }
  size_t body = std::max((size_t)(CAP + 16) * (VALS ? 12 : 4), (size_t)kThreads * VEC * 8);   // staging (+ alignment / vector-read slack) or long-row reduction
  size_t smem = ((((size_t)RB + 1) * 4 + 15) & ~(size_t)15) + ((body + 15) & ~(size_t)15);
  const unsigned grid = (unsigned)((A->nrow + RB - 1) / RB);
  const int co = fsb_knob("staged_carveout", -1);   // experiment knob: shared-memory carve-out in percent of the maximum (-1: the driver's choice)
#define FSB_STAGED_GO(KERN_)                                                                                                  \
  do {                                                                                                                        \
    if (smem > 48 * 1024) FSB_CUDA(cudaFuncSetAttribute(KERN_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    if (co >= 0) cudaFuncSetAttribute(KERN_, cudaFuncAttributePreferredSharedMemoryCarveout, co);                             \
    KERN_<<<grid, kThreads, smem, st>>>(A->nrow, A->row_ptr, A->cols, A->vals, dX, dY, R, col0, ncols, RB, CAP, g_l2mode, dZ, lambda, \
                                        ldx, xcol0, xtex, xtex_off);                                                          \
  } while (0)
  // MODE: 0 = LDG gathers, 1 = texture gathers.  (A mode 2 -- column indices by one LDS per chunk of entries + warp shuffles
  // instead of one broadcast LDS per gather, to take the index reads off the LSU data pipe -- was measured and removed:
  // C4 3.03 -> 4.70 ms, C2 5.00 -> 6.62 ms, profiles/r2z_idx_shfl_probe.jsonl.)
  if (xtex) { if (deep) FSB_STAGED_GO((csr_spmm_staged_deep_kernel<G, VEC, VALS, 1>)); else FSB_STAGED_GO((csr_spmm_staged_kernel<G, VEC, VALS, 1>)); }
  else      { if (deep) FSB_STAGED_GO((csr_spmm_staged_deep_kernel<G, VEC, VALS, 0>)); else FSB_STAGED_GO((csr_spmm_staged_kernel<G, VEC, VALS, 0>)); }
#undef FSB_STAGED_GO
  return FSB_OK;
}

template <int G, int VEC>
int launch_v(const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, int RB, int CAP, cudaStream_t st,
             const double* dZ, double lambda, bool deep, int ldx, int xcol0) {
  return A->has_vals ? launch<G, VEC, true>(A, dY, dX, R, col0, ncols, RB, CAP, st, dZ, lambda, deep, ldx, xcol0)
                     : launch<G, VEC, false>(A, dY, dX, R, col0, ncols, RB, CAP, st, dZ, lambda, deep, ldx, xcol0);
}

template <int G>
int launch_g(int vec, const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols, int RB, int CAP, cudaStream_t st,
             const double* dZ, double lambda, bool deep, int ldx, int xcol0) {
  switch (vec) {
    case 1: return launch_v<G, 1>(A, dY, dX, R, col0, ncols, RB, CAP, st, dZ, lambda, deep, ldx, xcol0);
    case 2: return launch_v<G, 2>(A, dY, dX, R, col0, ncols, RB, CAP, st, dZ, lambda, deep, ldx, xcol0);
    default: return launch_v<G, 4>(A, dY, dX, R, col0, ncols, RB, CAP, st, dZ, lambda, deep, ldx, xcol0);
  }
}

}  // namespace

// where the texture form of the gather is the default (profiles/r2z_tex_gathers.md)
// binary SpMV and R = 2 / 4 gain 1.4-2 % (0.752 -> 0.741, 0.764 -> 0.753, 0.816 -> 0.800 ms at C3's structure); gathers of
// 64 bytes and more per row lose 20-40 % (R = 8: 1.07 -> 1.29 ms, C4 R = 32: 3.25 -> 4.33 ms, C2: 5.18 -> 6.10 ms), and so
// do matrices with values on this kernel
static bool staged_tex_auto(const fsb_matrix* A, int R, int ncols, int vec) {
  (void)ncols; (void)vec;
  return !A->has_vals && R <= 4;
}

void fsb_csr_staged_set_tuning(int rb, int cap_mult) {
  g_rb = rb;
  g_cap_mult = cap_mult % 100;           // hundreds digit of cap_mult selects the L2 policy experiment:
  const int h = cap_mult / 100;                   // 1xx = no cache hints, 3xx..6xx = fractional evict_last on X (fsb_device.cuh),
  g_l2mode = h == 1 ? 0 : (h >= 3 && h <= 6) ? h : 1;   // otherwise X evict_last / stream evict_first
}

// one pass over columns [col0, col0+ncols) with sub-groups of g lanes x vec doubles
int fsb_launch_csr_spmm_staged(const fsb_matrix* A, double* dY, const double* dX, int R, int col0, int ncols,
                               int g, int vec, cudaStream_t st, const double* dZ, double lambda, bool deep, int ldx, int xcol0) {
  if (ldx <= 0) { ldx = R; xcol0 = col0; }   // the usual case: X is [ncol][R] like Y
  // rows per CTA: enough rows that every sub-group gets a few, bounded so that the index
  // run (~avg_nnz * RB entries) stays a small shared-memory footprint (several CTAs per SM)
  const int nt = kThreads / g;
  int rb = g_rb ? g_rb : std::max(nt * 4, 128);
  const double avg = std::max(A->avg_row_nnz, 1.0);
  const double budget = A->has_vals ? 2048.0 : 4096.0;     // staged entries per CTA (12 or 4 bytes each)
  while (rb > nt && rb > 8 && avg * rb > budget) rb /= 2;
  rb = std::max(rb, 1);
  int cap = (int)std::min<double>(avg * rb * (g_cap_mult ? g_cap_mult : 1.5) + 256, 3.0 * budget);
  cap = std::max((cap + 63) & ~63, 512);
  {   // texture-pipe gathers: knob "staged_tex" 1 = always, 0 = never, -1 (default) = where measured faster
    const int kt = fsb_knob("staged_tex", -1);
    g_use_tex = kt == 1 || (kt < 0 && staged_tex_auto(A, R, ncols, vec));
  }
  int rc;
  switch (g) {
    case 1: rc = launch_g<1>(vec, A, dY, dX, R, col0, ncols, rb, cap, st, dZ, lambda, deep, ldx, xcol0); break;
    case 2: rc = launch_g<2>(vec, A, dY, dX, R, col0, ncols, rb, cap, st, dZ, lambda, deep, ldx, xcol0); break;
    case 4: rc = launch_g<4>(vec, A, dY, dX, R, col0, ncols, rb, cap, st, dZ, lambda, deep, ldx, xcol0); break;
    case 8: rc = launch_g<8>(vec, A, dY, dX, R, col0, ncols, rb, cap, st, dZ, lambda, deep, ldx, xcol0); break;
    case 16: rc = launch_g<16>(vec, A, dY, dX, R, col0, ncols, rb, cap, st, dZ, lambda, deep, ldx, xcol0); break;
    default: rc = launch_g<32>(vec, A, dY, dX, R, col0, ncols, rb, cap, st, dZ, lambda, deep, ldx, xcol0); break;
  }
  FSB_TRY(rc);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

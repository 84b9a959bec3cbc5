// kernels_dense.cu -- dense building blocks of the block conjugate-gradient solver (sm_100a).
//
// Replaces the OpenMP reductions and vector loops of the reference:
//   pnormsq/pnormsq2/pouter2/pdot/pdot2sym linalg.h:15-73  -> gram_partial_kernel (+ reduction)
//   solve2sym linalg.h:77-88                               -> small_solve_kernel (R x R Cholesky)
//   update loops cg.h:60-63,70-73,148-154,165-170,176-179  -> cg_mix_kernel<MODE>
// All operands are tall-skinny row-major [n][R], R <= 32.  Every pass streams its operands
// from HBM exactly once; the arithmetic is R x R per row (2 GFLOP per pass at n = 1M, R = 32),
// which scalar FMAs through shared memory cannot feed at HBM speed (operand broadcasts saturate
// the LDS pipe).  The products therefore run on the fp64 tensor-core path, mma.sync m8n8k4
// (DMMA; tcgen05 has no fp64 kind), with fragments loaded STRAIGHT from global memory:
//   * Gram  G = Xa' Xb : the long dimension (rows) is the MMA k dimension.  Lane (g, t) of a warp
//     loads Xa[r0 + t][cols of slot g] -- for R = 32 two 128-bit loads per row, a warp covers four
//     full 128-byte lines -- and the same registers serve as A fragment (Xa') and B fragment (Xb);
//   * row mix  O (op)= I M : the R x R coefficient matrix lives in registers as B fragments for the
//     whole kernel; lane (g, t) loads 64 contiguous bytes of row r0 + g of I (the MMA k index is
//     permuted to make that contiguous) and the 8 x 8 accumulator tiles map to 16-byte pieces of
//     O's rows.  No shared-memory staging; the next 8-row block is prefetched into registers.
// Reductions use a fixed grid and a fixed combination order, so every rank of a multi-GPU
// solve gets bit-identical Gram matrices and takes the same branches.
// Kernels of the solver loop are predicated on the device-side status words (converged /
// breakdown), so the host may enqueue several iterations before it looks at the flags.
#include <stdint.h>

#include <algorithm>

#include "fsb_dense.h"
#include "fsb_internal.h"

namespace {

// experiment knob: minimum resident CTAs per SM promised to ptxas for the Gram / row-mix kernels (1 = let it use
// 150-200 registers, one CTA of 8 warps per SM; 2 = cap at 128 registers, which spills 100-470 bytes per thread
// and measured 0.1-0.4 ms slower per C5 iteration)
#ifndef FSB_DENSE_MINB
#define FSB_DENSE_MINB 1
#endif
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSms = 148;
constexpr int kMaxParts = kSms * 4;   // upper bound of the partial-Gram slots any launch uses

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ bool solver_stopped(const int* status) { return status && ((status[0] | status[1]) != 0); }

// column of slot i (0..7) in 8-column block blk of a Gram operand.  V32 (R == 32, 16-byte aligned
// rows): a lane's four slots are two adjacent column pairs, so they come from two 128-bit loads.
template <bool V32>
__device__ __forceinline__ int gram_col(int blk, int i) {
  return V32 ? 2 * i + (blk & 1) + 16 * (blk >> 1) : 8 * blk + i;
}

template <int NB, bool V32>
__device__ __forceinline__ void load_gram_frag(double (&f)[NB], const double* X, long long row, long long n, int R, int g) {
  if constexpr (V32) {
    static_assert(!V32 || NB == 4, "V32 needs R == 32");
    if (row < n) {
      const double2 lo = __ldg(reinterpret_cast<const double2*>(X + row * 32 + 2 * g));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(X + row * 32 + 16 + 2 * g));
      f[0] = lo.x; f[1] = lo.y; f[2] = hi.x; f[3] = hi.y;
    } else {
      f[0] = f[1] = f[2] = f[3] = 0.0;
    }
  } else {
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
      const int col = 8 * blk + g;
      f[blk] = (row < n && col < R) ? __ldg(X + row * R + col) : 0.0;
    }
  }
}

// Combine the per-warp Gram accumulators of a CTA in a fixed order (two rounds of four warps
// through shared memory) and write partial[cta][R*R].  acc[ab][bb][e] of lane (g, t) holds
// G[col(ab, g)][col(bb, 2t + e)]; UPPER: only blocks ab <= bb were accumulated (symmetric Gram),
// entries are mirrored.
template <int NB, bool V32, bool UPPER>
__device__ __forceinline__ void reduce_gram_cta(double* red, const double (&acc)[NB][NB][2], double* __restrict__ partial, int R) {
  constexpr int W = NB * 8;
  constexpr int PER = (W * W + kThreads - 1) / kThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  double tot[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) tot[k] = 0.0;
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    if ((warp >> 2) == round) {
      double* dst = red + (warp & 3) * W * W;
#pragma unroll
      for (int ab = 0; ab < NB; ++ab)
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          if (UPPER && bb < ab) continue;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int a = gram_col<V32>(ab, g), b = gram_col<V32>(bb, 2 * t + e);
            if (UPPER) {   // off-diagonal blocks fill their mirror block; diagonal blocks mirror their upper half
              if (ab != bb || a <= b) { dst[a * W + b] = acc[ab][bb][e]; dst[b * W + a] = acc[ab][bb][e]; }
            } else {
              dst[a * W + b] = acc[ab][bb][e];
            }
          }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int e = tid + k * kThreads;
      if (e < W * W) tot[k] += (red[e] + red[W * W + e]) + (red[2 * W * W + e] + red[3 * W * W + e]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int e = tid + k * kThreads;
    if (e < W * W) {
      const int a = e / W, b = e % W;
      if (a < R && b < R) partial[(size_t)blockIdx.x * R * R + a * R + b] = tot[k];
    }
  }
}

// partial[cta][R*R] = sum over this CTA's rows of Xa[i][a] * Xb[i][b]   (SYM: Xb == Xa)
template <int NB, bool V32, bool SYM>
__global__ void __launch_bounds__(kThreads, FSB_DENSE_MINB)
gram_partial_kernel(double* __restrict__ partial, const double* Xa, const double* Xb, long long n, int R) {
  __shared__ double red[4 * NB * 8 * NB * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  double acc[NB][NB][2];
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const long long nsteps = (n + 7) / 8, stride = (long long)gridDim.x * kWarps;
  long long step = (long long)blockIdx.x * kWarps + warp;
  double a[2][NB], b[2][NB], an[2][NB], bn[2][NB];
  if (step < nsteps) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      load_gram_frag<NB, V32>(a[s], Xa, step * 8 + 4 * s + t, n, R, g);
      if (!SYM) load_gram_frag<NB, V32>(b[s], Xb, step * 8 + 4 * s + t, n, R, g);
    }
  }
  for (; step < nsteps; step += stride) {
    const long long nx = step + stride;
    if (nx < nsteps) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        load_gram_frag<NB, V32>(an[s], Xa, nx * 8 + 4 * s + t, n, R, g);
        if (!SYM) load_gram_frag<NB, V32>(bn[s], Xb, nx * 8 + 4 * s + t, n, R, g);
      }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int ab = 0; ab < NB; ++ab)
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) {
          if (SYM && bb < ab) continue;
          dmma884(acc[ab][bb][0], acc[ab][bb][1], a[s][ab], SYM ? a[s][bb] : b[s][bb]);
        }
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int k = 0; k < NB; ++k) { a[s][k] = an[s][k]; if (!SYM) b[s][k] = bn[s][k]; }
  }
  reduce_gram_cta<NB, V32, SYM>(red, acc, partial, R);
}

// G[e] = sum over the partial Gram matrices, second stage of the reductions above.  A CTA owns 32
// consecutive elements; its eight warps each sum every eighth partial (four interleaved running sums:
// four loads in flight per thread) and are combined through shared memory in a fixed tree, so the
// result does not depend on scheduling.  (A first version with one thread per element and 4 CTAs
// took 130 us cold: a chain of ~300 dependent loads per thread.)
constexpr int kFinalThreads = 256;
__global__ void __launch_bounds__(kFinalThreads)
gram_final_kernel(double* __restrict__ G, const double* __restrict__ partial, int nparts, int RR) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (e < RR) {
    int p = grp;
    for (; p + 24 < nparts; p += 32) {
      s0 += partial[(size_t)p * RR + e];
      s1 += partial[(size_t)(p + 8) * RR + e];
      s2 += partial[(size_t)(p + 16) * RR + e];
      s3 += partial[(size_t)(p + 24) * RR + e];
    }
    for (; p < nparts; p += 8) s0 += partial[(size_t)p * RR + e];
  }
  red[grp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (grp == 0 && e < RR)
    G[e] = ((red[0][lane] + red[1][lane]) + (red[2][lane] + red[3][lane])) + ((red[4][lane] + red[5][lane]) + (red[6][lane] + red[7][lane]));
}
inline int final_grid(int RR) { return (RR + 31) / 32; }

__global__ void axpy_lambda_kernel(double* __restrict__ Y, const double* __restrict__ X, double lambda, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) Y[i] = fma(lambda, X[i], Y[i]);
}

// X = 0, Rm = P = B * diag(inorm)     (cg.h:44-48, 112-120)
__global__ void cg_init_kernel(double* __restrict__ X, double* __restrict__ Rm, double* __restrict__ P,
                               const double* __restrict__ B, const double* __restrict__ inorm, long long n, int R) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const double r = B[i] * inorm[i % R];
    X[i] = 0.0;
    Rm[i] = r;
    P[i] = r;
  }
}

// norms from the Gram diagonal: norm[k] = sqrt(G[k][k]) (or 1 when !normalise)
__global__ void cg_norms_kernel(double* __restrict__ norm, double* __restrict__ inorm, const double* __restrict__ G, int R, int normalise) {
  const int k = threadIdx.x;
  if (k >= R) return;
  const double nv = normalise ? sqrt(G[k * R + k]) : 1.0;
  norm[k] = nv;
  inorm[k] = 1.0 / nv;
}

// ---------------------------------------------------------------------------------------------
// O[i,:] (op)= I[i,:] * M for a tall row-major [n][R] operand and an R x R matrix M
// (M[k][j] = coefficient of input column k in output column j):
//   MODE 0: O += I M                      (X += P alpha,      cg.h:148-151)
//   MODE 1: O -= I M, partial = O' O      (R -= KP alpha cg.h:152-153, fused with R'R cg.h:156)
//   MODE 2: O  = Add + I M, I may alias O (P = R + P psi,     cg.h:165-170)
// One warp owns 8-row blocks: lane (g, t) holds row r0 + g.
template <int NB, bool V32>
struct MixFrag {
  static constexpr int NKS = 2 * NB;   // k steps of 4 input columns
  // input column handled by lane t at k step s
  static __device__ __forceinline__ int kcol(int s, int t) { return V32 ? 8 * t + s : 4 * s + t; }

  static __device__ __forceinline__ void load_m(double (&bf)[NKS][NB], const double* __restrict__ M, int R, int g, int t) {
#pragma unroll
    for (int s = 0; s < NKS; ++s)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int k = kcol(s, t), j = 8 * nb + g;
        bf[s][nb] = (k < R && j < R) ? M[k * R + j] : 0.0;
      }
  }
  template <bool NC>
  static __device__ __forceinline__ void load_in(double (&av)[NKS], const double* I, long long row, long long n, int R, int t) {
    if constexpr (V32) {
      if (row < n) {
        const double2* p = reinterpret_cast<const double2*>(I + row * 32 + 8 * t);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 v = NC ? __ldg(p + q) : p[q];
          av[2 * q] = v.x; av[2 * q + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int s = 0; s < NKS; ++s) av[s] = 0.0;
      }
    } else {
#pragma unroll
      for (int s = 0; s < NKS; ++s) {
        const int col = 4 * s + t;
        av[s] = (row < n && col < R) ? (NC ? __ldg(I + row * R + col) : I[row * R + col]) : 0.0;
      }
    }
  }
  static __device__ __forceinline__ void load_out(double (&c)[NB][2], const double* O, long long row, long long n, int R, int t) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      if constexpr (V32) {
        if (row < n) {
          const double2 v = *reinterpret_cast<const double2*>(O + row * 32 + 8 * nb + 2 * t);
          c[nb][0] = v.x; c[nb][1] = v.y;
        } else {
          c[nb][0] = c[nb][1] = 0.0;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * nb + 2 * t + e;
          c[nb][e] = (row < n && col < R) ? O[row * R + col] : 0.0;
        }
      }
    }
  }
  static __device__ __forceinline__ void store_out(double* O, const double (&c)[NB][2], long long row, long long n, int R, int t) {
    if (row >= n) return;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      if constexpr (V32) {
        *reinterpret_cast<double2*>(O + row * 32 + 8 * nb + 2 * t) = make_double2(c[nb][0], c[nb][1]);
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * nb + 2 * t + e;
          if (col < R) O[row * R + col] = c[nb][e];
        }
      }
    }
  }
};

constexpr int kScratchStride = 36;   // doubles per scratch row: conflict-free fragment reads, 16-byte aligned rows

template <int NB, bool V32, int MODE>
__global__ void __launch_bounds__(kThreads, FSB_DENSE_MINB)
cg_mix_kernel(double* O, const double* I, const double* Add, const double* __restrict__ M, double* __restrict__ partial,
              long long n, int R, const int* __restrict__ status) {
  using F = MixFrag<NB, V32>;
  constexpr int NKS = F::NKS;
  constexpr int W = NB * 8;
  constexpr int kSmem = MODE == 1 ? (4 * W * W > kWarps * 8 * kScratchStride ? 4 * W * W : kWarps * 8 * kScratchStride) : 1;
  __shared__ __align__(16) double smem[kSmem];
  if (solver_stopped(status)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  double bf[NKS][NB];
  F::load_m(bf, M, R, g, t);
  double gacc[NB][NB][2];   // MODE 1 only (dead otherwise); upper blocks used
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) gacc[i][j][0] = gacc[i][j][1] = 0.0;
  double* scratch = smem + warp * 8 * kScratchStride;
  const double* C0 = MODE == 2 ? Add : O;
  const long long nblk = (n + 7) / 8, stride = (long long)gridDim.x * kWarps;
  long long blk = (long long)blockIdx.x * kWarps + warp;
  double av[NKS], c[NB][2], avn[NKS], cn[NB][2];
  if (blk < nblk) {
    F::template load_in<MODE != 2>(av, I, blk * 8 + g, n, R, t);
    F::load_out(c, C0, blk * 8 + g, n, R, t);
  }
  for (; blk < nblk; blk += stride) {
    const long long nx = blk + stride;
    if (nx < nblk) {
      F::template load_in<MODE != 2>(avn, I, nx * 8 + g, n, R, t);
      F::load_out(cn, C0, nx * 8 + g, n, R, t);
    }
#pragma unroll
    for (int s = 0; s < NKS; ++s)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) dmma884(c[nb][0], c[nb][1], MODE == 1 ? -av[s] : av[s], bf[s][nb]);
    const long long row = blk * 8 + g;
    F::store_out(O, c, row, n, R, t);
    if constexpr (MODE == 1) {
      // Gram of the updated rows: re-distribute the 8 x W tile through per-warp scratch so that
      // lane (g, t) holds rows t, t + 4 of column 8*blk + g (the fragment layout of gram_partial)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const bool ok = row < n;
        // (columns beyond R come from zero-padded operands and are zero already)
        const double2 v = make_double2(ok ? c[nb][0] : 0.0, ok ? c[nb][1] : 0.0);
        *reinterpret_cast<double2*>(scratch + g * kScratchStride + 8 * nb + 2 * t) = v;
      }
      __syncwarp();
      double ga[2][NB];
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int k = 0; k < NB; ++k) ga[s][k] = scratch[(4 * s + t) * kScratchStride + 8 * k + g];
      __syncwarp();
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int ab = 0; ab < NB; ++ab)
#pragma unroll
          for (int bb = ab; bb < NB; ++bb) dmma884(gacc[ab][bb][0], gacc[ab][bb][1], ga[s][ab], ga[s][bb]);
    }
#pragma unroll
    for (int s = 0; s < NKS; ++s) av[s] = avn[s];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) { c[nb][0] = cn[nb][0]; c[nb][1] = cn[nb][1]; }
  }
  if constexpr (MODE == 1) {
    __syncthreads();   // scratch is re-used as the reduction buffer
    reduce_gram_cta<NB, false, true>(smem, gacc, partial, R);
  }
}

// lo[i][0:h] = src[i][0:h], hi[i][0:h] = src[i][h:2h]  (column halves of a [n][2h] operand, 16-byte pieces)
__global__ void split_halves_kernel(double2* __restrict__ lo, double2* __restrict__ hi, const double2* __restrict__ src,
                                    long long n, int h2 /* half width in double2 */, const int* __restrict__ status) {
  if (solver_stopped(status)) return;
  const long long tot = n * 2 * h2;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < tot; i += stride) {
    const long long row = i / (2 * h2);
    const int c = (int)(i - row * 2 * h2);
    const double2 v = src[i];
    if (c < h2) lo[row * h2 + c] = v; else hi[row * h2 + (c - h2)] = v;
  }
}

__global__ void cg_scale_cols_kernel(double* __restrict__ X, const double* __restrict__ norm, long long n, int R) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) X[i] *= norm[i % R];
}

// Solve A M = RHS for M (all R x R, row-major; A symmetric positive definite) by Cholesky
// A = L L' -- in ONE WARP: lane j keeps column j of A and of the right-hand side in registers, the
// current column of L is broadcast through shared memory; no block barriers, every inner loop is a
// run of independent FMAs.  Generalises the
// closed-form 2x2 solve2sym (linalg.h:77-88); matrices are padded to 32 x 32 with the identity.
// (A first version -- 1024 threads, three block barriers per pivot, substitution by dependent
// chains through shared memory -- took 50-75 us; this one is a few microseconds, which matters
// when an iteration is 2.5 ms on 8 GPUs.)
// status[0] is set to 1 when a pivot is not safely positive.  check != 0 also evaluates the
// stopping rule of bsbm_cg2 (cg.h:158): status[1] = 1 when every diagonal entry of RHS is
// <= thr; otherwise status[2] (completed iterations) is incremented.  A stopped solver
// (status[0] or status[1] set) is left untouched.
__global__ void __launch_bounds__(32, 1)
small_solve_kernel(double* __restrict__ M, const double* __restrict__ A, const double* __restrict__ RHS, int R,
                   int* __restrict__ status, int check, double thr) {
  __shared__ double S[32][33];      // L (lower triangle), written column by column during the factorisation
  __shared__ double col[32];        // column k of L, broadcast to the warp
  __shared__ double dinv[32];       // 1 / L[r][r]
  if (solver_stopped(status)) return;
  const int j = threadIdx.x;
  const unsigned full = 0xffffffffu;
  double a[32], b[32];              // column j of A and of the right-hand side, padded with the identity / zeros
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const bool in = i < R && j < R;
    a[i] = in ? A[i * R + j] : (i == j ? 1.0 : 0.0);
    b[i] = in ? RHS[i * R + j] : 0.0;
  }
  double ajj = 0.0, bjj = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) { ajj = (i == j) ? a[i] : ajj; bjj = (i == j) ? b[i] : bjj; }
  double amax = j < R ? fabs(ajj) : 0.0;                 // the identity padding does not take part in the pivot test
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) amax = fmax(amax, __shfl_xor_sync(full, amax, off));
  int done = 1;
  if (check) done = __all_sync(full, j >= R || bjj <= thr);
  int bad = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    double piv = __shfl_sync(full, a[k], k);             // A[k][k] after the previous updates: lane k, register k
    if (k < R && !(piv > 1e-14 * amax)) { bad = 1; piv = 1.0; }
    const double inv = rsqrt(piv), root = piv * inv;     // one slow fp64 operation per pivot instead of sqrt + divide
    if (j == k) {                                        // lane k publishes column k of L
#pragma unroll
      for (int i = 0; i < 32; ++i) col[i] = i > k ? a[i] * inv : (i == k ? root : 0.0);
      dinv[k] = inv;
    }
    __syncwarp();
    const double ljk = col[j];
    S[j][k] = ljk;
    if (j > k) {                                         // trailing update of column j (rows below the pivot)
#pragma unroll
      for (int i = k + 1; i < 32; ++i) a[i] = fma(-col[i], ljk, a[i]);
    }
    __syncwarp();
  }
  // lane j solves for column j of M with its right-hand side column in registers: forward (L z = b)
  // and backward (L' m = z) substitution, right-looking so that the updates of a step are independent
  // of each other.  The warp barrier after every step keeps ptxas from hoisting hundreds of
  // shared-memory loads ahead (which spills).
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const double x = b[r] * dinv[r];
    b[r] = x;
#pragma unroll
    for (int i = r + 1; i < 32; ++i) b[i] = fma(-S[i][r], x, b[i]);
    __syncwarp();
  }
#pragma unroll
  for (int r = 31; r >= 0; --r) {
    const double x = b[r] * dinv[r];
    b[r] = x;
#pragma unroll
    for (int i = 0; i < r; ++i) b[i] = fma(-S[r][i], x, b[i]);
    __syncwarp();
  }
  if (j < R) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < R) M[i * R + j] = b[i];
  }
  if (j == 0) {
    status[0] = bad;
    if (check) {
      status[1] = done;
      if (!done) status[2] += 1;
    }
  }
}

// status[1] = all diag(G) <= thr  (stopping rule alone)
__global__ void diag_check_kernel(const double* __restrict__ G, int R, double thr, int* __restrict__ status) {
  if (threadIdx.x == 0) {
    int done = 1;
    for (int d = 0; d < R; ++d)
      if (!(G[d * R + d] <= thr)) done = 0;
    status[1] = done;
  }
}

inline int grid_for(long long n) { return (int)std::min<long long>((n + 255) / 256, 148LL * 16); }
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// persistent grid: one CTA per resident slot, never more CTAs than the 8-row blocks need
template <typename K>
int query_occupancy(K kern) {
  int q = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kThreads, 0) != cudaSuccess || q < 1) q = 1;
  return std::min(q, kMaxParts / kSms);
}
inline int persistent_grid(int occ, long long n) {
  const long long need = ((n + 7) / 8 + kWarps - 1) / kWarps;
  return (int)std::max<long long>(1, std::min<long long>((long long)kSms * occ, need));
}

template <int NB, bool V32, bool SYM>
int launch_gram_k(double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st, int* nparts) {
  auto k = gram_partial_kernel<NB, V32, SYM>;
  static const int occ = query_occupancy(k);   // one per instantiation
  const int grid = persistent_grid(occ, n);
  k<<<grid, kThreads, 0, st>>>(dPartial, dXa, dXb, n, R);
  *nparts = grid;
  return FSB_OK;
}

template <int NB, bool V32>
int launch_gram(double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st, int* nparts) {
  return dXa == dXb ? launch_gram_k<NB, V32, true>(dPartial, dXa, dXb, n, R, st, nparts)
                    : launch_gram_k<NB, V32, false>(dPartial, dXa, dXb, n, R, st, nparts);
}

template <int NB, bool V32, int MODE>
int launch_mix(double* dO, const double* dI, const double* dAdd, const double* dM, double* dPartial, long n, int R,
               const int* dStatus, cudaStream_t st, int* nparts) {
  auto k = cg_mix_kernel<NB, V32, MODE>;
  static const int occ = query_occupancy(k);
  const int grid = persistent_grid(occ, n);
  k<<<grid, kThreads, 0, st>>>(dO, dI, dAdd, dM, dPartial, n, R, dStatus);
  if (nparts) *nparts = grid;
  return FSB_OK;
}

template <int MODE>
int dispatch_mix(double* dO, const double* dI, const double* dAdd, const double* dM, double* dPartial, long n, int R,
                 const int* dStatus, cudaStream_t st, int* nparts) {
  if (R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "row mix: R must be 1..32 (got %d)", R);
  if (n <= 0) { if (nparts) *nparts = 0; return FSB_OK; }
  const bool v32 = R == 32 && aligned16(dO) && aligned16(dI) && (MODE != 2 || aligned16(dAdd));
  int rc;
  if (v32) rc = launch_mix<4, true, MODE>(dO, dI, dAdd, dM, dPartial, n, R, dStatus, st, nparts);
  else if (R <= 8) rc = launch_mix<1, false, MODE>(dO, dI, dAdd, dM, dPartial, n, R, dStatus, st, nparts);
  else if (R <= 16) rc = launch_mix<2, false, MODE>(dO, dI, dAdd, dM, dPartial, n, R, dStatus, st, nparts);
  else if (R <= 24) rc = launch_mix<3, false, MODE>(dO, dI, dAdd, dM, dPartial, n, R, dStatus, st, nparts);
  else rc = launch_mix<4, false, MODE>(dO, dI, dAdd, dM, dPartial, n, R, dStatus, st, nparts);
  FSB_TRY(rc);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

}  // namespace

size_t fsb_dense_gram_scratch_bytes(int R) { return (size_t)kMaxParts * R * R * sizeof(double); }

int fsb_dense_gram_partial(double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st, int* nparts) {
  if (R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "gram: R must be 1..32 (got %d)", R);
  if (n <= 0) { *nparts = 0; return FSB_OK; }
  const bool v32 = R == 32 && aligned16(dXa) && aligned16(dXb);
  int rc;
  if (v32) rc = launch_gram<4, true>(dPartial, dXa, dXb, n, R, st, nparts);
  else if (R <= 8) rc = launch_gram<1, false>(dPartial, dXa, dXb, n, R, st, nparts);
  else if (R <= 16) rc = launch_gram<2, false>(dPartial, dXa, dXb, n, R, st, nparts);
  else if (R <= 24) rc = launch_gram<3, false>(dPartial, dXa, dXb, n, R, st, nparts);
  else rc = launch_gram<4, false>(dPartial, dXa, dXb, n, R, st, nparts);
  FSB_TRY(rc);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_gram_finalize(double* dG, const double* dPartial, int nparts, int R, cudaStream_t st) {
  gram_final_kernel<<<final_grid(R * R), kFinalThreads, 0, st>>>(dG, dPartial, nparts, R * R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_gram_into(double* dG, double* dPartial, const double* dXa, const double* dXb, long n, int R, cudaStream_t st) {
  int nparts = 0;
  FSB_TRY(fsb_dense_gram_partial(dPartial, dXa, dXb, n, R, st, &nparts));
  return fsb_dense_gram_finalize(dG, dPartial, nparts, R, st);
}

int fsb_dense_axpy_lambda(double* dY, const double* dX, double lambda, long n, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  axpy_lambda_kernel<<<grid_for(n), 256, 0, st>>>(dY, dX, lambda, n);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_cg_norms(double* dNorm, double* dInorm, const double* dG, int R, int normalise, cudaStream_t st) {
  cg_norms_kernel<<<1, 32, 0, st>>>(dNorm, dInorm, dG, R, normalise);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_cg_init(double* dX, double* dRm, double* dP, const double* dB, const double* dInorm, long n, int R, cudaStream_t st) {
  const long long tot = (long long)n * R;
  if (tot <= 0) return FSB_OK;
  cg_init_kernel<<<grid_for(tot), 256, 0, st>>>(dX, dRm, dP, dB, dInorm, tot, R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_mix_add(double* dO, const double* dI, const double* dM, long n, int R, const int* dStatus, cudaStream_t st) {
  return dispatch_mix<0>(dO, dI, nullptr, dM, nullptr, n, R, dStatus, st, nullptr);
}

int fsb_dense_mix_sub_gram(double* dO, const double* dI, const double* dM, double* dPartial, long n, int R, const int* dStatus,
                           cudaStream_t st, int* nparts) {
  return dispatch_mix<1>(dO, dI, nullptr, dM, dPartial, n, R, dStatus, st, nparts);
}

int fsb_dense_mix_set(double* dO, const double* dI, const double* dAdd, const double* dM, long n, int R, const int* dStatus, cudaStream_t st) {
  return dispatch_mix<2>(dO, dI, dAdd, dM, nullptr, n, R, dStatus, st, nullptr);
}

int fsb_dense_split_halves(double* dLo, double* dHi, const double* dSrc, long n, int R, const int* dStatus, cudaStream_t st) {
  if (R % 4 || !aligned16(dLo) || !aligned16(dHi) || !aligned16(dSrc)) return fsb_set_error(FSB_EINVAL, "split halves: R must be a multiple of 4, operands 16-byte aligned");
  if (n <= 0) return FSB_OK;
  split_halves_kernel<<<grid_for((long long)n * R / 2), 256, 0, st>>>(reinterpret_cast<double2*>(dLo), reinterpret_cast<double2*>(dHi),
                                                                   reinterpret_cast<const double2*>(dSrc), n, R / 4, dStatus);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_scale_cols(double* dX, const double* dNorm, long n, int R, cudaStream_t st) {
  const long long tot = (long long)n * R;
  if (tot <= 0) return FSB_OK;
  cg_scale_cols_kernel<<<grid_for(tot), 256, 0, st>>>(dX, dNorm, tot, R);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_small_solve(double* dM, double* dA, double* dRHS, const double* dPartA, int nA, const double* dPartRHS, int nRHS, int R,
                          int* dStatus, int check, double thr, cudaStream_t st) {
  // second stage of the Gram reductions first (their own small parallel kernel), then the one-warp solve
  if (nA > 0) FSB_TRY(fsb_dense_gram_finalize(dA, dPartA, nA, R, st));
  if (nRHS > 0) FSB_TRY(fsb_dense_gram_finalize(dRHS, dPartRHS, nRHS, R, st));
  small_solve_kernel<<<1, 32, 0, st>>>(dM, dA, dRHS, R, dStatus, check, thr);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

int fsb_dense_diag_check(const double* dG, int R, double thr, int* dStatus, cudaStream_t st) {
  diag_check_kernel<<<1, 32, 0, st>>>(dG, R, thr, dStatus);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

extern "C" int fsb_gram_dev(double* G_host, const double* dXa, const double* dXb, long n, int R, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!G_host || !dXa || !dXb || n < 0) return fsb_set_error(FSB_EINVAL, "fsb_gram_dev: bad argument");
  cudaStream_t st = fsb_pick_stream(stream);
  double *dG = nullptr, *dPart = nullptr;
  FSB_CUDA(cudaMalloc(&dG, (size_t)R * R * 8));
  cudaError_t e = cudaMalloc(&dPart, fsb_dense_gram_scratch_bytes(R));
  int rc = e == cudaSuccess ? fsb_dense_gram_into(dG, dPart, dXa, dXb, n, R, st) : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  if (rc == FSB_OK) {
    e = cudaMemcpyAsync(G_host, dG, (size_t)R * R * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "gram readback", __FILE__, __LINE__);
  }
  cudaFree(dG); cudaFree(dPart);
  return rc;
}

extern "C" int fsb_rowmix_dev(int mode, double* dO, const double* dI, const double* dAdd, const double* dM, double* G_host,
                              long n, int R, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!dO || !dI || !dM || n < 0 || mode < 0 || mode > 2 || (mode == 2 && !dAdd)) return fsb_set_error(FSB_EINVAL, "fsb_rowmix_dev: bad argument");
  if (R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "fsb_rowmix_dev: R must be 1..32 (got %d)", R);
  cudaStream_t st = fsb_pick_stream(stream);
  if (mode == 0) return fsb_dense_mix_add(dO, dI, dM, n, R, nullptr, st);
  if (mode == 2) return fsb_dense_mix_set(dO, dI, dAdd, dM, n, R, nullptr, st);
  double *dG = nullptr, *dPart = nullptr;
  FSB_CUDA(cudaMalloc(&dG, (size_t)R * R * 8));
  cudaError_t e = cudaMalloc(&dPart, fsb_dense_gram_scratch_bytes(R));
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  int np = 0;
  if (rc == FSB_OK) rc = fsb_dense_mix_sub_gram(dO, dI, dM, dPart, n, R, nullptr, st, &np);
  if (rc == FSB_OK && G_host) {
    gram_final_kernel<<<final_grid(R * R), kFinalThreads, 0, st>>>(dG, dPart, np, R * R);
    fsb_count_launch();
    e = cudaMemcpyAsync(G_host, dG, (size_t)R * R * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "rowmix readback", __FILE__, __LINE__);
  }
  cudaFree(dG); cudaFree(dPart);
  return rc;
}

namespace {
__global__ void diff_kernel(double* d, const double* x, const double* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = x[i] - y[i];
}
}  // namespace

extern "C" int fsb_gram_host(double* G, const double* Xa, const double* Xb, long n, int R) {
  FSB_TRY(fsb_require_device());
  if (!G || !Xa || !Xb || n < 0 || R < 1 || R > 32) return fsb_set_error(FSB_EINVAL, "fsb_gram_host: bad argument");
  cudaStream_t st = fsb_default_stream();
  const size_t bytes = std::max<size_t>((size_t)n * R, 1) * 8;
  double *da = nullptr, *db = nullptr;
  FSB_CUDA(cudaMalloc(&da, bytes));
  int rc = FSB_OK;
  cudaError_t e = cudaMemcpyAsync(da, Xa, (size_t)n * R * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && Xb != Xa) {
    e = cudaMalloc(&db, bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db, Xb, (size_t)n * R * 8, cudaMemcpyHostToDevice, st);
  }
  if (e != cudaSuccess) rc = fsb_cuda_error(e, "fsb_gram_host staging", __FILE__, __LINE__);
  if (rc == FSB_OK) rc = fsb_gram_dev(G, da, db ? db : da, n, R, (void*)st);
  cudaFree(da); cudaFree(db);
  return rc;
}

extern "C" int fsb_dist_host(double* out, const double* x, const double* y, long n) {
  FSB_TRY(fsb_require_device());
  if (!out || !x || !y || n < 0) return fsb_set_error(FSB_EINVAL, "fsb_dist_host: bad argument");
  cudaStream_t st = fsb_default_stream();
  const size_t bytes = std::max<size_t>((size_t)n, 1) * 8;
  double *dx = nullptr, *dy = nullptr;
  FSB_CUDA(cudaMalloc(&dx, bytes));
  cudaError_t e = cudaMalloc(&dy, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dy, y, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "fsb_dist_host staging", __FILE__, __LINE__);
  if (rc == FSB_OK && n > 0) {
    diff_kernel<<<grid_for(n), 256, 0, st>>>(dx, dx, dy, n);
    fsb_count_launch();
  }
  double g = 0.0;
  if (rc == FSB_OK) rc = fsb_gram_dev(&g, dx, dx, n, 1, (void*)st);
  cudaFree(dx); cudaFree(dy);
  if (rc == FSB_OK) *out = sqrt(g);
  return rc;
}

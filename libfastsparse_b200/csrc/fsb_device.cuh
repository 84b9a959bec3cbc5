// fsb_device.cuh -- device-side load/store helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>

namespace fsbdev {

// streaming int load: matrix indices are read exactly once per product
__device__ __forceinline__ int ld_stream_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream_s32x4(const int* p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// L2 cache policy word for ld/st ...L2::cache_hint: 0 = normal, 1 = evict_last (keep the dense
// operand resident), 2 = evict_first (streams that are touched once)
__device__ __forceinline__ unsigned long long make_l2_policy(int kind) {
  unsigned long long p;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  // experiment (knob cap_mult 3xx..6xx of fsb_tune_csr_algo): keep only a FRACTION of the operand's lines (chosen by
  // address hash) and let the rest stream, for operands about twice the L2 a gather operand gets.  Measured at C2: every
  // fraction is slower than evict_last 1.0 (4.74-4.96 ms): 0.75 5.05-5.22, 0.5 5.22-5.43, 0.375 5.38-5.49 ms
  // (profiles/r2z_l2_fraction_probe.jsonl)
  else if (kind == 3) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.5;" : "=l"(p));
  else if (kind == 4) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.75;" : "=l"(p));
  else if (kind == 5) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.375;" : "=l"(p));
  else if (kind == 6) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.5;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int ld_stream_s32_pol(const int* p, unsigned long long pol) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double ld_stream_f64_pol(const double* p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}

// gather of VEC consecutive doubles of a dense operand row (read-only path).
// VEC = 4 is one 256-bit LDG (sm_100+); the L2::evict_last form keeps the dense
// operand resident in L2 against the streaming matrix/Y traffic.
template <int VEC> struct XLoad;
template <> struct XLoad<1> {
  static __device__ __forceinline__ void ld(double* v, const double* p) {
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v[0]) : "l"(p));
  }
  static __device__ __forceinline__ void ldp(double* v, const double* p, unsigned long long pol) {
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v[0]) : "l"(p), "l"(pol));
  }
};
template <> struct XLoad<2> {
  static __device__ __forceinline__ void ld(double* v, const double* p) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
  }
  static __device__ __forceinline__ void ldp(double* v, const double* p, unsigned long long pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[0]), "=d"(v[1]) : "l"(p), "l"(pol));
  }
};
#ifndef FSB_X_EVICT_LAST
#define FSB_X_EVICT_LAST 1
#endif
template <> struct XLoad<4> {
  static __device__ __forceinline__ void ld(double* v, const double* p) {
#if FSB_X_EVICT_LAST
    asm volatile("ld.global.nc.L2::evict_last.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#else
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#endif
  }
  static __device__ __forceinline__ void ldp(double* v, const double* p, unsigned long long pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p), "l"(pol));
  }
};

// streaming (evict-first) store of VEC doubles: outputs are written once
template <int VEC> struct YStore;
template <> struct YStore<1> {
  static __device__ __forceinline__ void st(double* p, const double* v) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v[0]) : "memory");
  }
};
template <> struct YStore<2> {
  static __device__ __forceinline__ void st(double* p, const double* v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
  }
};
template <> struct YStore<4> {
  static __device__ __forceinline__ void st(double* p, const double* v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p + 2), "d"(v[2]), "d"(v[3]) : "memory");
  }
};

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous run global -> shared, completion on an mbarrier.
// src / dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar,
                                            unsigned long long l2_policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "l"(l2_policy) : "memory");
}

// fire-and-forget fp64 add at L2 (REDG.E.ADD.F64)
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double shfl_f64(unsigned mask, double v, int src, int width) {
  return __shfl_sync(mask, v, src, width);
}
__device__ __forceinline__ double shfl_xor_f64(unsigned mask, double v, int off, int width) {
  return __shfl_xor_sync(mask, v, off, width);
}

}  // namespace fsbdev

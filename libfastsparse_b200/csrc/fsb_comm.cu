// fsb_comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-process OpenMP and has no communication layer at all
// (SURVEY 2a).  The path shards by rows: A x needs no exchange; the
// per-shard partial of A'(...) and the CG Gram matrices are sum-allreduced.  NCCL is
// bound with dlopen so that (a) the library loads on machines without NCCL and (b)
// inside a torch process the already-loaded (torch-bundled) libnccl is reused instead
// of mixing two copies.
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "fsb_internal.h"

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_reduce_scatter)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
typedef int (*fn_group)(void);
typedef int (*fn_destroy)(nccl_comm);
typedef const char* (*fn_errstr)(int);

struct Nccl {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_reduce_scatter reduce_scatter = nullptr;
  fn_allgather allgather = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

nccl_comm g_comm = nullptr;
int g_nranks = 1, g_rank = 0;
std::mutex g_mu;

int load_nccl() {
  if (g_nccl.lib) return FSB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) return fsb_set_error(FSB_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.reduce_scatter = (fn_reduce_scatter)dlsym(g_nccl.lib, "ncclReduceScatter");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.group_start = (fn_group)dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.group_end = (fn_group)dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.reduce_scatter || !g_nccl.allgather || !g_nccl.destroy) {
    g_nccl.lib = nullptr;
    return fsb_set_error(FSB_ENCCL, "libnccl is missing a required symbol");
  }
  return FSB_OK;
}

int nccl_fail(int code, const char* what) {
  return fsb_set_error(FSB_ENCCL, "NCCL error %d (%s) in %s", code, g_nccl.errstr ? g_nccl.errstr(code) : "?", what);
}

}  // namespace

bool fsb_comm_active() { return g_comm != nullptr && g_nranks > 1; }

// recv[recvcount] = this rank's slice of the element-wise sum of every rank's send[nranks*recvcount]
int fsb_comm_reduce_scatter_sum(const double* send, double* recv, size_t recvcount, cudaStream_t st) {
  if (!fsb_comm_active()) return fsb_set_error(FSB_ENCCL, "reduce-scatter without an active communicator");
  const int rc = g_nccl.reduce_scatter(send, recv, recvcount, /*ncclFloat64*/ 8, /*ncclSum*/ 0, g_comm, st);
  return rc ? nccl_fail(rc, "ncclReduceScatter") : FSB_OK;
}

// several collectives issued between these two calls are fused by NCCL into one launch
int fsb_comm_group_start() {
  if (!fsb_comm_active() || !g_nccl.group_start) return FSB_OK;
  const int rc = g_nccl.group_start();
  return rc ? nccl_fail(rc, "ncclGroupStart") : FSB_OK;
}
int fsb_comm_group_end() {
  if (!fsb_comm_active() || !g_nccl.group_end) return FSB_OK;
  const int rc = g_nccl.group_end();
  return rc ? nccl_fail(rc, "ncclGroupEnd") : FSB_OK;
}

// recv[nranks*sendcount] = concatenation of every rank's send[sendcount] in rank order
int fsb_comm_allgather(const double* send, double* recv, size_t sendcount, cudaStream_t st) {
  if (!fsb_comm_active()) return fsb_set_error(FSB_ENCCL, "all-gather without an active communicator");
  const int rc = g_nccl.allgather(send, recv, sendcount, /*ncclFloat64*/ 8, g_comm, st);
  return rc ? nccl_fail(rc, "ncclAllGather") : FSB_OK;
}

extern "C" {

int fsb_comm_unique_id(void* id_out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!id_out) return fsb_set_error(FSB_EINVAL, "fsb_comm_unique_id: null argument");
  FSB_TRY(load_nccl());
  nccl_uid id;
  const int rc = g_nccl.get_uid(&id);
  if (rc) return nccl_fail(rc, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof id);
  return FSB_OK;
}

int fsb_comm_init(int nranks, int rank, const void* id) {
  FSB_TRY(fsb_require_device());
  std::lock_guard<std::mutex> lk(g_mu);
  if (nranks < 1 || rank < 0 || rank >= nranks || !id) return fsb_set_error(FSB_EINVAL, "fsb_comm_init: bad arguments");
  if (g_comm) return fsb_set_error(FSB_EINVAL, "fsb_comm_init: communicator already active");
  FSB_TRY(load_nccl());
  nccl_uid uid;
  memcpy(&uid, id, sizeof uid);
  const int rc = g_nccl.init_rank(&g_comm, nranks, uid, rank);
  if (rc) { g_comm = nullptr; return nccl_fail(rc, "ncclCommInitRank"); }
  g_nranks = nranks;
  g_rank = rank;
  return FSB_OK;
}

int fsb_comm_finalize(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) {
    cudaDeviceSynchronize();
    g_nccl.destroy(g_comm);
  }
  g_comm = nullptr;
  g_nranks = 1;
  g_rank = 0;
  return FSB_OK;
}

int fsb_comm_size(void) { return g_nranks; }
int fsb_comm_rank(void) { return g_rank; }

int fsb_allreduce_sum_dev(double* dBuf, long count, void* stream) {
  FSB_TRY(fsb_require_device());
  if (!fsb_comm_active() || count <= 0) return FSB_OK;
  const int rc = g_nccl.allreduce(dBuf, dBuf, (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, g_comm, fsb_pick_stream(stream));
  if (rc) return nccl_fail(rc, "ncclAllReduce");
  return FSB_OK;
}

int fsb_partition_rows(int nrow, const int* row_ptr, int nparts, int* bounds) {
  if (nrow < 0 || nparts < 1 || !row_ptr || !bounds) return fsb_set_error(FSB_EINVAL, "fsb_partition_rows: bad arguments");
  const long nnz = row_ptr[nrow];
  bounds[0] = 0;
  int r = 0;
  for (int p = 1; p < nparts; ++p) {
    // first row whose prefix reaches p/nparts of the entries (ties keep rows balanced for empty matrices)
    const long target = nnz > 0 ? (long)(((__int128)nnz * p) / nparts) : 0;
    if (nnz == 0) { r = (int)(((long)nrow * p) / nparts); }
    else { while (r < nrow && row_ptr[r] < target) ++r; }
    bounds[p] = r < bounds[p - 1] ? bounds[p - 1] : r;
  }
  bounds[nparts] = nrow;
  return FSB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Peer-memory collectives over NVLink / NVSwitch (one process per GPU on one node).
//
// A symmetric buffer is cudaMalloc'ed by every rank and mapped into every other rank's address space through CUDA IPC
// (the unique ids of the handles travel through one NCCL all-gather).  On top of it:
//   fsb_p2p_allgather_chunks -- the all-gather of the block-CG search directions P: every rank STORES its slices
//     straight into the replicated buffer of all G ranks (peer stores over NVLink, 16 bytes per thread), then publishes
//     an epoch flag to each peer (release at system scope after the last CTA's stores); a one-thread wait kernel on the
//     consumer's stream acquires the G flags before the product that reads P starts.  No staging copies, no
//     intermediate protocol buffers: the bytes cross the switch once, at the rate the SMs can push them.
// Safe reuse without double buffering: a rank can only reach its next push after the reduce-scatter of the current
// iteration, which needs every rank's A'(A P) partial, i.e. every rank has finished reading the previous P.
// The wait kernel gives up after ~15 s (a peer died) and raises a device-side error flag instead of hanging the GPU.
namespace {

constexpr int kMaxPeers = 8;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct PushArgs {
  double2* dst[kMaxPeers];                  // replicated buffer of every rank (this rank's own included)
  unsigned long long* flags[kMaxPeers];     // flag array of every rank: flags[g][src_rank]
};

// src: this rank's local slices [C][slice2] (double2 units); destination offset of slice (c, rank): c*chunk2 + rank*slice2
// The epoch lives on the device (epoch_ptr, this rank's): the kernel publishes *epoch_ptr + 1 and stores it back, so the
// same launch can be replayed from a CUDA graph (the block-CG iteration is captured once and replayed).
__global__ void __launch_bounds__(256) p2p_push_kernel(const double2* __restrict__ src, PushArgs a, long long slice2, long long chunk2, int C,
                                                       int G, int rank, unsigned long long* __restrict__ epoch_ptr, unsigned int* __restrict__ counter) {
  const unsigned long long epoch = *epoch_ptr + 1;     // every CTA reads it before any CTA can be "last"
  const int g = blockIdx.y;
  double2* __restrict__ dst = a.dst[g];
  const long long n = (long long)C * slice2;
  // one 16-byte load and one peer store per thread and step: with ~600 CTAs in flight the NVLink store path, not the
  // load latency, is the limit (a four-way unrolled variant measured 0.214 vs 0.220 ms on 2 GPUs but 0.38 vs 0.35 ms on 8)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / slice2, k = i - c * slice2;
    dst[c * chunk2 + (long long)rank * slice2 + k] = src[i];
  }
  // the last CTA to finish publishes the epoch: every CTA's stores -> system fence -> counter; last CTA -> fence -> flags
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    last = atomicAdd(counter, 1u) == total - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < G) st_release_sys(a.flags[threadIdx.x] + rank, epoch);
    if (threadIdx.x == 0) { *counter = 0; *epoch_ptr = epoch; }
  }
}

// one warp: lane g waits for rank g's flag; err[0] is set when a peer never shows up
__global__ void p2p_wait_kernel(const unsigned long long* __restrict__ flags, int G, const unsigned long long* __restrict__ epoch_ptr,
                                int* __restrict__ err) {
  const int g = threadIdx.x;
  if (g >= G) return;
  const unsigned long long epoch = *epoch_ptr;         // already advanced by this rank's push kernel (stream order)
  const long long t0 = clock64();
  while (ld_acquire_sys(flags + g) < epoch) {
    if (clock64() - t0 > 30000000000LL) { atomicExch(err, 1); break; }    // ~15 s at 2 GHz
    __nanosleep(64);
  }
}

// Small sum-allreduce (the R x R Gram matrices of the block CG, <= kSmallMax doubles) in ONE single-CTA kernel: store
// my vector into slot [parity][rank] of every rank, publish the epoch, wait for the G epochs, add the G slots in rank
// order (identical bits on every rank -- the solver's ranks take identical branches without exchanging flags).
// Two slot sets alternate by epoch parity: a rank can run at most one allreduce ahead of the slowest reader.
constexpr int kSmallMax = 1024;
struct SmallArgs {
  double* slots[kMaxPeers];                 // every rank's slot area: [2][kMaxPeers][kSmallMax]
  unsigned long long* flags[kMaxPeers];     // every rank's flag array for this channel: flags[g][src_rank]
};
__global__ void __launch_bounds__(256) p2p_small_allreduce_kernel(double* __restrict__ buf, int n, SmallArgs a, int G, int rank,
                                                                  unsigned long long* __restrict__ epoch_ptr, int* __restrict__ err) {
  const unsigned long long epoch = *epoch_ptr + 1;
  const int par = (int)(epoch & 1);
  for (int g = 0; g < G; ++g) {
    double* dst = a.slots[g] + ((size_t)par * kMaxPeers + rank) * kSmallMax;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < G) {
    st_release_sys(a.flags[threadIdx.x] + rank, epoch);
    const long long t0 = clock64();
    while (ld_acquire_sys(a.flags[rank] + threadIdx.x) < epoch) {
      if (clock64() - t0 > 30000000000LL) { atomicExch(err, 1); break; }
      __nanosleep(32);
    }
  }
  __syncthreads();
  const double* mine = a.slots[rank] + (size_t)par * kMaxPeers * kSmallMax;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = mine[i];
    for (int g = 1; g < G; ++g) s += mine[(size_t)g * kSmallMax + i];
    buf[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ptr = epoch;
}

// Reduce-scatter by PULL: after a rank has produced chunk c of its full-length partial it signals the chunk's epoch to
// every peer; the owner of a slice then reads the G partial slices straight out of the peers' buffers (NVLink loads)
// and adds them in rank order, plus lambda * add[] (the "+ lambda P" of the CG operator).
__global__ void p2p_signal_kernel(SmallArgs a, int G, int rank, int word0, unsigned long long* __restrict__ epoch_ptr) {
  const unsigned long long epoch = *epoch_ptr + 1;
  __syncwarp();
  if (threadIdx.x < G) st_release_sys(a.flags[threadIdx.x] + word0 + rank, epoch);
  __syncwarp();
  if (threadIdx.x == 0) *epoch_ptr = epoch;
}

struct PullArgs { const double2* src[kMaxPeers]; };
__global__ void __launch_bounds__(256) p2p_pull_sum_kernel(double2* __restrict__ out, PullArgs a, long long n2, int G,
                                                           const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ epoch_ptr,
                                                           const double2* __restrict__ add, double lambda, int* __restrict__ err) {
  if (threadIdx.x < G) {
    const unsigned long long epoch = *epoch_ptr;       // this rank's own signal count for the chunk (already advanced)
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > 30000000000LL) { atomicExch(err, 1); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  // NVLink loads have a few microseconds of latency: keep kUnroll x G independent 16-byte loads in flight per thread
  constexpr int kUnroll = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n2; i0 += stride * kUnroll) {
    double2 v[kUnroll][kMaxPeers];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
#pragma unroll
      for (int g = 0; g < kMaxPeers; ++g)
        if (g < G && i < n2) v[u][g] = a.src[g][i];
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i >= n2) break;
      double2 s = v[u][0];
#pragma unroll
      for (int g = 1; g < kMaxPeers; ++g)
        if (g < G) { s.x += v[u][g].x; s.y += v[u][g].y; }
      if (add) { const double2 p = add[i]; s.x = fma(lambda, p.x, s.x); s.y = fma(lambda, p.y, s.y); }
      out[i] = s;
    }
  }
}

}  // namespace

struct fsb_p2p {
  int G = 1, rank = 0;
  size_t bytes = 0;
  void* local = nullptr;
  void* peer[kMaxPeers] = {};               // peer[g]: rank g's buffer mapped here (peer[rank] == local)
  unsigned long long* flags = nullptr;      // [kMaxPeers] at the head of this rank's flag page
  unsigned long long* peer_flags[kMaxPeers] = {};
  unsigned int* counter = nullptr;
  int* err = nullptr;
  unsigned long long* epoch_dev = nullptr;  // epochs published so far (device-resident: graph-replayable)
  // small-allreduce channel
  double* slots = nullptr;
  double* peer_slots[kMaxPeers] = {};
  unsigned long long* epoch2_dev = nullptr;
  unsigned long long* epoch3_dev = nullptr;   // [8]: per-channel signal counts (reduce-scatter by pull)
};

namespace {
// every rank contributes one IPC handle; out[g] = rank g's allocation mapped into this process.  COLLECTIVE: a rank
// whose own allocation failed still takes part (mine == nullptr is announced as "no handle"), so nobody hangs.
int exchange_ipc(void* mine, void* out[kMaxPeers], cudaStream_t st) {
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof h);
  int ok = (mine && cudaIpcGetMemHandle(&h, mine) == cudaSuccess) ? 1 : 0;
  if (!ok) cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  constexpr int kWords = 16;                // 64 bytes of handle + status, in doubles
  double hbuf[kWords] = {};
  memcpy(hbuf, &h, 64);
  hbuf[8] = ok;
  double *dsend = nullptr, *drecv = nullptr;
  FSB_CUDA(cudaMalloc(&dsend, kWords * 8));
  cudaError_t e = cudaMalloc(&drecv, (size_t)kWords * 8 * g_nranks);
  int rc = e == cudaSuccess ? FSB_OK : fsb_cuda_error(e, "cudaMalloc", __FILE__, __LINE__);
  double all[kWords * kMaxPeers] = {};
  if (rc == FSB_OK) {
    e = cudaMemcpyAsync(dsend, hbuf, sizeof hbuf, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) rc = fsb_comm_allgather(dsend, drecv, kWords, st);
    if (e == cudaSuccess && rc == FSB_OK) e = cudaMemcpyAsync(all, drecv, (size_t)kWords * 8 * g_nranks, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == FSB_OK) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fsb_cuda_error(e, "IPC handle exchange", __FILE__, __LINE__);
  }
  cudaFree(dsend); cudaFree(drecv);
  FSB_TRY(rc);
  for (int g = 0; g < g_nranks; ++g)
    if (all[g * kWords + 8] != 1.0) return fsb_set_error(FSB_ENCCL, "peer memory: rank %d could not export an IPC handle", g);
  for (int g = 0; g < g_nranks; ++g) {
    if (g == g_rank) { out[g] = mine; continue; }
    cudaIpcMemHandle_t ph;
    memcpy(&ph, &all[g * kWords], 64);
    e = cudaIpcOpenMemHandle(&out[g], ph, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fsb_set_error(FSB_ENCCL, "peer memory: cannot map rank %d's buffer (%s)", g, cudaGetErrorString(e));
    }
  }
  return FSB_OK;
}
}  // namespace

// Collective (every rank, same size, same order).  On failure on ANY rank every rank returns an error (the status
// is agreed through an allreduce) and the caller keeps using NCCL.
int fsb_p2p_create(fsb_p2p** out, size_t bytes, cudaStream_t st) {
  *out = nullptr;
  if (!fsb_comm_active() || g_nranks > kMaxPeers) return fsb_set_error(FSB_ENCCL, "peer memory: needs an active communicator of at most %d ranks", kMaxPeers);
  fsb_p2p* p = new fsb_p2p();
  p->G = g_nranks; p->rank = g_rank; p->bytes = bytes;
  int rc = FSB_OK;
  void* flagpage = nullptr;
  cudaError_t e = cudaMalloc(&p->local, std::max<size_t>(bytes, 256));
  if (e == cudaSuccess) e = cudaMalloc(&flagpage, 4096);
  if (e == cudaSuccess) e = cudaMemsetAsync(flagpage, 0, 4096, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { cudaGetLastError(); rc = FSB_ECUDA; }
  void* pf[kMaxPeers] = {};
  void* ps[kMaxPeers] = {};
  void* slots = nullptr;
  const size_t slot_bytes = (size_t)2 * kMaxPeers * kSmallMax * sizeof(double);
  if (rc == FSB_OK && cudaMalloc(&slots, slot_bytes) != cudaSuccess) { cudaGetLastError(); rc = FSB_ECUDA; }
  // the three exchanges always run, whatever happened locally: they are collective
  const int rc1 = exchange_ipc(rc == FSB_OK ? p->local : nullptr, p->peer, st);
  const int rc2 = exchange_ipc(rc == FSB_OK ? flagpage : nullptr, pf, st);
  const int rc3 = exchange_ipc(rc == FSB_OK ? slots : nullptr, ps, st);
  if (rc == FSB_OK) rc = rc1 != FSB_OK ? rc1 : (rc2 != FSB_OK ? rc2 : rc3);
  // agree on the outcome
  double flag = rc == FSB_OK ? 0.0 : 1.0, *dflag = nullptr;
  if (cudaMalloc(&dflag, 8) == cudaSuccess) {
    cudaMemcpyAsync(dflag, &flag, 8, cudaMemcpyHostToDevice, st);
    fsb_allreduce_sum_dev(dflag, 1, (void*)st);
    cudaMemcpyAsync(&flag, dflag, 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(dflag);
  }
  if (flag != 0.0 || rc != FSB_OK) {
    for (int g = 0; g < p->G; ++g) {
      if (g != p->rank && p->peer[g]) cudaIpcCloseMemHandle(p->peer[g]);
      if (g != p->rank && pf[g]) cudaIpcCloseMemHandle(pf[g]);
      if (g != p->rank && ps[g]) cudaIpcCloseMemHandle(ps[g]);
    }
    cudaFree(p->local); cudaFree(flagpage); cudaFree(slots);
    delete p;
    return rc != FSB_OK ? rc : fsb_set_error(FSB_ENCCL, "peer memory: another rank could not map the buffers");
  }
  p->flags = (unsigned long long*)flagpage;
  for (int g = 0; g < p->G; ++g) p->peer_flags[g] = (unsigned long long*)pf[g];
  p->counter = (unsigned int*)((char*)flagpage + 2048);
  p->err = (int*)((char*)flagpage + 2048 + 64);
  p->epoch_dev = (unsigned long long*)((char*)flagpage + 2048 + 128);
  p->epoch2_dev = (unsigned long long*)((char*)flagpage + 2048 + 192);
  p->epoch3_dev = (unsigned long long*)((char*)flagpage + 2048 + 256);
  p->slots = (double*)slots;
  for (int g = 0; g < p->G; ++g) p->peer_slots[g] = (double*)ps[g];
  *out = p;
  return FSB_OK;
}

void fsb_p2p_destroy(fsb_p2p* p) {
  if (!p) return;
  cudaDeviceSynchronize();
  for (int g = 0; g < p->G; ++g) {
    if (g == p->rank) continue;
    if (p->peer[g]) cudaIpcCloseMemHandle(p->peer[g]);
    if (p->peer_flags[g]) cudaIpcCloseMemHandle(p->peer_flags[g]);
    if (p->peer_slots[g]) cudaIpcCloseMemHandle(p->peer_slots[g]);
  }
  cudaFree(p->local);
  cudaFree(p->flags);
  cudaFree(p->slots);
  delete p;
}

void* fsb_p2p_local(fsb_p2p* p) { return p->local; }

// all-gather of C chunks: rank r's slice of chunk c (slice doubles at loc + c*slice) lands at offset c*G*slice + r*slice of
// EVERY rank's buffer.  Returns after enqueueing push + wait on st; the data is complete for kernels that follow on st.
int fsb_p2p_allgather_chunks(fsb_p2p* p, const double* loc, int C, long slice, cudaStream_t st) {
  if (slice % 2 || ((uintptr_t)loc & 15)) return fsb_set_error(FSB_EINVAL, "peer all-gather: slices must be 16-byte aligned");
  if ((size_t)C * p->G * slice * 8 > p->bytes) return fsb_set_error(FSB_EINVAL, "peer all-gather: buffer too small");
  PushArgs a;
  for (int g = 0; g < kMaxPeers; ++g) { a.dst[g] = g < p->G ? (double2*)p->peer[g] : nullptr; a.flags[g] = g < p->G ? p->peer_flags[g] : nullptr; }
  const long long n2 = (long long)C * slice / 2;
  const int bx = (int)std::max<long long>(1, std::min<long long>((n2 + 255) / 256, (148 * 4) / p->G + 1));
  dim3 grid(bx, p->G);
  p2p_push_kernel<<<grid, 256, 0, st>>>((const double2*)loc, a, slice / 2, (long long)p->G * slice / 2, C, p->G, p->rank, p->epoch_dev, p->counter);
  FSB_KERNEL_CHECK();
  p2p_wait_kernel<<<1, 32, 0, st>>>(p->flags, p->G, p->epoch_dev, p->err);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

// in-place sum-allreduce of n <= 1024 doubles over the peer-mapped slots (one kernel, ~10 us); identical bits on every rank
int fsb_p2p_allreduce_small(fsb_p2p* p, double* buf, int n, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  if (n > kSmallMax) return fsb_set_error(FSB_EINVAL, "peer allreduce: at most %d doubles", kSmallMax);
  SmallArgs a;
  for (int g = 0; g < kMaxPeers; ++g) {
    a.slots[g] = g < p->G ? p->peer_slots[g] : nullptr;
    a.flags[g] = g < p->G ? p->peer_flags[g] + 64 : nullptr;     // second flag channel: words 64.. of the flag page
  }
  p2p_small_allreduce_kernel<<<1, 256, 0, st>>>(buf, n, a, p->G, p->rank, p->epoch2_dev, p->err);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

// "chunk `channel` of my buffer is complete": publish the channel's next epoch to every rank (channel < 8)
int fsb_p2p_signal(fsb_p2p* p, int channel, cudaStream_t st) {
  if (channel < 0 || channel >= 8) return fsb_set_error(FSB_EINVAL, "peer signal: channel out of range");
  SmallArgs a;
  for (int g = 0; g < kMaxPeers; ++g) { a.slots[g] = nullptr; a.flags[g] = g < p->G ? p->peer_flags[g] : nullptr; }
  p2p_signal_kernel<<<1, 32, 0, st>>>(a, p->G, p->rank, 128 + channel * 8, p->epoch3_dev + channel);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

// out[0..n) = sum over ranks g (in rank order) of rank g's buffer[elem_offset .. +n)  (+ lambda * add[0..n) when add != nullptr),
// after every rank has signalled `channel` as often as this rank has.  n even, offsets 16-byte aligned.
int fsb_p2p_pull_sum(fsb_p2p* p, int channel, double* out, size_t elem_offset, long n, const double* add, double lambda, cudaStream_t st) {
  if (n <= 0) return FSB_OK;
  if (n % 2 || elem_offset % 2 || ((uintptr_t)out & 15) || (add && ((uintptr_t)add & 15)))
    return fsb_set_error(FSB_EINVAL, "peer reduce-scatter: operands must be 16-byte aligned");
  if ((elem_offset + (size_t)n) * 8 > p->bytes) return fsb_set_error(FSB_EINVAL, "peer reduce-scatter: range outside the buffer");
  PullArgs a;
  for (int g = 0; g < kMaxPeers; ++g) a.src[g] = g < p->G ? (const double2*)((const double*)p->peer[g] + elem_offset) : nullptr;
  const long long n2 = n / 2;
  const int grid = (int)std::max<long long>(1, std::min<long long>((n2 + 255) / 256, 148));   // 160 registers: one CTA per SM
  p2p_pull_sum_kernel<<<grid, 256, 0, st>>>((double2*)out, a, n2, p->G, p->flags + 128 + channel * 8, p->epoch3_dev + channel, (const double2*)add,
                                             lambda, p->err);
  FSB_KERNEL_CHECK();
  return FSB_OK;
}

// non-zero when a wait kernel timed out (a peer never published its epoch)
int fsb_p2p_check(fsb_p2p* p, cudaStream_t st) {
  int h = 0;
  FSB_CUDA(cudaMemcpyAsync(&h, p->err, sizeof h, cudaMemcpyDeviceToHost, st));
  FSB_CUDA(cudaStreamSynchronize(st));
  if (h) return fsb_set_error(FSB_ENCCL, "peer all-gather: a rank did not arrive (timeout)");
  return FSB_OK;
}

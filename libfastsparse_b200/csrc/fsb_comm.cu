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

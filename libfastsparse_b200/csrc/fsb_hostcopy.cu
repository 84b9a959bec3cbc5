// fsb_hostcopy.cu -- host <-> device copies for the host-pointer (drop-in) entry points.
//
// A C caller of the reference allocates every operand with malloc (bench_a_mul_b.c:125-139): PAGEABLE memory.
// cudaMemcpyAsync from / to pageable memory degrades to the driver's synchronous staged copy (one thread, small
// chunks: a fraction of the PCIe rate), and it cannot overlap anything.  Here pageable operands go through a
// ring of pinned bounce buffers: the DMA engine moves bounce <-> HBM at wire speed while a few host threads move
// bounce <-> the caller's array, so the two legs overlap and the host leg is not one core's memcpy rate.
// Pinned or registered pointers (torch pinned tensors, cudaHostAlloc, cudaHostRegister) are detected with
// cudaPointerGetAttributes and copied directly.  Nothing is ever registered behind the caller's back, so there is
// no lifetime coupling with the caller's malloc / free.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include <emmintrin.h>
#include <omp.h>

#include "fsb_internal.h"

namespace {

// Small pieces on purpose: the DMA engine writes a piece into host memory (through the last-level cache where the
// platform has DDIO) and a few microseconds later the copy threads read it back -- a 4 MB piece is still cache
// resident then, a 32 MB one has been evicted to DRAM and costs the host a second trip over its memory bus.  The
// copy out of the bounce buffer uses non-temporal stores (no read-for-ownership of the destination lines).  Measured
// on the bench box (16 vCPUs), 2.56 GB result: 32 MB pieces + memcpy 92 ms per product against 51 ms pinned.
constexpr int kBuffers = 6;
size_t bounce_bytes() {          // FSB_BOUNCE_KB overrides the piece size (experiments); fixed after the first transfer
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("FSB_BOUNCE_KB");
    v = (size_t)std::max(64L, e ? atol(e) : 4096L) << 10;
  }
  return v;
}
#define kBounceBytes (bounce_bytes())

struct Ring {
  char* buf[kBuffers] = {};
  cudaEvent_t ev[kBuffers] = {};
  bool ready = false;
};
Ring g_ring;
std::mutex g_ring_mu;   // one pageable transfer at a time (the ring is shared)

int ring_init() {
  if (g_ring.ready) return FSB_OK;
  for (int i = 0; i < kBuffers; ++i) {
    FSB_CUDA(cudaHostAlloc((void**)&g_ring.buf[i], kBounceBytes, cudaHostAllocDefault));
    FSB_CUDA(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
  }
  g_ring.ready = true;
  return FSB_OK;
}

int copy_threads() {
  static int n = 0;
  if (n == 0) {
    const char* e = getenv("FSB_COPY_THREADS");
    int want = e ? atoi(e) : 8;
    int hw = (int)std::thread::hardware_concurrency();
    // one process per GPU: the ranks of a node share its cores (8 ranks x 8 copy threads on 32 cores: 188 ms per
    // product instead of 30, profiles/r2h_bench_c2_n8.json)
    if (hw > 0 && fsb_comm_size() > 1) hw = std::max(2, hw / fsb_comm_size());
    if (hw > 0) want = std::min(want, hw);
    n = std::max(want, 1);
  }
  return n;
}

// copy with non-temporal (streaming) stores: the destination is written once and not read again by this thread
void stream_copy(char* dst, const char* src, size_t n) {
  size_t i = 0;
  const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
  if (head && head <= n) { memcpy(dst, src, head); i = head; }
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
    const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
    _mm_stream_si128((__m128i*)(dst + i), a); _mm_stream_si128((__m128i*)(dst + i + 16), b);
    _mm_stream_si128((__m128i*)(dst + i + 32), c); _mm_stream_si128((__m128i*)(dst + i + 48), d);
  }
  if (i < n) memcpy(dst + i, src + i, n - i);
  _mm_sfence();
}

// bounce <-> caller memory with a few threads (explicit num_threads: torchrun exports OMP_NUM_THREADS=1);
// streaming = true for bounce -> caller (the D2H leg)
void host_copy(void* dst, const void* src, size_t bytes, bool streaming) {
  const int nt = copy_threads();
  const size_t piece = (size_t)256 << 10;
  if (nt == 1 || bytes < 2 * piece) {
    if (streaming) stream_copy((char*)dst, (const char*)src, bytes); else memcpy(dst, src, bytes);
    return;
  }
  const long np = (long)((bytes + piece - 1) / piece);
#pragma omp parallel for num_threads(nt) schedule(static)
  for (long p = 0; p < np; ++p) {
    const size_t off = (size_t)p * piece, n = std::min(piece, bytes - off);
    if (streaming) stream_copy((char*)dst + off, (const char*)src + off, n); else memcpy((char*)dst + off, (const char*)src + off, n);
  }
}

}  // namespace

bool fsb_host_is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// dst (device) <- src (host).  Pinned source: one asynchronous copy on st.  Pageable source: pipelined through the
// ring; returns when the source has been read completely (the device copies are still ordered on st).
int fsb_h2d(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (!bytes) return FSB_OK;
  if (!fsb_host_is_pageable(src) || bytes < ((size_t)1 << 20)) {
    FSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return FSB_OK;
  }
  std::lock_guard<std::mutex> lk(g_ring_mu);
  FSB_TRY(ring_init());
  int k = 0;
  for (size_t off = 0; off < bytes; off += kBounceBytes, k = (k + 1) % kBuffers) {
    const size_t n = std::min(kBounceBytes, bytes - off);
    if (off >= kBuffers * kBounceBytes) FSB_CUDA(cudaEventSynchronize(g_ring.ev[k]));   // the copy that last used this buffer
    host_copy(g_ring.buf[k], (const char*)src + off, n, false);
    FSB_CUDA(cudaMemcpyAsync((char*)dst + off, g_ring.buf[k], n, cudaMemcpyHostToDevice, st));
    FSB_CUDA(cudaEventRecord(g_ring.ev[k], st));
  }
  // the ring may be reused by the next transfer only when these copies have drained it
  for (int i = 0; i < kBuffers; ++i) FSB_CUDA(cudaEventSynchronize(g_ring.ev[i]));
  return FSB_OK;
}

// Copies `nseg` device segments to host memory in order.  Segment i may be read once ready[i] (an event recorded on
// the producing stream; nullptr = already ordered on st) has fired.  Pinned destination: asynchronous copies on st,
// returns immediately.  Pageable destination: pipelined through the ring, returns when every byte is in place.
int fsb_d2h_segments(int nseg, void* const* dst, const void* const* src, const size_t* bytes, const cudaEvent_t* ready, cudaStream_t st) {
  bool pageable = false;
  for (int i = 0; i < nseg && !pageable; ++i) pageable = bytes[i] >= ((size_t)1 << 20) && fsb_host_is_pageable(dst[i]);
  if (!pageable) {
    for (int i = 0; i < nseg; ++i) {
      if (!bytes[i]) continue;
      if (ready && ready[i]) FSB_CUDA(cudaStreamWaitEvent(st, ready[i], 0));
      FSB_CUDA(cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDeviceToHost, st));
    }
    return FSB_OK;
  }
  std::lock_guard<std::mutex> lk(g_ring_mu);
  FSB_TRY(ring_init());
  // flatten into pieces of at most one bounce buffer
  struct Piece { char* h; const char* d; size_t n; int seg; };
  size_t total = 0;
  for (int i = 0; i < nseg; ++i) total += (bytes[i] + kBounceBytes - 1) / kBounceBytes;
  Piece* pc = (Piece*)malloc(std::max<size_t>(total, 1) * sizeof(Piece));
  if (!pc) return fsb_set_error(FSB_ENOMEM, "fsb_d2h_segments: out of host memory");
  size_t np = 0;
  for (int i = 0; i < nseg; ++i)
    for (size_t off = 0; off < bytes[i]; off += kBounceBytes) pc[np++] = Piece{(char*)dst[i] + off, (const char*)src[i] + off, std::min(kBounceBytes, bytes[i] - off), i};
  int rc = FSB_OK;
  int waited_seg = -1;
  auto enqueue = [&](size_t p) -> cudaError_t {
    const int k = (int)(p % kBuffers);
    if (ready && pc[p].seg != waited_seg) {
      waited_seg = pc[p].seg;
      if (ready[waited_seg]) { cudaError_t e = cudaStreamWaitEvent(st, ready[waited_seg], 0); if (e != cudaSuccess) return e; }
    }
    cudaError_t e = cudaMemcpyAsync(g_ring.buf[k], pc[p].d, pc[p].n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(g_ring.ev[k], st);
    return e;
  };
  cudaError_t e = cudaSuccess;
  for (size_t p = 0; p < np && p < (size_t)kBuffers && e == cudaSuccess; ++p) e = enqueue(p);
  for (size_t p = 0; p < np && e == cudaSuccess; ++p) {
    const int k = (int)(p % kBuffers);
    e = cudaEventSynchronize(g_ring.ev[k]);
    if (e != cudaSuccess) break;
    host_copy(pc[p].h, g_ring.buf[k], pc[p].n, true);
    if (p + kBuffers < np) e = enqueue(p + kBuffers);
  }
  if (e != cudaSuccess) rc = fsb_cuda_error(e, "fsb_d2h_segments", __FILE__, __LINE__);
  free(pc);
  return rc;
}

int fsb_d2h(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  void* d[1] = {dst};
  const void* s[1] = {src};
  size_t b[1] = {bytes};
  return fsb_d2h_segments(1, d, s, b, nullptr, st);
}

// Many small host arrays <-> one contiguous device range (the per-block arrays of BlockedSBM / BlockedSDM: tens of
// thousands of ~40 KB malloc'd pieces).  One cudaMemcpyAsync per piece from pageable memory costs ~8 us each; here the
// pieces are packed into / unpacked from the pinned ring, so the DMA engine sees 4 MB transfers.
namespace {
struct Seg { const char* host; size_t ring_off, len; };
}

int fsb_h2d_gather(void* dst, const void* const* srcs, const size_t* bytes, long n, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ring_mu);
  FSB_TRY(ring_init());
  std::vector<Seg> segs;
  size_t dev_off = 0, fill = 0;
  long piece = 0;
  auto flush = [&]() -> int {
    if (!fill) return FSB_OK;
    const int k = (int)(piece % kBuffers);
    if (piece >= kBuffers) FSB_CUDA(cudaEventSynchronize(g_ring.ev[k]));     // the copy that last used this buffer
    char* buf = g_ring.buf[k];
    const long ns = (long)segs.size();
#pragma omp parallel for num_threads(copy_threads()) schedule(static) if (fill >= ((size_t)1 << 20))
    for (long i = 0; i < ns; ++i) memcpy(buf + segs[(size_t)i].ring_off, segs[(size_t)i].host, segs[(size_t)i].len);
    FSB_CUDA(cudaMemcpyAsync((char*)dst + dev_off, buf, fill, cudaMemcpyHostToDevice, st));
    FSB_CUDA(cudaEventRecord(g_ring.ev[k], st));
    dev_off += fill; fill = 0; segs.clear(); ++piece;
    return FSB_OK;
  };
  for (long i = 0; i < n; ++i) {
    size_t done = 0;
    while (done < bytes[i]) {
      const size_t take = std::min(bytes[i] - done, kBounceBytes - fill);
      segs.push_back(Seg{(const char*)srcs[i] + done, fill, take});
      fill += take; done += take;
      if (fill == kBounceBytes) FSB_TRY(flush());
    }
  }
  FSB_TRY(flush());
  for (int i = 0; i < kBuffers; ++i) FSB_CUDA(cudaEventSynchronize(g_ring.ev[i]));
  return FSB_OK;
}

int fsb_d2h_scatter(void* const* dsts, const void* src, const size_t* bytes, long n, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_ring_mu);
  FSB_TRY(ring_init());
  size_t total = 0;
  for (long i = 0; i < n; ++i) total += bytes[i];
  const long npieces = (long)((total + kBounceBytes - 1) / kBounceBytes);
  auto enqueue = [&](long p) -> cudaError_t {
    const int k = (int)(p % kBuffers);
    const size_t off = (size_t)p * kBounceBytes, len = std::min(kBounceBytes, total - off);
    cudaError_t e = cudaMemcpyAsync(g_ring.buf[k], (const char*)src + off, len, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(g_ring.ev[k], st);
    return e;
  };
  cudaError_t e = cudaSuccess;
  for (long p = 0; p < npieces && p < kBuffers && e == cudaSuccess; ++p) e = enqueue(p);
  long arr = 0;          // walk the destination arrays in step with the pieces
  size_t arr_done = 0;
  for (long p = 0; p < npieces && e == cudaSuccess; ++p) {
    const int k = (int)(p % kBuffers);
    e = cudaEventSynchronize(g_ring.ev[k]);
    if (e != cudaSuccess) break;
    const size_t len = std::min(kBounceBytes, total - (size_t)p * kBounceBytes);
    std::vector<Seg> segs;
    size_t off = 0;
    while (off < len) {
      while (arr < n && arr_done == bytes[arr]) { ++arr; arr_done = 0; }
      const size_t take = std::min(bytes[arr] - arr_done, len - off);
      segs.push_back(Seg{(const char*)dsts[arr] + arr_done, off, take});
      arr_done += take; off += take;
    }
    const char* buf = g_ring.buf[k];
    const long ns = (long)segs.size();
#pragma omp parallel for num_threads(copy_threads()) schedule(static) if (len >= ((size_t)1 << 20))
    for (long i = 0; i < ns; ++i) memcpy(const_cast<char*>(segs[(size_t)i].host), buf + segs[(size_t)i].ring_off, segs[(size_t)i].len);
    if (p + kBuffers < npieces) e = enqueue(p + kBuffers);
  }
  if (e != cudaSuccess) return fsb_cuda_error(e, "fsb_d2h_scatter", __FILE__, __LINE__);
  return FSB_OK;
}

// fsb_dropin.cpp -- residency cache behind the drop-in headers (include/fastsparse/*.h).
//
// The reference's structs are caller-owned and their layout is frozen (struct BinaryCSR is dumped raw into
// .csr.bin, csr.h:104), so a device handle cannot live inside them.  Each host structure is mapped to its HBM copy
// by the addresses of its arrays (SURVEY 8b "Ownership").  The reference reads the caller's arrays on every call,
// so a caller may edit them in place (a sampler that re-weights vals[]) or free and re-allocate them at the same
// address; the cache must never answer with the old matrix.  Validation is therefore EXACT by default:
//
//   * every entry stores a 64-bit hash of the FULL content of every array (multi-lane multiply-rotate hash over
//     1 MiB blocks, block hashes combined in order, so the value does not depend on the thread count);
//   * a lookup first compares shapes and a cheap sampled fingerprint (a mismatch is stale at once), then checks the
//     full hash.  Small structures are hashed inline.  Large ones (the hash streams the arrays once at host memory
//     speed -- about the cost of the PCIe upload it avoids) are hashed by a background worker WHILE the product
//     runs on the cached copy; the drop-in header then calls fsb_cache_settle(): if any handle it used turns out
//     stale the entry is dropped and the header repeats the call, which uploads the new content.  The result of
//     the stale run is overwritten (outputs are always fully overwritten), so the caller only ever sees a product
//     of the arrays as they are now -- at the cost of one wasted product when an edit is detected;
//   * mutating drop-in entry points (sort_*, transpose, free_*) still call fsb_cache_drop at once.
//
// Modes (environment): FSB_CACHE=1 / unset -- exact (above); FSB_CACHE=fast -- sampled fingerprint only (256 strided
// samples per array: in-place edits between samples go unnoticed; for callers that promise not to edit);
// FSB_CACHE=0 -- no caching: every call uploads, the handles live until the same thread's fsb_cache_settle().
// Eviction: least-recently-used entries are released when the cached matrices exceed FSB_CACHE_MAX_MB of HBM
// (default 65536) or 64 entries, so arrays released with plain free() do not pin HBM for the life of the process.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <omp.h>

#include "../../include/fsb.h"

namespace {

// ------------------------------------------------------------------------------------------------ hashing
inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h * 0xff51afd7ed558ccdull;
}

constexpr size_t kHashBlock = (size_t)1 << 20;

// one block: 4 independent lanes of (lane + word * P1) rotl 31 * P2 (the xxh64 round), then folded
uint64_t hash_block(const unsigned char* p, size_t n) {
  const uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full;
  uint64_t a = 1, b = 2, c = 3, d = 4;
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    uint64_t w[4];
    memcpy(w, p + i, 32);
    a = rotl(a + w[0] * P2, 31) * P1;
    b = rotl(b + w[1] * P2, 31) * P1;
    c = rotl(c + w[2] * P2, 31) * P1;
    d = rotl(d + w[3] * P2, 31) * P1;
  }
  uint64_t h = rotl(a, 1) + rotl(b, 7) + rotl(c, 12) + rotl(d, 18) + (uint64_t)n;
  for (; i < n; ++i) h = (h ^ p[i]) * P1;
  h ^= h >> 33; h *= P2; h ^= h >> 29;
  return h;
}

int hash_threads() {
  static int n = 0;
  if (!n) {
    const char* e = getenv("FSB_HASH_THREADS");
    int want = e ? atoi(e) : 6;    // the hash shares the host with the copy threads of the product it runs beside
    const int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0) want = std::min(want, hw);
    n = std::max(want, 1);
  }
  return n;
}

uint64_t hash_bytes(uint64_t seed, const void* ptr, size_t bytes) {
  if (!ptr || !bytes) return mix(seed, 0x5eed);
  const unsigned char* p = (const unsigned char*)ptr;
  const long nb = (long)((bytes + kHashBlock - 1) / kHashBlock);
  if (nb == 1) return mix(seed, hash_block(p, bytes));
  std::vector<uint64_t> hb((size_t)nb);
#pragma omp parallel for num_threads(hash_threads()) schedule(static)
  for (long b = 0; b < nb; ++b) hb[(size_t)b] = hash_block(p + (size_t)b * kHashBlock, std::min(kHashBlock, bytes - (size_t)b * kHashBlock));
  uint64_t h = seed;
  for (long b = 0; b < nb; ++b) h = mix(h, hb[(size_t)b]);
  return h;
}

template <typename T>
uint64_t sample(uint64_t h, const T* p, long n) {
  if (!p || n <= 0) return mix(h, 0x5eed);
  const long step = n > 256 ? n / 256 : 1;
  for (long i = 0; i < n; i += step) {
    uint64_t bits = 0;
    memcpy(&bits, &p[i], sizeof(T));
    h = mix(h, bits);
  }
  uint64_t last = 0;
  memcpy(&last, &p[n - 1], sizeof(T));
  return mix(h, last);
}

// ------------------------------------------------------------------------------------------------ entries
enum Mode { kOff = 0, kExact = 1, kFast = 2 };
Mode cache_mode() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("FSB_CACHE");
    m = !e ? kExact : (e[0] == '0' ? kOff : ((e[0] == 'f' || e[0] == 'F') ? kFast : kExact));
  }
  return (Mode)m;
}

struct Entry {
  int kind;                 // 1 csr, 2 coo, 3 cbcsr, 4 blocked
  const void* k0;
  const void* k1;
  const void* k2;
  long nnz;
  int nrow, ncol, extra;
  uint64_t print;           // sampled fingerprint
  uint64_t full;            // hash of the full content
  fsb_matrix_t h;
  uint64_t stamp;           // LRU clock
  int pins;                 // pending validations / users between a lookup and fsb_cache_settle
};

std::vector<Entry*> g_entries;
std::mutex g_mu;
uint64_t g_clock = 0;

// ---- background hashing: one persistent worker (its OpenMP team persists with it)
struct Job {
  std::function<uint64_t()> fn;
  uint64_t result = 0;
  bool done = false;
  std::mutex mu;
  std::condition_variable cv;
};
// the worker is detached and may be waiting when the process exits: its queue state is never destroyed
std::mutex& g_q_mu = *new std::mutex;
std::condition_variable& g_q_cv = *new std::condition_variable;
std::vector<std::shared_ptr<Job>>& g_queue = *new std::vector<std::shared_ptr<Job>>;
bool g_worker_started = false;

void worker_main() {
  for (;;) {
    std::shared_ptr<Job> j;
    {
      std::unique_lock<std::mutex> lk(g_q_mu);
      g_q_cv.wait(lk, [] { return !g_queue.empty(); });
      j = g_queue.front();
      g_queue.erase(g_queue.begin());
    }
    const uint64_t r = j->fn();
    {
      std::lock_guard<std::mutex> lk(j->mu);
      j->result = r;
      j->done = true;
    }
    j->cv.notify_all();
  }
}

std::shared_ptr<Job> submit(std::function<uint64_t()> fn) {
  auto j = std::make_shared<Job>();
  j->fn = std::move(fn);
  {
    std::lock_guard<std::mutex> lk(g_q_mu);
    if (!g_worker_started) {
      std::thread(worker_main).detach();
      g_worker_started = true;
    }
    g_queue.push_back(j);
  }
  g_q_cv.notify_one();
  return j;
}

// ---- what one drop-in call (one thread) has taken from the cache and not settled yet
struct Pending {
  Entry* e;                       // pinned entry (nullptr for a transient handle)
  std::shared_ptr<Job> job;       // full-hash validation in flight (may be null: validated inline)
  fsb_matrix_t transient;         // FSB_CACHE=0: handle owned by this call
};
thread_local std::vector<Pending> tl_pending;

constexpr size_t kInlineHashBytes = (size_t)4 << 20;   // below this the full hash is computed inline (< 1 ms)

size_t max_cached_bytes() {
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("FSB_CACHE_MAX_MB");
    v = (size_t)(e ? atol(e) : 65536) << 20;
  }
  return v;
}

void drop_entry_locked(size_t i) {
  fsb_matrix_free(g_entries[i]->h);
  delete g_entries[i];
  g_entries.erase(g_entries.begin() + (long)i);
}

void evict_locked(const Entry* keep) {
  for (;;) {
    size_t total = 0;
    for (Entry* e : g_entries) total += (size_t)fsb_matrix_bytes(e->h);
    if (total <= max_cached_bytes() && g_entries.size() <= 64) return;
    long victim = -1;
    for (size_t i = 0; i < g_entries.size(); ++i) {
      Entry* e = g_entries[i];
      if (e == keep || e->pins > 0) continue;
      if (victim < 0 || e->stamp < g_entries[(size_t)victim]->stamp) victim = (long)i;
    }
    if (victim < 0) return;
    drop_entry_locked((size_t)victim);
  }
}

// Common body of the fsb_cache_* entry points.
//   bytes      : total size of the arrays (decides inline vs background hashing)
//   fingerprint: cheap sampled fingerprint, computed by the caller
//   full_hash  : computes the full-content hash (may run on the worker: must only touch the caller's arrays)
//   upload     : makes a fresh device handle
fsb_matrix_t acquire(int kind, const void* k0, const void* k1, const void* k2, long nnz, int nrow, int ncol, int extra,
                     size_t bytes, uint64_t fingerprint, std::function<uint64_t()> full_hash,
                     std::function<int(fsb_matrix_t*)> upload) {
  const Mode mode = cache_mode();
  if (mode == kOff) {
    fsb_matrix_t h = nullptr;
    if (upload(&h) != FSB_OK) return nullptr;
    tl_pending.push_back(Pending{nullptr, nullptr, h});   // released by this thread's fsb_cache_settle()
    return h;
  }
  std::unique_lock<std::mutex> lk(g_mu);
  Entry* hit = nullptr;
  for (size_t i = 0; i < g_entries.size(); ++i) {
    Entry* e = g_entries[i];
    if (e->kind == kind && e->k0 == k0 && e->k1 == k1 && e->k2 == k2) {
      const bool same = e->nnz == nnz && e->nrow == nrow && e->ncol == ncol && e->extra == extra && e->print == fingerprint;
      if (same) { hit = e; break; }
      if (e->pins == 0) drop_entry_locked(i);       // same addresses, different content: stale
      else e->k0 = e->k1 = e->k2 = nullptr;        // still in use by another call: orphan it, LRU will take it
      break;
    }
  }
  if (hit) {
    hit->stamp = ++g_clock;
    if (mode == kFast) return hit->h;
    if (bytes <= kInlineHashBytes) {
      hit->pins++;               // keep the entry alive while the lock is released
      lk.unlock();
      const uint64_t now = full_hash();
      lk.lock();
      hit->pins--;
      if (now == hit->full) return hit->h;
      for (size_t i = 0; i < g_entries.size(); ++i)
        if (g_entries[i] == hit) {
          if (hit->pins == 0) drop_entry_locked(i); else hit->k0 = hit->k1 = hit->k2 = nullptr;
          break;
        }
      hit = nullptr;
    } else {
      // optimistic: hand the cached copy out now, verify while the product runs; fsb_cache_settle() decides
      hit->pins++;
      tl_pending.push_back(Pending{hit, submit(full_hash), nullptr});
      return hit->h;
    }
  }
  lk.unlock();
  // miss (or stale): upload and hash the content as it is now
  fsb_matrix_t h = nullptr;
  std::shared_ptr<Job> job;
  if (mode == kExact && bytes > kInlineHashBytes) job = submit(full_hash);   // hash while the upload runs
  const int rc = upload(&h);
  uint64_t full = 0;
  if (job) {
    std::unique_lock<std::mutex> jl(job->mu);
    job->cv.wait(jl, [&] { return job->done; });
    full = job->result;
  } else if (mode == kExact && rc == FSB_OK) {
    full = full_hash();
  }
  if (rc != FSB_OK) return nullptr;
  lk.lock();
  Entry* e = new Entry{kind, k0, k1, k2, nnz, nrow, ncol, extra, fingerprint, full, h, ++g_clock, 0};
  g_entries.push_back(e);
  evict_locked(e);
  return h;
}

}  // namespace

extern "C" {

fsb_matrix_t fsb_cache_csr(int nrow, int ncol, long nnz, const int* row_ptr, const int* cols, const double* vals) {
  const uint64_t fp = sample(sample(sample(0x1, row_ptr, (long)nrow + 1), cols, nnz), vals, vals ? nnz : 0);
  const size_t brp = ((size_t)nrow + 1) * 4, bc = (size_t)std::max(nnz, 0L) * 4, bv = vals ? (size_t)std::max(nnz, 0L) * 8 : 0;
  return acquire(1, row_ptr, cols, vals, nnz, nrow, ncol, 0, brp + bc + bv, fp,
                 [=] { return hash_bytes(hash_bytes(hash_bytes(0x11, row_ptr, brp), cols, bc), vals, bv); },
                 [=](fsb_matrix_t* h) { return fsb_csr_upload(h, nrow, ncol, nnz, row_ptr, cols, vals); });
}

fsb_matrix_t fsb_cache_coo(int nrow, int ncol, long nnz, const int* rows, const int* cols, const double* vals) {
  const uint64_t fp = sample(sample(sample(0x2, rows, nnz), cols, nnz), vals, vals ? nnz : 0);
  const size_t bi = (size_t)std::max(nnz, 0L) * 4, bv = vals ? (size_t)std::max(nnz, 0L) * 8 : 0;
  return acquire(2, rows, cols, vals, nnz, nrow, ncol, 0, 2 * bi + bv, fp,
                 [=] { return hash_bytes(hash_bytes(hash_bytes(0x22, rows, bi), cols, bi), vals, bv); },
                 [=](fsb_matrix_t* h) { return fsb_csr_upload_coo(h, nrow, ncol, nnz, rows, cols, vals); });
}

fsb_matrix_t fsb_cache_cbcsr(int nrow, int ncol, int nblocks, int colblocksize, long nnz, const int* row_ptr, const int* cols) {
  const long ncell1 = (long)nblocks * nrow + 1;
  const uint64_t fp = sample(sample(0x3, row_ptr, ncell1), cols, nnz);
  const size_t brp = (size_t)ncell1 * 4, bc = (size_t)std::max(nnz, 0L) * 4;
  return acquire(3, row_ptr, cols, nullptr, nnz, nrow, ncol, nblocks, brp + bc, fp,
                 [=] { return hash_bytes(hash_bytes(0x33 + (uint64_t)colblocksize, row_ptr, brp), cols, bc); },
                 [=](fsb_matrix_t* h) { return fsb_cbcsr_upload(h, nrow, ncol, nblocks, colblocksize, nnz, row_ptr, cols); });
}

fsb_matrix_t fsb_cache_blocked(int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz,
                               int* const* rows, int* const* cols, double* const* vals) {
  uint64_t fp = sample(sample(0x4, start_row, (long)nblocks + 1), blk_nnz, nblocks);
  long nnz = 0;
  for (int b = 0; b < nblocks; ++b) {
    nnz += blk_nnz[b];
    fp = mix(fp, (uint64_t)(uintptr_t)rows[b]);
    const int m = blk_nnz[b];
    if (m > 0) {   // first, middle and last entry of every block (re-sorting a block moves them)
      const int pick[3] = {0, m / 2, m - 1};
      for (int q = 0; q < 3; ++q) {
        fp = mix(fp, ((uint64_t)(uint32_t)rows[b][pick[q]] << 32) | (uint32_t)cols[b][pick[q]]);
        if (vals) { uint64_t bits; memcpy(&bits, &vals[b][pick[q]], 8); fp = mix(fp, bits); }
      }
    }
  }
  const size_t bytes = (size_t)nnz * (vals ? 16 : 8) + ((size_t)nblocks * 2 + 1) * 4;
  auto full = [=] {
    // per-block hashes in parallel (blocks are many and small), combined in block order
    std::vector<uint64_t> hb((size_t)std::max(nblocks, 1), 0);
#pragma omp parallel for num_threads(hash_threads()) schedule(dynamic, 16)
    for (int b = 0; b < nblocks; ++b) {
      const size_t m = (size_t)blk_nnz[b];
      uint64_t h = mix(0x44, hash_block((const unsigned char*)rows[b], m * 4));
      h = mix(h, hash_block((const unsigned char*)cols[b], m * 4));
      if (vals) h = mix(h, hash_block((const unsigned char*)vals[b], m * 8));
      hb[(size_t)b] = h;
    }
    uint64_t h = hash_bytes(hash_bytes(0x44, start_row, ((size_t)nblocks + 1) * 4), blk_nnz, (size_t)nblocks * 4);
    for (int b = 0; b < nblocks; ++b) h = mix(h, hb[(size_t)b]);
    return h;
  };
  return acquire(4, start_row, rows, vals, nnz, nrow, ncol, nblocks, bytes, fp, full,
                 [=](fsb_matrix_t* h) { return fsb_blocked_upload(h, nrow, ncol, nblocks, start_row, blk_nnz, rows, cols, vals); });
}

// End of one drop-in call on this thread: waits for the validations started by its fsb_cache_* lookups and
// releases what the call held.  Returns the number of handles that turned out STALE (their entries are dropped):
// the caller must then repeat the call -- the next lookup uploads the current content.
int fsb_cache_settle(void) {
  int stale = 0;
  std::vector<Pending> mine;
  mine.swap(tl_pending);
  for (Pending& p : mine) {
    if (p.transient) { fsb_matrix_free(p.transient); continue; }
    uint64_t now = 0;
    if (p.job) {
      std::unique_lock<std::mutex> lk(p.job->mu);
      p.job->cv.wait(lk, [&] { return p.job->done; });
      now = p.job->result;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    p.e->pins--;
    const bool is_stale = p.job && now != p.e->full;
    if (is_stale) ++stale;
    const bool orphan = p.e->k0 == nullptr;        // dropped or found stale by another call while this one used it
    if (is_stale || orphan) {
      for (size_t i = 0; i < g_entries.size(); ++i)
        if (g_entries[i] == p.e) {
          if (p.e->pins == 0) drop_entry_locked(i); else p.e->k0 = p.e->k1 = p.e->k2 = nullptr;
          break;
        }
    }
  }
  return stale;
}

void fsb_cache_drop(const void* key_ptr) {
  if (!key_ptr) return;
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < g_entries.size();) {
    Entry* e = g_entries[i];
    if (e->k0 == key_ptr || e->k1 == key_ptr || e->k2 == key_ptr) {
      if (e->pins == 0) { drop_entry_locked(i); continue; }
      e->k0 = e->k1 = e->k2 = nullptr;
    }
    ++i;
  }
}

void fsb_cache_clear(void) {
  fsb_cache_settle();
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = g_entries.size(); i-- > 0;)
    if (g_entries[i]->pins == 0) drop_entry_locked(i);
}

/* cache statistics for tests: entries held, HBM bytes they pin */
int fsb_cache_stats(long* entries, long* bytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  long b = 0;
  for (Entry* e : g_entries) b += fsb_matrix_bytes(e->h);
  if (entries) *entries = (long)g_entries.size();
  if (bytes) *bytes = b;
  return FSB_OK;
}

}  // extern "C"

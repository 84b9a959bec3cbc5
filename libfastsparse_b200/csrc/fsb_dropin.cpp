// fsb_dropin.cpp -- residency cache behind the drop-in headers (include/fastsparse/*.h).
//
// The reference's structs are caller-owned and their layout is frozen (struct BinaryCSR
// is dumped raw into .csr.bin, csr.h:104), so a device handle cannot live inside them.
// Instead each host structure is mapped to its HBM copy by the addresses of its arrays
// (SURVEY 8b "Ownership").  An entry is (re)validated with a cheap fingerprint -- the
// shape plus up to 256 strided samples of every array -- so that a structure that was
// re-sorted in place, or freed and re-allocated at the same address, is uploaded again
// instead of silently reusing stale data.  Mutating drop-in entry points (sort_*,
// transpose, free_*) also call fsb_cache_drop explicitly.  FSB_CACHE=0 disables the
// cache (upload on every call).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/fsb.h"

namespace {

struct Entry {
  int kind;                 // 1 csr, 2 coo, 3 cbcsr, 4 blocked
  const void* k0;
  const void* k1;
  const void* k2;
  long nnz;
  int nrow, ncol, extra;
  uint64_t print;
  fsb_matrix_t h;
};

std::vector<Entry> g_entries;
std::mutex g_mu;

bool cache_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FSB_CACHE");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h * 0xff51afd7ed558ccdull;
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

fsb_matrix_t lookup(int kind, const void* k0, const void* k1, const void* k2, long nnz, int nrow, int ncol, int extra, uint64_t print) {
  for (size_t i = 0; i < g_entries.size(); ++i) {
    Entry& e = g_entries[i];
    if (e.kind == kind && e.k0 == k0 && e.k1 == k1 && e.k2 == k2) {
      if (e.nnz == nnz && e.nrow == nrow && e.ncol == ncol && e.extra == extra && e.print == print) return e.h;
      fsb_matrix_free(e.h);             // same addresses, different content: stale
      g_entries.erase(g_entries.begin() + i);
      return nullptr;
    }
  }
  return nullptr;
}

void remember(int kind, const void* k0, const void* k1, const void* k2, long nnz, int nrow, int ncol, int extra, uint64_t print, fsb_matrix_t h) {
  g_entries.push_back(Entry{kind, k0, k1, k2, nnz, nrow, ncol, extra, print, h});
}

// one uncached handle kept alive until the next call when caching is off
fsb_matrix_t g_transient = nullptr;
fsb_matrix_t transient(fsb_matrix_t h) {
  if (g_transient) fsb_matrix_free(g_transient);
  g_transient = h;
  return h;
}

}  // namespace

extern "C" {

fsb_matrix_t fsb_cache_csr(int nrow, int ncol, long nnz, const int* row_ptr, const int* cols, const double* vals) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t fp = sample(sample(sample(0x1, row_ptr, (long)nrow + 1), cols, nnz), vals, vals ? nnz : 0);
  fsb_matrix_t h = cache_enabled() ? lookup(1, row_ptr, cols, vals, nnz, nrow, ncol, 0, fp) : nullptr;
  if (h) return h;
  if (fsb_csr_upload(&h, nrow, ncol, nnz, row_ptr, cols, vals) != FSB_OK) return nullptr;
  if (!cache_enabled()) return transient(h);
  remember(1, row_ptr, cols, vals, nnz, nrow, ncol, 0, fp, h);
  return h;
}

fsb_matrix_t fsb_cache_coo(int nrow, int ncol, long nnz, const int* rows, const int* cols, const double* vals) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t fp = sample(sample(sample(0x2, rows, nnz), cols, nnz), vals, vals ? nnz : 0);
  fsb_matrix_t h = cache_enabled() ? lookup(2, rows, cols, vals, nnz, nrow, ncol, 0, fp) : nullptr;
  if (h) return h;
  if (fsb_csr_upload_coo(&h, nrow, ncol, nnz, rows, cols, vals) != FSB_OK) return nullptr;
  if (!cache_enabled()) return transient(h);
  remember(2, rows, cols, vals, nnz, nrow, ncol, 0, fp, h);
  return h;
}

fsb_matrix_t fsb_cache_cbcsr(int nrow, int ncol, int nblocks, int colblocksize, long nnz, const int* row_ptr, const int* cols) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t fp = sample(sample(0x3, row_ptr, (long)nblocks * nrow + 1), cols, nnz);
  fsb_matrix_t h = cache_enabled() ? lookup(3, row_ptr, cols, nullptr, nnz, nrow, ncol, nblocks, fp) : nullptr;
  if (h) return h;
  if (fsb_cbcsr_upload(&h, nrow, ncol, nblocks, colblocksize, nnz, row_ptr, cols) != FSB_OK) return nullptr;
  if (!cache_enabled()) return transient(h);
  remember(3, row_ptr, cols, nullptr, nnz, nrow, ncol, nblocks, fp, h);
  return h;
}

fsb_matrix_t fsb_cache_blocked(int nrow, int ncol, int nblocks, const int* start_row, const int* blk_nnz,
                               int* const* rows, int* const* cols, double* const* vals) {
  std::lock_guard<std::mutex> lk(g_mu);
  uint64_t fp = sample(sample(0x4, start_row, (long)nblocks + 1), blk_nnz, nblocks);
  long nnz = 0;
  for (int b = 0; b < nblocks; ++b) {
    nnz += blk_nnz[b];
    fp = mix(fp, (uint64_t)(uintptr_t)rows[b]);
    // every block: first, middle and last entry (cheap, and re-sorting a block moves them)
    const int m = blk_nnz[b];
    if (m > 0) {
      const int pick[3] = {0, m / 2, m - 1};
      for (int q = 0; q < 3; ++q) {
        fp = mix(fp, ((uint64_t)(uint32_t)rows[b][pick[q]] << 32) | (uint32_t)cols[b][pick[q]]);
        if (vals) { uint64_t bits; memcpy(&bits, &vals[b][pick[q]], 8); fp = mix(fp, bits); }
      }
    }
  }
  fsb_matrix_t h = cache_enabled() ? lookup(4, start_row, rows, vals, nnz, nrow, ncol, nblocks, fp) : nullptr;
  if (h) return h;
  if (fsb_blocked_upload(&h, nrow, ncol, nblocks, start_row, blk_nnz, rows, cols, vals) != FSB_OK) return nullptr;
  if (!cache_enabled()) return transient(h);
  remember(4, start_row, rows, vals, nnz, nrow, ncol, nblocks, fp, h);
  return h;
}

void fsb_cache_drop(const void* key_ptr) {
  if (!key_ptr) return;
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < g_entries.size();) {
    Entry& e = g_entries[i];
    if (e.k0 == key_ptr || e.k1 == key_ptr || e.k2 == key_ptr) {
      fsb_matrix_free(e.h);
      g_entries.erase(g_entries.begin() + i);
    } else {
      ++i;
    }
  }
}

void fsb_cache_clear(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (Entry& e : g_entries) fsb_matrix_free(e.h);
  g_entries.clear();
  if (g_transient) { fsb_matrix_free(g_transient); g_transient = nullptr; }
}

}  // extern "C"

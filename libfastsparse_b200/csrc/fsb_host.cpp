// fsb_host.cpp -- host-side structure of the drop-in: constructors, Hilbert maths,
// sorting and file formats.  SURVEY 8b: "construction / load entry points stay host C
// (bit-exact)".  None of this needs a GPU, none of it is on the measured hot path, and
// none of it computes a sparse x dense product (there is no CPU fallback for those).
//
// Every routine reproduces the reference's OUTPUT exactly (array contents, orders,
// file bytes); the code is written for this library: flat counting sorts with an
// explicit cursor array, an iterative quicksort, fixed-width file headers.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/fsb.h"

int fsb_set_error(int code, const char* fmt, ...);

extern "C" {

// ---------------------------------------------------------------- Hilbert maths
// hilbert.h:11-13: 1 << (int)ceil(log2(x)) == smallest power of two >= x for x >= 1
int fsb_host_ceil_pow2(int x) {
  if (x <= 1) return 1;
  return 1 << (32 - __builtin_clz((unsigned)(x - 1)));
}

// hilbert.h:16-27 with the quadrant transform of hilbert.h:45-57 applied with the
// current sub-square size (the reference's convention; x, y are not masked first)
long fsb_host_xy2d(int n, int x, int y) {
  long d = 0;
  for (long s = n / 2; s > 0; s /= 2) {
    const int rx = (x & s) > 0, ry = (y & s) > 0;
    d += s * s * (long)((3 * rx) ^ ry);
    if (!ry) {
      if (rx) { x = (int)s - 1 - x; y = (int)s - 1 - y; }
      const int t = x; x = y; y = t;
    }
  }
  return d;
}

// hilbert.h:30-42
void fsb_host_d2xy(int n, long d, int* xo, int* yo) {
  int x = 0, y = 0;
  long t = d;
  for (int s = 1; s < n; s *= 2) {
    const int rx = (int)(1 & (t / 2));
    const int ry = (int)(1 & (t ^ rx));
    if (!ry) {
      if (rx) { x = s - 1 - x; y = s - 1 - y; }
      const int w = x; x = y; y = w;
    }
    x += s * rx;
    y += s * ry;
    t /= 4;
  }
  *xo = x;
  *yo = y;
}

// hilbert.h:60-65
long fsb_host_row_xy2d(int n, int x, int y) {
  const long nsq = (long)n * n;
  return fsb_host_xy2d(n, y % n, x) + nsq * (long)(y / n);
}

// hilbert.h:68-75
void fsb_host_row_d2xy(int n, long d, int* x, int* y) {
  const long nsq = (long)n * n;
  const int tile = (int)(d / nsq);
  fsb_host_d2xy(n, d % nsq, y, x);
  *y += tile * n;
}

// ---------------------------------------------------------------- sorting
// Ascending sort of keys with an optional co-moved payload.  Same scheme as
// quickSort.h:10-57 / quickSortD.h:12-71 (middle element as pivot parked at the left
// end, two-pointer scan with <= / >, insertion sort for ranges shorter than 11), so the
// placement of payloads under duplicate keys matches too.  Iterative, O(log n) stack.
void fsb_host_sort_keys(long* a, double* v, long n) {
  if (n < 2) return;
  struct Range { long l, r; };
  std::vector<Range> todo;
  todo.push_back({0, n - 1});
  while (!todo.empty()) {
    Range cur = todo.back();
    todo.pop_back();
    long l = cur.l, r = cur.r;
    while (r - l >= 10) {
      const long m = (l + r) / 2;
      const long pivot = a[m];
      a[m] = a[l]; a[l] = pivot;
      if (v) { const double t = v[m]; v[m] = v[l]; v[l] = t; }
      long i = l, j = r + 1;
      for (;;) {
        do { ++i; } while (i <= r && a[i] <= pivot);
        do { --j; } while (a[j] > pivot);
        if (i >= j) break;
        const long t = a[i]; a[i] = a[j]; a[j] = t;
        if (v) { const double tv = v[i]; v[i] = v[j]; v[j] = tv; }
      }
      { const long t = a[l]; a[l] = a[j]; a[j] = t; }
      if (v) { const double tv = v[l]; v[l] = v[j]; v[j] = tv; }
      // left part now (depth first, as the reference recurses), right part later
      todo.push_back({j + 1, r});
      r = j - 1;
    }
    for (long k = l + 1; k <= r; ++k) {   // insertion sort of the short range
      const long key = a[k];
      const double pv = v ? v[k] : 0.0;
      long q = k - 1;
      while (q >= l && key < a[q]) {
        a[q + 1] = a[q];
        if (v) v[q + 1] = v[q];
        --q;
      }
      a[q + 1] = key;
      if (v) v[q + 1] = pv;
    }
  }
}

// ---------------------------------------------------------------- constructors
// csr.h:30-67 / 375-422: stable counting sort by row
int fsb_host_csr_from_coo(long nnz, int nrow, const int* rows, const int* cols, const double* vals,
                          int* row_ptr, int* out_cols, double* out_vals) {
  if (nnz < 0 || nrow < 0 || !row_ptr || (nnz > 0 && (!rows || !cols || !out_cols))) return fsb_set_error(FSB_EINVAL, "fsb_host_csr_from_coo: bad arguments");
  if (nnz > INT32_MAX) return fsb_set_error(FSB_EINVAL, "fsb_host_csr_from_coo: nnz exceeds int32 row_ptr range");
  memset(row_ptr, 0, ((size_t)nrow + 1) * sizeof(int));
  for (long i = 0; i < nnz; ++i) row_ptr[rows[i] + 1]++;
  for (int r = 0; r < nrow; ++r) row_ptr[r + 1] += row_ptr[r];
  std::vector<int> next(row_ptr, row_ptr + nrow);
  for (long i = 0; i < nnz; ++i) {
    const int d = next[rows[i]]++;
    out_cols[d] = cols[i];
    if (vals) out_vals[d] = vals[i];
  }
  return FSB_OK;
}

int fsb_host_cbcsr_nblocks(int ncol, int colblocksize) {  // cbcsr.h:27
  return (int)ceil(ncol / (double)colblocksize);
}

// cbcsr.h:16-65: stable counting sort by cell = (col / colblocksize) * nrow + row
int fsb_host_cbcsr_from_coo(int colblocksize, long nnz, int nrow, int ncol, const int* rows, const int* cols,
                            int* row_ptr, int* out_cols) {
  if (colblocksize <= 0 || nnz < 0 || nrow < 0 || ncol < 0 || !row_ptr) return fsb_set_error(FSB_EINVAL, "fsb_host_cbcsr_from_coo: bad arguments");
  const long ncell = (long)fsb_host_cbcsr_nblocks(ncol, colblocksize) * nrow;
  if (ncell >= INT32_MAX || nnz > INT32_MAX) return fsb_set_error(FSB_EINVAL, "fsb_host_cbcsr_from_coo: nblocks*nrow exceeds the int32 cell range (cbcsr.h:41)");
  memset(row_ptr, 0, ((size_t)ncell + 1) * sizeof(int));
  for (long i = 0; i < nnz; ++i) row_ptr[(long)(cols[i] / colblocksize) * nrow + rows[i] + 1]++;
  for (long c = 0; c < ncell; ++c) row_ptr[c + 1] += row_ptr[c];
  std::vector<int> next(row_ptr, row_ptr + ncell);
  for (long i = 0; i < nnz; ++i) out_cols[next[(long)(cols[i] / colblocksize) * nrow + rows[i]]++] = cols[i];
  return FSB_OK;
}

int fsb_host_blocked_nblocks(int nrow, int block_size) {  // sparse.h:179
  return (int)ceil(nrow / (double)block_size);
}

// sparse.h:175-195: block boundaries and per-block counts
int fsb_host_blocked_count(long nnz, int nrow, int block_size, const int* rows, int* start_row, int* blk_nnz) {
  if (block_size <= 0 || nnz < 0 || nrow < 0 || !start_row) return fsb_set_error(FSB_EINVAL, "fsb_host_blocked_count: bad arguments");
  const int nb = fsb_host_blocked_nblocks(nrow, block_size);
  for (int b = 0; b < nb; ++b) { start_row[b] = b * block_size; blk_nnz[b] = 0; }
  start_row[nb] = nrow;
  for (long j = 0; j < nnz; ++j) blk_nnz[rows[j] / block_size]++;
  return FSB_OK;
}

// sparse.h:196-212 / dsparse.h:153-170: bucket entries into their block, COO order kept
int fsb_host_blocked_fill(long nnz, int block_size, const int* rows, const int* cols, const double* vals, int nblocks,
                          int* const* rows_out, int* const* cols_out, double* const* vals_out) {
  if (block_size <= 0 || nnz < 0 || (nnz > 0 && (!rows || !cols || !rows_out || !cols_out))) return fsb_set_error(FSB_EINVAL, "fsb_host_blocked_fill: bad arguments");
  std::vector<int> fill((size_t)(nblocks > 0 ? nblocks : 1), 0);
  for (long j = 0; j < nnz; ++j) {
    const int b = rows[j] / block_size;
    const int k = fill[b]++;
    rows_out[b][k] = rows[j];
    cols_out[b][k] = cols[j];
    if (vals) vals_out[b][k] = vals[j];
  }
  return FSB_OK;
}

// sparse.h:142-161 (sort_sbm) / dsparse.h:96-115 (sort_sdm)
int fsb_host_sort_coo_hilbert(int nrow, int ncol, long nnz, int* rows, int* cols, double* vals) {
  if (nnz < 0 || (nnz > 0 && (!rows || !cols))) return fsb_set_error(FSB_EINVAL, "fsb_host_sort_coo_hilbert: bad arguments");
  const int n = fsb_host_ceil_pow2(nrow > ncol ? nrow : ncol);
  std::vector<long> h((size_t)nnz);
#pragma omp parallel for schedule(static)
  for (long j = 0; j < nnz; ++j) h[j] = fsb_host_xy2d(n, rows[j], cols[j]);
  fsb_host_sort_keys(h.data(), vals, nnz);
#pragma omp parallel for schedule(static)
  for (long j = 0; j < nnz; ++j) fsb_host_d2xy(n, h[j], &rows[j], &cols[j]);
  return FSB_OK;
}

// one block of sort_bsbm (sparse.h:215-236) / sort_bsdm (dsparse.h:193-216)
int fsb_host_sort_block_hilbert(int start_row, int nrows_in_block, long nnz, int* rows, int* cols, double* vals) {
  if (nnz < 0 || (nnz > 0 && (!rows || !cols))) return fsb_set_error(FSB_EINVAL, "fsb_host_sort_block_hilbert: bad arguments");
  const int n = fsb_host_ceil_pow2(nrows_in_block);
  std::vector<long> h((size_t)nnz);
  for (long j = 0; j < nnz; ++j) h[j] = fsb_host_row_xy2d(n, rows[j] - start_row, cols[j]);
  fsb_host_sort_keys(h.data(), vals, nnz);
  for (long j = 0; j < nnz; ++j) {
    fsb_host_row_d2xy(n, h[j], &rows[j], &cols[j]);
    rows[j] += start_row;
  }
  return FSB_OK;
}

// one block of sort_bsbm_byrow (sparse.h:238-256)
int fsb_host_sort_block_byrow(int ncol, long nnz, int* rows, int* cols) {
  if (nnz < 0 || ncol <= 0 || (nnz > 0 && (!rows || !cols))) return fsb_set_error(FSB_EINVAL, "fsb_host_sort_block_byrow: bad arguments");
  std::vector<long> h((size_t)nnz);
  for (long j = 0; j < nnz; ++j) h[j] = rows[j] * (long)ncol + (long)cols[j];
  fsb_host_sort_keys(h.data(), nullptr, nnz);
  for (long j = 0; j < nnz; ++j) { rows[j] = (int)(h[j] / ncol); cols[j] = (int)(h[j] % ncol); }
  return FSB_OK;
}

// ---------------------------------------------------------------- file formats
// read_sbm sparse.h:112-139 / read_sdm dsparse.h:64-93: int64 nrow, ncol, nnz; int32
// rows[nnz], cols[nnz] (1-based on disk); optional float64 vals[nnz]
// one native 8-byte integer from an open stream (read_long utils.h:4-12); *ok = 0 on a short read
long fsb_host_read_long(void* file, int* ok) {
  int64_t v = 0;
  const bool good = file && fread(&v, sizeof v, 1, static_cast<FILE*>(file)) == 1;
  if (ok) *ok = good ? 1 : 0;
  if (!good) fsb_set_error(FSB_EIO, "File reading error for a long. File is corrupt.");
  return (long)v;
}

int fsb_host_read_coo(const char* path, long* nrow, long* ncol, long* nnz, int* rows, int* cols, double* vals) {
  FILE* f = fopen(path, "rb");
  if (!f) return fsb_set_error(FSB_EIO, "File error: %s", path ? path : "(null)");
  int64_t hdr[3];
  if (fread(hdr, 8, 3, f) != 3) {
    fclose(f);
    return fsb_set_error(FSB_EIO, "File reading error for a long. File is corrupt.");
  }
  if (nrow) *nrow = hdr[0];
  if (ncol) *ncol = hdr[1];
  if (nnz) *nnz = hdr[2];
  int rc = FSB_OK;
  if (rows && cols) {
    const size_t n = (size_t)hdr[2];
    if (fread(rows, 4, n, f) != n || fread(cols, 4, n, f) != n || (vals && fread(vals, 8, n, f) != n))
      rc = fsb_set_error(FSB_EIO, "File read error: %s", path);
    else
      for (size_t i = 0; i < n; ++i) { rows[i]--; cols[i]--; }
  }
  fclose(f);
  return rc;
}

// serialize_to_file csr.h:97-113
static const char kCsrTag[] = "BINARY_CSR: struct BinaryCSR, int[nrow], int[nnz]\n";

int fsb_host_write_csr_bin(const char* path, const void* struct_image, int nrow, long nnz, const int* row_ptr, const int* cols) {
  FILE* f = fopen(path, "w+");
  if (!f) return fsb_set_error(FSB_EIO, "cannot open %s for writing", path ? path : "(null)");
  bool ok = fputs(kCsrTag, f) >= 0 && fputs("struct BinaryCSR\n", f) >= 0;
  ok = ok && fwrite(struct_image, 32, 1, f) == 1;
  ok = ok && fprintf(f, "int[%d]\n", nrow + 1) > 0;
  ok = ok && fwrite(row_ptr, 4, (size_t)nrow + 1, f) == (size_t)nrow + 1;
  ok = ok && fprintf(f, "int[%ld]\n", nnz) > 0;
  ok = ok && (nnz == 0 || fwrite(cols, 4, (size_t)nnz, f) == (size_t)nnz);
  fclose(f);
  return ok ? FSB_OK : fsb_set_error(FSB_EIO, "short write to %s", path);
}

// deserialize_from_file csr.h:117-146.  Two-call protocol: row_ptr == NULL reads the
// struct image only (so the caller can allocate nrow+1 / nnz entries).
int fsb_host_read_csr_bin(const char* path, void* struct_image, int* row_ptr, int* cols) {
  FILE* f = fopen(path, "r");
  if (!f) return fsb_set_error(FSB_EIO, "File error: %s", path ? path : "(null)");
  char line[256], want[64];
  struct Image { int nrow, ncol; long nnz; void* p0; void* p1; } img;
  int rc = FSB_OK;
  auto expect = [&](const char* s, const char* what) {
    if (!fgets(line, sizeof line, f) || strncmp(line, s, sizeof line)) {
      rc = fsb_set_error(FSB_EIO, "ERROR: could not read data from file, %s", what);
      return false;
    }
    return true;
  };
  if (expect(kCsrTag, "Invalid file format or version") && expect("struct BinaryCSR\n", "struct data corrupted")) {
    if (fread(&img, sizeof img, 1, f) != 1) rc = fsb_set_error(FSB_EIO, "ERROR: could not read data from file, struct data corrupted");
  }
  if (rc == FSB_OK) {
    memcpy(struct_image, &img, sizeof img);
    if (row_ptr) {
      snprintf(want, sizeof want, "int[%d]\n", img.nrow + 1);
      if (expect(want, "nrow data corrupted") && fread(row_ptr, 4, (size_t)img.nrow + 1, f) != (size_t)img.nrow + 1)
        rc = fsb_set_error(FSB_EIO, "ERROR: could not read data from file, nrow data corrupted");
      if (rc == FSB_OK) {
        snprintf(want, sizeof want, "int[%ld]\n", img.nnz);
        if (expect(want, "cols data corrupted") && fread(cols, 4, (size_t)img.nnz, f) != (size_t)img.nnz)
          rc = fsb_set_error(FSB_EIO, "ERROR: could not read data from file, cols data corrupted");
      }
    }
  }
  fclose(f);
  return rc;
}

}  // extern "C"

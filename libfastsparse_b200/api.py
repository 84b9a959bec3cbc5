"""Python mirror of the reference's C interface (csr.h, sparse.h, dsparse.h, cbcsr.h,
cg.h, linalg.h, hilbert.h) on top of the C ABI.  Function names, argument order (output
array first, filled in place) and error behaviour follow the reference so the parity
tests read like test_sparse.c.  Host structures hold numpy arrays; construction and
sorting run the library's bit-exact host routines (fsb_host_*), every multiply / solve
runs on the GPU through the *_host entry points (host buffers in, host buffers out).

`DeviceMatrix` is the device-pointer face of the same library for callers that keep
their dense operands in HBM (torch tensors are used purely as device memory).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import FsbError, c_dbl_p, c_int_p, c_long_p, check, handle, lib

__all__ = [
    "SparseBinaryMatrix", "SparseDoubleMatrix", "BinaryCSR", "CSR", "ColBinaryCSR", "BlockedSBM", "BlockedSDM",
    "DeviceMatrix", "new_sbm", "free_sbm", "new_transpose", "transpose", "read_sbm", "sort_sbm", "A_mul_B", "At_mul_B",
    "new_bsbm", "sort_bsbm", "sort_bsbm_byrow", "bsbm_A_mul_B", "bsbm_A_mul_B2", "bsbm_A_mul_B4", "bsbm_A_mul_Bn",
    "new_sdm", "sdm_transpose", "read_sdm", "sort_sdm", "sdm_A_mul_B", "sdm_At_mul_B", "new_bsdm", "sort_bsdm", "bsdm_A_mul_B",
    "new_bcsr", "bcsr_from_sbm", "free_bcsr", "serialize_to_file", "deserialize_from_file", "bcsr_A_mul_B", "bcsr_A_mul_B2",
    "bcsr_A_mul_B4", "bcsr_A_mul_B8", "bcsr_A_mul_B8_auto", "bcsr_A_mul_Bn", "bcsr_A_mul_B32n", "bcsr_AA_mul_B",
    "parallel_bcsr_AA_mul_B", "bcsr_At_mul_Bn", "new_csr", "free_csr", "csr_A_mul_B", "csr_A_mul_Bn", "csr_At_mul_Bn",
    "new_cbcsr", "cbcsr_from_sbm", "cbcsr_A_mul_B", "cbcsr_A_mul_Bn", "bsbm_AtA", "bsbm_cg", "bsbm_cg2", "bsbm_cgn",
    "pnormsq", "pnormsq2", "pouter2", "pdot", "pdot2sym", "solve2sym", "dist", "ceilPower2", "xy2d", "d2xy", "row_xy2d",
    "row_d2xy", "quickSort", "quickSortD", "partition_rows", "synth_coo_host", "device_count", "launch_count",
    "comm_init_from_torch", "comm_finalize", "allreduce_sum", "cg_shard_layout",
]


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dbl_p)


def _out(y, n):
    if not (isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous and y.size >= n):
        raise ValueError(f"output must be a C-contiguous float64 array with at least {n} elements")
    return y


def device_count() -> int:
    return lib().fsb_device_count()


def launch_count() -> int:
    return lib().fsb_launch_count()


# ----------------------------------------------------------------------------- host structures
class _Resident:
    """Host structure with a device twin.

    The drop-in-named functions below go through the library's residency cache exactly like the C headers
    (fsb_cache_* + fsb_cache_settle): the host arrays stay the source of truth and are re-validated by content on every
    call, so editing them in place between calls is safe.  `_dev()` is an explicit, privately owned snapshot used by
    DeviceMatrix.of() for callers that keep a matrix resident and promise not to edit the host copy."""

    _h = None

    def _cached(self):
        raise NotImplementedError

    def _drop(self):
        if self._h:
            lib().fsb_matrix_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self._drop()
        except Exception:
            pass


class SparseBinaryMatrix(_Resident):      # sparse.h:11-18
    def __init__(self, nrow, ncol, rows, cols):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.rows, self.cols = _i32(rows), _i32(cols)
        self.vals = None

    @property
    def nnz(self):
        return int(self.rows.size)

    def _dev(self):
        if not self._h:
            h = handle()
            check(lib().fsb_csr_upload_coo(C.byref(h), self.nrow, self.ncol, self.nnz, _ip(self.rows), _ip(self.cols), _dp(self.vals)))
            self._h = h
        return self._h

    def _cached(self):
        return lib().fsb_cache_coo(self.nrow, self.ncol, self.nnz, _ip(self.rows), _ip(self.cols), _dp(self.vals))


class SparseDoubleMatrix(SparseBinaryMatrix):   # dsparse.h:11-19
    def __init__(self, nrow, ncol, rows, cols, vals):
        super().__init__(nrow, ncol, rows, cols)
        self.vals = _f64(vals)


class BinaryCSR(_Resident):               # csr.h:15-22
    def __init__(self, nrow, ncol, row_ptr, cols, vals=None):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.row_ptr, self.cols = _i32(row_ptr), _i32(cols)
        self.vals = None if vals is None else _f64(vals)

    @property
    def nnz(self):
        return int(self.row_ptr[self.nrow])

    def _dev(self):
        if not self._h:
            h = handle()
            check(lib().fsb_csr_upload(C.byref(h), self.nrow, self.ncol, self.nnz, _ip(self.row_ptr), _ip(self.cols), _dp(self.vals)))
            self._h = h
        return self._h

    def _cached(self):
        return lib().fsb_cache_csr(self.nrow, self.ncol, self.nnz, _ip(self.row_ptr), _ip(self.cols), _dp(self.vals))


class CSR(BinaryCSR):                     # csr.h:358-366
    pass


class ColBinaryCSR(_Resident):            # cbcsr.h:5-14
    def __init__(self, nrow, ncol, nblocks, colblocksize, row_ptr, cols):
        self.nrow, self.ncol, self.nblocks, self.colblocksize = int(nrow), int(ncol), int(nblocks), int(colblocksize)
        self.row_ptr, self.cols = _i32(row_ptr), _i32(cols)

    @property
    def nnz(self):
        return int(self.row_ptr[self.nblocks * self.nrow])

    def _dev(self):
        if not self._h:
            h = handle()
            check(lib().fsb_cbcsr_upload(C.byref(h), self.nrow, self.ncol, self.nblocks, self.colblocksize, self.nnz,
                                         _ip(self.row_ptr), _ip(self.cols)))
            self._h = h
        return self._h

    def _cached(self):
        return lib().fsb_cache_cbcsr(self.nrow, self.ncol, self.nblocks, self.colblocksize, self.nnz, _ip(self.row_ptr), _ip(self.cols))


class BlockedSBM(_Resident):              # sparse.h:163-172
    def __init__(self, nrow, ncol, start_row, nnz, rows, cols, vals=None):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.start_row, self.nnz = _i32(start_row), _i32(nnz)
        self.rows, self.cols, self.vals = rows, cols, vals     # lists of per-block arrays

    @property
    def nblocks(self):
        return int(self.start_row.size - 1)

    def _ptrs(self, arrs, typ):
        n = max(self.nblocks, 1)
        out = (typ * n)()
        for b in range(self.nblocks):
            out[b] = arrs[b].ctypes.data_as(typ)
        return out

    def _dev(self):
        if not self._h:
            h = handle()
            rp, cp = self._ptrs(self.rows, c_int_p), self._ptrs(self.cols, c_int_p)
            vp = self._ptrs(self.vals, c_dbl_p) if self.vals is not None else None
            check(lib().fsb_blocked_upload(C.byref(h), self.nrow, self.ncol, self.nblocks, _ip(self.start_row), _ip(self.nnz), rp, cp, vp))
            self._h = h
        return self._h

    def _cached(self):
        # the pointer tables must outlive the call (the cache may hash the blocks on a worker thread until settle)
        self._tables = (self._ptrs(self.rows, c_int_p), self._ptrs(self.cols, c_int_p),
                        self._ptrs(self.vals, c_dbl_p) if self.vals is not None else None)
        return lib().fsb_cache_blocked(self.nrow, self.ncol, self.nblocks, _ip(self.start_row), _ip(self.nnz), *self._tables)


class BlockedSDM(BlockedSBM):             # dsparse.h:119-129
    pass


def _dropin(structs, run):
    """The drop-in call protocol of include/fastsparse/*.h (FSB_DROPIN_CALL): take the handles from the residency cache,
    run, settle; repeat when a handle turned out to be a stale copy of arrays that were edited in place."""
    while True:
        hs = [st._cached() for st in structs]
        if any(not h for h in hs):
            lib().fsb_cache_settle()
            check(lib().fsb_last_error_code() or 1)
        check(run(*hs))
        if lib().fsb_cache_settle() == 0:
            return


# ----------------------------------------------------------------------------- hilbert.h / quickSort*.h
def ceilPower2(x):
    return lib().fsb_host_ceil_pow2(int(x))


def xy2d(n, x, y):
    return lib().fsb_host_xy2d(int(n), int(x), int(y))


def d2xy(n, d):
    a, b = C.c_int(), C.c_int()
    lib().fsb_host_d2xy(int(n), int(d), C.byref(a), C.byref(b))
    return a.value, b.value


def row_xy2d(n, x, y):
    return lib().fsb_host_row_xy2d(int(n), int(x), int(y))


def row_d2xy(n, d):
    a, b = C.c_int(), C.c_int()
    lib().fsb_host_row_d2xy(int(n), int(d), C.byref(a), C.byref(b))
    return a.value, b.value


def quickSort(a):
    """in-place ascending sort of an int64 array (quickSort.h:10-24)"""
    assert a.dtype == np.int64 and a.flags.c_contiguous
    lib().fsb_host_sort_keys(a.ctypes.data_as(c_long_p), None, a.size)


def quickSortD(a, v):
    """in-place sort of int64 keys with a float64 payload (quickSortD.h:12-26)"""
    assert a.dtype == np.int64 and v.dtype == np.float64 and a.size == v.size
    lib().fsb_host_sort_keys(a.ctypes.data_as(c_long_p), _dp(v), a.size)


# ----------------------------------------------------------------------------- sparse.h
def new_sbm(nrow, ncol, nnz, rows, cols):                     # sparse.h:21-29
    assert len(rows) == nnz and len(cols) == nnz
    return SparseBinaryMatrix(nrow, ncol, rows, cols)


def free_sbm(A):                                              # sparse.h:31-34
    A._drop()


def new_transpose(A):                                         # sparse.h:38-46 (aliases the arrays)
    return SparseBinaryMatrix(A.ncol, A.nrow, A.cols, A.rows)


def transpose(A):                                             # sparse.h:48-55
    A._drop()
    A.rows, A.cols = A.cols, A.rows
    A.nrow, A.ncol = A.ncol, A.nrow


def _read_coo(path, with_vals):
    nrow, ncol, nnz = C.c_long(), C.c_long(), C.c_long()
    check(lib().fsb_host_read_coo(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), None, None, None))
    rows = np.zeros(nnz.value, np.int32)
    cols = np.zeros(nnz.value, np.int32)
    vals = np.zeros(nnz.value, np.float64) if with_vals else None
    check(lib().fsb_host_read_coo(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), _ip(rows), _ip(cols), _dp(vals)))
    return nrow.value, ncol.value, rows, cols, vals


def read_sbm(filename):                                       # sparse.h:112-139
    nrow, ncol, rows, cols, _ = _read_coo(filename, False)
    return SparseBinaryMatrix(nrow, ncol, rows, cols)


def sort_sbm(A):                                              # sparse.h:142-161
    A._drop()
    check(lib().fsb_sort_coo_hilbert_auto(A.nrow, A.ncol, A.nnz, _ip(A.rows), _ip(A.cols), None))


def A_mul_B(y, A, x):                                         # sparse.h:58-65 / dsparse.h:43-51
    x = _f64(x)
    assert x.size >= A.ncol
    yo = _out(y, A.nrow)
    _dropin([A], lambda h: lib().fsb_spmm_host(h, _dp(yo), _dp(x), 1))


def At_mul_B(y, A, x):                                        # sparse.h:68-75 / dsparse.h:54-62
    x = _f64(x)
    assert x.size >= A.nrow
    yo = _out(y, A.ncol)
    _dropin([A], lambda h: lib().fsb_spmm_t_host(h, _dp(yo), _dp(x), 1))


def _new_blocked(A, block_size, cls):
    L = lib()
    nb = L.fsb_host_blocked_nblocks(A.nrow, block_size)
    start_row = np.zeros(nb + 1, np.int32)
    cnt = np.zeros(max(nb, 1), np.int32)
    check(L.fsb_host_blocked_count(A.nnz, A.nrow, block_size, _ip(A.rows), _ip(start_row), _ip(cnt)))
    rows = [np.zeros(int(cnt[b]), np.int32) for b in range(nb)]
    cols = [np.zeros(int(cnt[b]), np.int32) for b in range(nb)]
    vals = [np.zeros(int(cnt[b]), np.float64) for b in range(nb)] if A.vals is not None else None
    B = cls(A.nrow, A.ncol, start_row, cnt[:nb], rows, cols, vals)
    check(L.fsb_host_blocked_fill(A.nnz, block_size, _ip(A.rows), _ip(A.cols), _dp(A.vals), nb, B._ptrs(rows, c_int_p),
                                  B._ptrs(cols, c_int_p), B._ptrs(vals, c_dbl_p) if vals is not None else None))
    return B


def new_bsbm(A, block_size):                                  # sparse.h:175-213
    return _new_blocked(A, block_size, BlockedSBM)


def _sort_blocked(B, order, device=None):
    """device=None: the drop-in's choice (device at >= FSB_SORT_DEVICE_MIN entries when a GPU is present); True / False force it."""
    B._drop()
    rp, cp = B._ptrs(B.rows, c_int_p), B._ptrs(B.cols, c_int_p)
    vp = B._ptrs(B.vals, c_dbl_p) if B.vals is not None else None
    if device is False:
        for b in range(B.nblocks):
            if order == 1:
                check(lib().fsb_host_sort_block_hilbert(int(B.start_row[b]), int(B.start_row[b + 1] - B.start_row[b]), int(B.nnz[b]),
                                                        _ip(B.rows[b]), _ip(B.cols[b]), _dp(B.vals[b]) if B.vals is not None else None))
            else:
                check(lib().fsb_host_sort_block_byrow(B.ncol, int(B.nnz[b]), _ip(B.rows[b]), _ip(B.cols[b])))
        return
    fn = lib().fsb_sort_blocked if device else lib().fsb_sort_blocked_auto
    check(fn(B.nrow, B.ncol, B.nblocks, _ip(B.start_row), _ip(B.nnz), rp, cp, vp, order))


def sort_bsbm(B, device=None):                                # sparse.h:215-236 / dsparse.h:193-216
    _sort_blocked(B, 1, device)


def sort_bsbm_byrow(B, device=None):                          # sparse.h:238-256
    _sort_blocked(B, 2, device)


def bsbm_A_mul_Bn(y, B, x, ncol):                             # sparse.h:318-336
    x = _f64(x)
    assert x.size >= B.ncol * ncol
    yo = _out(y, B.nrow * ncol)
    _dropin([B], lambda h: lib().fsb_spmm_host(h, _dp(yo), _dp(x), ncol))


def bsbm_A_mul_B(y, B, x):                                    # sparse.h:259-273
    bsbm_A_mul_Bn(y, B, x, 1)


def bsbm_A_mul_B2(y, B, x):                                   # sparse.h:276-293
    bsbm_A_mul_Bn(y, B, x, 2)


def bsbm_A_mul_B4(y, B, x):                                   # sparse.h:296-315
    bsbm_A_mul_Bn(y, B, x, 4)


# ----------------------------------------------------------------------------- dsparse.h
def new_sdm(nrow, ncol, nnz, rows, cols, vals):               # dsparse.h:22-31
    assert len(rows) == nnz and len(cols) == nnz and len(vals) == nnz
    return SparseDoubleMatrix(nrow, ncol, rows, cols, vals)


sdm_transpose = transpose                                     # dsparse.h:33-40


def read_sdm(filename):                                       # dsparse.h:64-93
    nrow, ncol, rows, cols, vals = _read_coo(filename, True)
    return SparseDoubleMatrix(nrow, ncol, rows, cols, vals)


def sort_sdm(A):                                              # dsparse.h:96-115
    A._drop()
    check(lib().fsb_sort_coo_hilbert_auto(A.nrow, A.ncol, A.nnz, _ip(A.rows), _ip(A.cols), _dp(A.vals)))


sdm_A_mul_B = A_mul_B                                         # dsparse.h:43-51
sdm_At_mul_B = At_mul_B                                       # dsparse.h:54-62


def new_bsdm(A, block_size):                                  # dsparse.h:132-173
    return _new_blocked(A, block_size, BlockedSDM)


sort_bsdm = sort_bsbm                                         # dsparse.h:193-216
bsdm_A_mul_B = bsbm_A_mul_B                                   # dsparse.h:176-191


# ----------------------------------------------------------------------------- csr.h
def new_bcsr(nnz, nrow, ncol, rows, cols):                    # csr.h:30-67
    rows, cols = _i32(rows), _i32(cols)
    row_ptr = np.zeros(nrow + 1, np.int32)
    out_cols = np.zeros(max(nnz, 1), np.int32)
    check(lib().fsb_host_csr_from_coo(nnz, nrow, _ip(rows), _ip(cols), None, _ip(row_ptr), _ip(out_cols), None))
    return BinaryCSR(nrow, ncol, row_ptr, out_cols[:nnz])


def bcsr_from_sbm(sbm):                                       # csr.h:69-74
    return new_bcsr(sbm.nnz, sbm.nrow, sbm.ncol, sbm.rows, sbm.cols)


def free_bcsr(A):                                             # csr.h:24-28
    A._drop()


def new_csr(nnz, nrow, ncol, rows, cols, vals):               # csr.h:375-422
    rows, cols, vals = _i32(rows), _i32(cols), _f64(vals)
    row_ptr = np.zeros(nrow + 1, np.int32)
    out_cols = np.zeros(max(nnz, 1), np.int32)
    out_vals = np.zeros(max(nnz, 1), np.float64)
    check(lib().fsb_host_csr_from_coo(nnz, nrow, _ip(rows), _ip(cols), _dp(vals), _ip(row_ptr), _ip(out_cols), _dp(out_vals)))
    return CSR(nrow, ncol, row_ptr, out_cols[:nnz], out_vals[:nnz])


free_csr = free_bcsr                                          # csr.h:368-373


class _BcsrImage(C.Structure):                                # the raw 32-byte struct BinaryCSR image
    _fields_ = [("nrow", C.c_int), ("ncol", C.c_int), ("nnz", C.c_long), ("row_ptr", C.c_void_p), ("cols", C.c_void_p)]


def serialize_to_file(bcsr, filename):                        # csr.h:97-113
    img = _BcsrImage(bcsr.nrow, bcsr.ncol, bcsr.nnz, bcsr.row_ptr.ctypes.data, bcsr.cols.ctypes.data)
    check(lib().fsb_host_write_csr_bin(filename.encode(), C.byref(img), bcsr.nrow, bcsr.nnz, _ip(bcsr.row_ptr), _ip(bcsr.cols)))


def deserialize_from_file(filename):                          # csr.h:117-146
    img = _BcsrImage()
    check(lib().fsb_host_read_csr_bin(filename.encode(), C.byref(img), None, None))
    row_ptr = np.zeros(img.nrow + 1, np.int32)
    cols = np.zeros(max(img.nnz, 1), np.int32)
    check(lib().fsb_host_read_csr_bin(filename.encode(), C.byref(img), _ip(row_ptr), _ip(cols)))
    return BinaryCSR(img.nrow, img.ncol, row_ptr, cols[:img.nnz])


def bcsr_A_mul_Bn(Y, A, X, ncol):                             # csr.h:257-280 / 441-465
    X = _f64(X)
    assert X.size >= A.ncol * ncol
    Yo = _out(Y, A.nrow * ncol)
    _dropin([A], lambda h: lib().fsb_spmm_host(h, _dp(Yo), _dp(X), ncol))


def bcsr_A_mul_B32n(Y, A, X, ncol):                           # csr.h:283-302
    assert ncol <= 32
    bcsr_A_mul_Bn(Y, A, X, ncol)


def bcsr_A_mul_B(y, A, x):                                    # csr.h:149-161
    bcsr_A_mul_Bn(y, A, x, 1)


def bcsr_A_mul_B2(Y, A, X):                                   # csr.h:164-181
    bcsr_A_mul_Bn(Y, A, X, 2)


def bcsr_A_mul_B4(Y, A, X):                                   # csr.h:184-202
    bcsr_A_mul_Bn(Y, A, X, 4)


def bcsr_A_mul_B8(Y, A, X):                                   # csr.h:205-223
    bcsr_A_mul_Bn(Y, A, X, 8)


bcsr_A_mul_B8_auto = bcsr_A_mul_B8                            # csr.h:225-254


def bcsr_At_mul_Bn(Y, A, X, ncol):
    """Y[ncol_A][ncol] = A' X -- CSR-side transposed product the reference lacks (SURVEY 8b "New")."""
    X = _f64(X)
    assert X.size >= A.nrow * ncol
    Yo = _out(Y, A.ncol * ncol)
    _dropin([A], lambda h: lib().fsb_spmm_t_host(h, _dp(Yo), _dp(X), ncol))


def bcsr_AA_mul_B(y, A, x, mode=0):                           # csr.h:305-319
    x = _f64(x)
    assert x.size >= A.ncol
    yo = _out(y, A.ncol)
    _dropin([A], lambda h: lib().fsb_ata_host(h, _dp(yo), _dp(x), 1, 0.0, mode))


def parallel_bcsr_AA_mul_B(y, A, x, ytmp=None):               # csr.h:323-355 (ytmp: unused scratch of the CPU version)
    bcsr_AA_mul_B(y, A, x, mode=1)


csr_A_mul_Bn = bcsr_A_mul_Bn                                  # csr.h:441-465
csr_A_mul_B = bcsr_A_mul_B                                    # csr.h:425-438
csr_At_mul_Bn = bcsr_At_mul_Bn


# ----------------------------------------------------------------------------- cbcsr.h
def new_cbcsr(colblocksize, nnz, nrow, ncol, rows, cols):     # cbcsr.h:16-65
    rows, cols = _i32(rows), _i32(cols)
    nb = lib().fsb_host_cbcsr_nblocks(ncol, colblocksize)
    row_ptr = np.zeros(nb * nrow + 1, np.int32)
    out_cols = np.zeros(max(nnz, 1), np.int32)
    check(lib().fsb_host_cbcsr_from_coo(colblocksize, nnz, nrow, ncol, _ip(rows), _ip(cols), _ip(row_ptr), _ip(out_cols)))
    return ColBinaryCSR(nrow, ncol, nb, colblocksize, row_ptr, out_cols[:nnz])


def cbcsr_from_sbm(sbm, colblocksize):                        # cbcsr.h:67-73
    return new_cbcsr(colblocksize, sbm.nnz, sbm.nrow, sbm.ncol, sbm.rows, sbm.cols)


def cbcsr_A_mul_Bn(Y, A, X, ncol):
    """n-RHS product on the column-blocked format (the reference has R = 1 only)."""
    X = _f64(X)
    assert X.size >= A.ncol * ncol
    Yo = _out(Y, A.nrow * ncol)
    _dropin([A], lambda h: lib().fsb_spmm_host(h, _dp(Yo), _dp(X), ncol))


def cbcsr_A_mul_B(y, A, x):                                   # cbcsr.h:76-106
    cbcsr_A_mul_Bn(y, A, x, 1)


# ----------------------------------------------------------------------------- linalg.h (reductions run on the GPU)
def _gram(Xa, Xb, n, R):
    G = np.zeros((R, R), np.float64)
    a = _f64(Xa).reshape(-1)
    b = a if Xb is Xa else _f64(Xb).reshape(-1)
    assert a.size >= n * R and b.size >= n * R
    check(lib().fsb_gram_host(_dp(G), _dp(a), _dp(b), n, R))
    return G


def pnormsq(x, n):                                            # linalg.h:15-22
    return float(_gram(x, x, n, 1)[0, 0])


def pdot(x, y, n):                                            # linalg.h:51-58
    return float(_gram(x, y, n, 1)[0, 0])


def pnormsq2(normsq, X, n):                                   # linalg.h:24-34
    G = _gram(X, X, n, 2)
    normsq[0], normsq[1] = G[0, 0], G[1, 1]


def pouter2(outer, X, n):                                     # linalg.h:37-49
    G = _gram(X, X, n, 2)
    outer[0], outer[1], outer[2] = G[0, 0], G[1, 1], G[0, 1]


def pdot2sym(D, X, Y, n):                                     # linalg.h:61-73
    G = _gram(X, Y, n, 2)
    D[0], D[1], D[2] = G[0, 0], G[1, 1], G[0, 1]


def solve2sym(X, A, RHS):                                     # linalg.h:77-88 (scalar host control code)
    dinv = 1.0 / (A[0] * A[1] - A[2] * A[2])
    i00, i11, i01 = dinv * A[1], dinv * A[0], -dinv * A[2]
    X[0] = i00 * RHS[0] + i01 * RHS[1]
    X[1] = i01 * RHS[0] + i11 * RHS[1]
    X[2] = i00 * RHS[2] + i01 * RHS[3]
    X[3] = i01 * RHS[2] + i11 * RHS[3]


def dist(x, y, n):                                            # linalg.h:6-13
    out = C.c_double(0.0)
    check(lib().fsb_dist_host(C.byref(out), _dp(_f64(x)), _dp(_f64(y)), n))
    return out.value


# ----------------------------------------------------------------------------- cg.h
def _check_pair(A, At):
    if A.nrow != At.ncol or A.ncol != At.nrow:      # cg.h:32-36: the reference prints this and exit(1)s
        raise FsbError(1, "A (%d x %d) and At (%d x %d) must be transposes of each other." % (A.nrow, A.ncol, At.nrow, At.ncol))


def bsbm_AtA(y, A, At, x, tmp, lam):                          # cg.h:9-22
    _check_pair(A, At)
    x = _f64(x)
    assert x.size >= A.ncol
    yo = _out(y, A.ncol)
    to = _out(tmp, A.nrow) if tmp is not None else None
    _dropin([A, At], lambda ha, ht: lib().fsb_ata_pair_host(ha, ht, _dp(yo), _dp(x), 1, float(lam), _dp(to) if to is not None else None))


def bsbm_cgn(X, A, At, B, ncol, lam, tol, max_iter=0):
    """Block CG with `ncol` right-hand sides (R <= 32): the generalisation of bsbm_cg2 (SURVEY 8b "New").
    Returns the iteration counter the reference stores in *out_iter."""
    _check_pair(A, At)
    B = _f64(B)
    assert B.size >= A.ncol * ncol
    it = C.c_int(0)
    Xo = _out(X, A.ncol * ncol)
    _dropin([A, At], lambda ha, ht: lib().fsb_cg_host(ha, ht, _dp(Xo), _dp(B), ncol, float(lam), float(tol), int(max_iter), C.byref(it)))
    return it.value


def bsbm_cg(x, A, At, b, lam, tol):                           # cg.h:25-82
    return bsbm_cgn(x, A, At, b, 1, lam, tol)


def bsbm_cg2(X, A, At, B, lam, tol):                          # cg.h:85-187
    return bsbm_cgn(X, A, At, B, 2, lam, tol)


# ----------------------------------------------------------------------------- multi-GPU helpers / synthetic inputs
def partition_rows(row_ptr, nparts):
    """nnz-balanced contiguous row partition (host): bounds[p]..bounds[p+1] = rows of part p."""
    row_ptr = _i32(row_ptr)
    bounds = np.zeros(nparts + 1, np.int32)
    check(lib().fsb_partition_rows(row_ptr.size - 1, _ip(row_ptr), nparts, _ip(bounds)))
    return bounds


def synth_coo_host(seed, dist_kind, nnz, nrow, ncol, with_vals=False):
    rows = np.zeros(nnz, np.int32)
    cols = np.zeros(nnz, np.int32)
    vals = np.zeros(nnz, np.float64) if with_vals else None
    check(lib().fsb_synth_coo_host(seed, dist_kind, nnz, nrow, ncol, _ip(rows), _ip(cols), _dp(vals)))
    return rows, cols, vals


def _torch_stream():
    import torch
    s = torch.cuda.current_stream().cuda_stream
    return s if s else 1          # 0 is the legacy default stream: cudaStreamLegacy == (cudaStream_t)1


class DeviceMatrix:
    """A sparse matrix resident in HBM plus device-pointer products (torch tensors = device memory only)."""

    def __init__(self, h, owner=None):
        self.h = h
        self._owner = owner     # keeps a host structure (and its handle) alive when borrowed
        fmt, nrow, ncol, nnz, hv, nb = C.c_int(), C.c_int(), C.c_int(), C.c_long(), C.c_int(), C.c_int()
        check(lib().fsb_matrix_info(h, C.byref(fmt), C.byref(nrow), C.byref(ncol), C.byref(nnz), C.byref(hv), C.byref(nb)))
        self.format, self.nrow, self.ncol, self.nnz, self.has_vals, self.nblocks = fmt.value, nrow.value, ncol.value, nnz.value, bool(hv.value), nb.value

    @classmethod
    def of(cls, host_struct):
        return cls(host_struct._dev(), owner=host_struct)

    @classmethod
    def from_coo_tensors(cls, nrow, ncol, rows, cols, vals=None):
        h = handle()
        check(lib().fsb_csr_from_coo_dev(C.byref(h), nrow, ncol, rows.numel(), rows.data_ptr(), cols.data_ptr(),
                                         vals.data_ptr() if vals is not None else None))
        return cls(h)

    @classmethod
    def synth(cls, seed, dist_kind, nnz, nrow, ncol, with_vals=False, keep_coo=False):
        import torch
        torch.cuda.synchronize()
        rows = torch.empty(nnz, dtype=torch.int32, device="cuda")
        cols = torch.empty(nnz, dtype=torch.int32, device="cuda")
        vals = torch.empty(nnz, dtype=torch.float64, device="cuda") if with_vals else None
        check(lib().fsb_synth_coo_dev(seed, dist_kind, nnz, nrow, ncol, rows.data_ptr(), cols.data_ptr(),
                                      vals.data_ptr() if with_vals else None, _torch_stream()))
        torch.cuda.synchronize()
        m = cls.from_coo_tensors(nrow, ncol, rows, cols, vals)
        if keep_coo:
            m.coo = (rows, cols, vals)
        return m

    @classmethod
    def blocked_from_coo_tensors(cls, nrow, ncol, rows, cols, vals, block_size, order=1):
        """BlockedSBM/SDM built on the device; order 0 COO, 1 Hilbert (sort_bsbm), 2 by row (sort_bsbm_byrow)."""
        h = handle()
        check(lib().fsb_blocked_from_coo_dev(C.byref(h), nrow, ncol, rows.numel(), rows.data_ptr(), cols.data_ptr(),
                                             vals.data_ptr() if vals is not None else None, block_size, order))
        return cls(h)

    @classmethod
    def cbcsr_from_coo_tensors(cls, nrow, ncol, rows, cols, colblocksize):
        h = handle()
        check(lib().fsb_cbcsr_from_coo_dev(C.byref(h), nrow, ncol, rows.numel(), rows.data_ptr(), cols.data_ptr(), colblocksize))
        return cls(h)

    @classmethod
    def load_coo_file(cls, path, with_vals=False):
        """read_sbm / read_sdm file -> HBM-resident CSR, without a host copy of the matrix."""
        h = handle()
        check(lib().fsb_csr_load_coo_file(C.byref(h), str(path).encode(), int(with_vals)))
        return cls(h)

    @classmethod
    def load_csr_bin(cls, path):
        """.csr.bin (serialize_to_file) -> HBM-resident binary CSR, without a host copy of the matrix."""
        h = handle()
        check(lib().fsb_csr_load_bin_file(C.byref(h), str(path).encode(), None))
        return cls(h)

    def row_slice(self, r0, r1):
        h = handle()
        check(lib().fsb_csr_row_slice(C.byref(h), self.h, int(r0), int(r1)))
        return DeviceMatrix(h)

    def download_csr(self):
        row_ptr = np.zeros(self.nrow + 1, np.int32)
        cols = np.zeros(max(self.nnz, 1), np.int32)
        vals = np.zeros(max(self.nnz, 1), np.float64) if self.has_vals else None
        check(lib().fsb_csr_download(self.h, _ip(row_ptr), _ip(cols), _dp(vals)))
        return row_ptr, cols[: self.nnz], (vals[: self.nnz] if vals is not None else None)

    def set_row_sharded(self, flag=True):
        check(lib().fsb_matrix_set_row_sharded(self.h, int(flag)))

    def bytes(self):
        return lib().fsb_matrix_bytes(self.h)

    def tuning(self, transposed=False):
        """(R, column passes, deep build) chosen by the staged SpMM's per-handle autotune; R = 0 before the first product."""
        R, p, d = C.c_int(), C.c_int(), C.c_int()
        check(lib().fsb_matrix_tuning(self.h, int(transposed), C.byref(R), C.byref(p), C.byref(d)))
        return R.value, p.value, bool(d.value)

    def spmm(self, dX, R, out=None):
        import torch
        if out is None:
            out = torch.empty(self.nrow * R, dtype=torch.float64, device=dX.device)
        check(lib().fsb_spmm_dev(self.h, out.data_ptr(), dX.data_ptr(), R, _torch_stream()))
        return out

    def spmm_t(self, dX, R, out=None):
        import torch
        if out is None:
            out = torch.empty(self.ncol * R, dtype=torch.float64, device=dX.device)
        check(lib().fsb_spmm_t_dev(self.h, out.data_ptr(), dX.data_ptr(), R, _torch_stream()))
        return out

    def ata(self, dX, R, lam=0.0, mode=0, out=None, tmp=None):
        import torch
        if out is None:
            out = torch.empty(self.ncol * R, dtype=torch.float64, device=dX.device)
        check(lib().fsb_ata_dev(self.h, out.data_ptr(), dX.data_ptr(), R, float(lam), tmp.data_ptr() if tmp is not None else None,
                                mode, _torch_stream()))
        return out

    def noise_rhs(self, R, lam, seed, At=None, out=None):
        """B = A'N + sqrt(lam) E with fresh device-generated noise (one Macau sampling step's right-hand side)."""
        import torch
        if out is None:
            out = torch.empty(self.ncol * R, dtype=torch.float64, device="cuda")
        check(lib().fsb_noise_rhs_dev(self.h, At.h if At is not None else None, out.data_ptr(), R, float(lam), int(seed), _torch_stream()))
        return out

    def cg(self, dB, R, lam, tol, At=None, max_iter=0, out=None):
        import torch
        if out is None:
            out = torch.empty(self.ncol * R, dtype=torch.float64, device=dB.device)
        it = C.c_int(0)
        check(lib().fsb_cg_dev(self.h, At.h if At is not None else None, out.data_ptr(), dB.data_ptr(), R, float(lam), float(tol),
                               int(max_iter), C.byref(it), _torch_stream()))
        return out, it.value

    def free(self):
        if self.h and self._owner is None:
            lib().fsb_matrix_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ----------------------------------------------------------------------------- multi-GPU plumbing
def comm_init_from_torch():
    """Create the library's NCCL communicator over the ranks of the current torch.distributed
    process group (one process per GPU): rank 0 makes the unique id, torch broadcasts the bytes."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib().fsb_comm_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    check(lib().fsb_comm_init(world, rank, raw))
    return world, rank


def cg_shard_layout(F, R, G):
    """(C, s, Fc, Fp, nloc): chunk / slice geometry of the multi-GPU block CG's F-sharded vectors (fsb.h)."""
    Cc = C.c_int(); s, Fc, Fp, nloc = C.c_long(), C.c_long(), C.c_long(), C.c_long()
    check(lib().fsb_cg_shard_layout(int(F), int(R), int(G), C.byref(Cc), C.byref(s), C.byref(Fc), C.byref(Fp), C.byref(nloc)))
    return Cc.value, s.value, Fc.value, Fp.value, nloc.value


def comm_finalize():
    check(lib().fsb_comm_finalize())


def allreduce_sum(t):
    """In-place sum-allreduce of a float64 CUDA tensor over the library communicator."""
    check(lib().fsb_allreduce_sum_dev(t.data_ptr(), t.numel(), _torch_stream()))
    return t

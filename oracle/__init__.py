"""CPU oracle bindings -- TEST INFRASTRUCTURE ONLY.

`oracle.O`   : ctypes handle on oracle/_build/libfsoracle.so (our plain-C
               restatement of the reference algorithms, oracle/fsoracle.c).
`oracle.REF` : ctypes handle on oracle/_ref/libfsref_<isa>.so (the UNMODIFIED
               reference headers behind flat-array wrappers, oracle/ref_wrap.c),
               or None when it was never built.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
--impl reference) may import this package.  The product package
libfastsparse_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_build", "libfsoracle.so")

c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)
c_dbl_p = C.POINTER(C.c_double)


class Blocked(C.Structure):
    """fso_blocked (oracle/fsoracle.h): flattened BlockedSBM / BlockedSDM."""

    _fields_ = [
        ("nrow", C.c_int), ("ncol", C.c_int), ("nblocks", C.c_int),
        ("start_row", c_int_p), ("blk_nnz", c_int_p), ("blk_off", c_long_p),
        ("rows", c_int_p), ("cols", c_int_p), ("vals", c_dbl_p),
    ]


class BlockedMatrix:
    """numpy-owning companion of `Blocked`."""

    def __init__(self, nrow, ncol, nnz, block_size, with_vals):
        nb = int(np.ceil(nrow / float(block_size))) if nrow > 0 else 0
        self.block_size = block_size
        self.start_row = np.zeros(nb + 1, np.int32)
        self.blk_nnz = np.zeros(max(nb, 1), np.int32)
        self.blk_off = np.zeros(nb + 1, np.int64)
        self.rows = np.zeros(max(nnz, 1), np.int32)
        self.cols = np.zeros(max(nnz, 1), np.int32)
        self.vals = np.zeros(max(nnz, 1), np.float64) if with_vals else None
        self.nnz = nnz
        self.c = Blocked(nrow, ncol, nb, ip(self.start_row), ip(self.blk_nnz), lp(self.blk_off),
                         ip(self.rows), ip(self.cols), dp(self.vals))

    @property
    def nrow(self): return self.c.nrow
    @property
    def ncol(self): return self.c.ncol
    @property
    def nblocks(self): return self.c.nblocks

    def ref(self):
        return C.byref(self.c)

    def copy(self):
        out = BlockedMatrix(self.c.nrow, self.c.ncol, self.nnz, self.block_size, self.vals is not None)
        out.start_row[:] = self.start_row; out.blk_nnz[:] = self.blk_nnz; out.blk_off[:] = self.blk_off
        out.rows[:] = self.rows; out.cols[:] = self.cols
        if self.vals is not None:
            out.vals[:] = self.vals
        return out


def ip(a):
    return None if a is None else a.ctypes.data_as(c_int_p)


def lp(a):
    return None if a is None else a.ctypes.data_as(c_long_p)


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dbl_p)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(force: bool = False) -> None:
    """Compile oracle/_build (always) and oracle/_ref (when /root/reference exists)."""
    if force or not os.path.exists(_ORACLE_SO) or \
            os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "fsoracle.c")):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def _cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        need = ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")
        return all(n in txt for n in need)
    except OSError:
        return False


def _load_oracle():
    stale = not os.path.exists(_ORACLE_SO) or any(
        os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, f)) for f in ("fsoracle.c", "fsoracle.h"))
    if stale:
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    L = C.CDLL(_ORACLE_SO)
    L.fso_xy2d.restype = C.c_long
    L.fso_row_xy2d.restype = C.c_long
    for f in ("fso_dist", "fso_normsq", "fso_dot"):
        getattr(L, f).restype = C.c_double
    L.fso_xy2d.argtypes = [C.c_int] * 3
    L.fso_row_xy2d.argtypes = [C.c_int] * 3
    L.fso_d2xy.argtypes = [C.c_int, C.c_long, c_int_p, c_int_p]
    L.fso_row_d2xy.argtypes = [C.c_int, C.c_long, c_int_p, c_int_p]
    L.fso_sort_keys.argtypes = [c_long_p, C.c_long]
    L.fso_sort_keys_vals.argtypes = [c_long_p, c_dbl_p, C.c_long]
    L.fso_csr_from_coo.argtypes = [C.c_long, C.c_int, c_int_p, c_int_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p]
    L.fso_cbcsr_from_coo.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
    L.fso_blocked_from_coo.argtypes = [C.c_long, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, C.POINTER(Blocked)]
    L.fso_sort_coo_hilbert.argtypes = [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]
    L.fso_coo_A_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p, c_dbl_p]
    L.fso_coo_At_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p, c_dbl_p]
    L.fso_csr_A_mul_Bn.argtypes = [c_dbl_p, C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, C.c_int]
    L.fso_bcsr_AA_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p]
    L.fso_cbcsr_A_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p]
    L.fso_blocked_A_mul_Bn.argtypes = [c_dbl_p, C.POINTER(Blocked), c_dbl_p, C.c_int]
    L.fso_dist.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.fso_normsq.argtypes = [c_dbl_p, C.c_int]
    L.fso_dot.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.fso_normsq2.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.fso_outer2.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.fso_dot2sym.argtypes = [c_dbl_p, c_dbl_p, c_dbl_p, C.c_int]
    L.fso_solve2sym.argtypes = [c_dbl_p, c_dbl_p, c_dbl_p]
    L.fso_blocked_AtA.argtypes = [c_dbl_p, C.POINTER(Blocked), C.POINTER(Blocked), c_dbl_p, c_dbl_p, C.c_double]
    L.fso_blocked_cg.argtypes = [c_dbl_p, C.POINTER(Blocked), C.POINTER(Blocked), c_dbl_p, C.c_double, C.c_double]
    L.fso_blocked_cg2.argtypes = [c_dbl_p, C.POINTER(Blocked), C.POINTER(Blocked), c_dbl_p, C.c_double, C.c_double]
    L.fso_synth_coo.argtypes = [C.c_ulonglong, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p]
    L.fso_synth_coo.restype = None
    L.fso_read_coo_file.argtypes = [C.c_char_p, c_long_p, c_long_p, c_long_p, c_int_p, c_int_p, c_dbl_p]
    L.fso_write_csr_bin.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p]
    L.fso_read_csr_bin.argtypes = [C.c_char_p, c_int_p, c_int_p, c_long_p, c_int_p, c_int_p]
    return L


def _load_ref():
    isa = "v4" if _cpu_has_avx512() else "v3"
    path = os.path.join(_HERE, "_ref", f"libfsref_{isa}.so")
    if not os.path.exists(path):
        return None, None
    L = C.CDLL(path)
    L.ref_xy2d.restype = C.c_long
    L.ref_row_xy2d.restype = C.c_long
    for f in ("ref_dist", "ref_pnormsq", "ref_pdot"):
        getattr(L, f).restype = C.c_double
    L.ref_xy2d.argtypes = [C.c_int] * 3
    L.ref_row_xy2d.argtypes = [C.c_int] * 3
    L.ref_d2xy.argtypes = [C.c_int, C.c_long, c_int_p, c_int_p]
    L.ref_row_d2xy.argtypes = [C.c_int, C.c_long, c_int_p, c_int_p]
    L.ref_quickSort.argtypes = [c_long_p, C.c_long]
    L.ref_quickSortD.argtypes = [c_long_p, c_dbl_p, C.c_long]
    L.ref_new_bcsr.argtypes = [C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
    L.ref_new_csr.argtypes = [C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p]
    L.ref_new_cbcsr.argtypes = [C.c_int, C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
    L.ref_new_bsbm.argtypes = [C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, C.c_int, C.POINTER(Blocked)]
    L.ref_new_bsdm.argtypes = [C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, C.c_int, C.POINTER(Blocked)]
    L.ref_sort_sbm.argtypes = [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p]
    L.ref_sort_sdm.argtypes = [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]
    L.ref_A_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]
    L.ref_At_mul_B.argtypes = L.ref_A_mul_B.argtypes
    L.ref_sdm_A_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p, c_dbl_p]
    L.ref_sdm_At_mul_B.argtypes = L.ref_sdm_A_mul_B.argtypes
    L.ref_bcsr_mul.argtypes = [C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p, C.c_int]
    L.ref_bcsr_AA_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]
    L.ref_parallel_bcsr_AA_mul_B.argtypes = L.ref_bcsr_AA_mul_B.argtypes
    L.ref_parallel_bcsr_AA_mul_B_scratch.argtypes = L.ref_bcsr_AA_mul_B.argtypes + [c_dbl_p]
    L.ref_csr_mul.argtypes = [C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p, c_dbl_p, C.c_int]
    L.ref_cbcsr_A_mul_B.argtypes = [c_dbl_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]
    L.ref_bsbm_mul.argtypes = [C.c_int, c_dbl_p, C.POINTER(Blocked), c_dbl_p, C.c_int]
    L.ref_bsdm_A_mul_B.argtypes = [c_dbl_p, C.POINTER(Blocked), c_dbl_p]
    L.ref_dist.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.ref_pnormsq.argtypes = [c_dbl_p, C.c_int]
    L.ref_pdot.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.ref_pnormsq2.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.ref_pouter2.argtypes = [c_dbl_p, c_dbl_p, C.c_int]
    L.ref_pdot2sym.argtypes = [c_dbl_p, c_dbl_p, c_dbl_p, C.c_int]
    L.ref_solve2sym.argtypes = [c_dbl_p, c_dbl_p, c_dbl_p]
    L.ref_bsbm_AtA.argtypes = [c_dbl_p, C.POINTER(Blocked), C.POINTER(Blocked), c_dbl_p, c_dbl_p, C.c_double]
    L.ref_bsbm_cg.argtypes = [c_dbl_p, C.POINTER(Blocked), C.POINTER(Blocked), c_dbl_p, C.c_double, C.c_double]
    L.ref_bsbm_cg2.argtypes = L.ref_bsbm_cg.argtypes
    L.ref_read_sbm.argtypes = [C.c_char_p, c_long_p, c_long_p, c_long_p, c_int_p, c_int_p]
    L.ref_read_sdm.argtypes = [C.c_char_p, c_long_p, c_long_p, c_long_p, c_int_p, c_int_p, c_dbl_p]
    L.ref_serialize_to_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p]
    L.ref_deserialize_from_file.argtypes = [C.c_char_p, c_int_p, c_int_p, c_long_p, c_int_p, c_int_p]
    for f in ("ref_sort_bsbm", "ref_sort_bsbm_byrow", "ref_sort_bsdm"):
        getattr(L, f).argtypes = [C.POINTER(Blocked)]
    return L, isa


O = _load_oracle()
REF, REF_ISA = _load_ref()


# ---------------------------------------------------------------------------
# numpy-level conveniences (oracle side)
# ---------------------------------------------------------------------------
def csr_from_coo(nrow, rows, cols, vals=None):
    rows, cols = i32(rows), i32(cols)
    nnz = rows.size
    row_ptr = np.zeros(nrow + 1, np.int32)
    out_cols = np.zeros(max(nnz, 1), np.int32)
    out_vals = np.zeros(max(nnz, 1), np.float64) if vals is not None else None
    O.fso_csr_from_coo(nnz, nrow, ip(rows), ip(cols), dp(f64(vals)) if vals is not None else None,
                       ip(row_ptr), ip(out_cols), dp(out_vals))
    return row_ptr, out_cols[:nnz], (out_vals[:nnz] if vals is not None else None)


def cbcsr_from_coo(nrow, ncol, colblocksize, rows, cols):
    rows, cols = i32(rows), i32(cols)
    nnz = rows.size
    nb = O.fso_cbcsr_nblocks(ncol, colblocksize)
    row_ptr = np.zeros(nb * nrow + 1, np.int32)
    out_cols = np.zeros(max(nnz, 1), np.int32)
    O.fso_cbcsr_from_coo(colblocksize, nnz, nrow, ncol, ip(rows), ip(cols), ip(row_ptr), ip(out_cols))
    return nb, row_ptr, out_cols[:nnz]


def blocked_from_coo(nrow, ncol, block_size, rows, cols, vals=None):
    rows, cols = i32(rows), i32(cols)
    B = BlockedMatrix(nrow, ncol, rows.size, block_size, vals is not None)
    O.fso_blocked_from_coo(rows.size, nrow, ncol, block_size, ip(rows), ip(cols),
                           dp(f64(vals)) if vals is not None else None, B.ref())
    return B


def synth_coo(seed, dist, nnz, nrow, ncol, with_vals=False, j0=0):
    """Entries [j0, j0+nnz) of the benchmark's counter-based COO stream (same bits as the product's generator)."""
    rows = np.empty(max(nnz, 1), np.int32)
    cols = np.empty(max(nnz, 1), np.int32)
    vals = np.empty(max(nnz, 1), np.float64) if with_vals else None
    O.fso_synth_coo(seed, dist, j0, nnz, nrow, ncol, ip(rows), ip(cols), dp(vals))
    return rows[:nnz], cols[:nnz], (vals[:nnz] if with_vals else None)


def csr_mul(nrow, row_ptr, cols, vals, X, R):
    X = f64(X)
    Y = np.zeros(nrow * R, np.float64)
    O.fso_csr_A_mul_Bn(dp(Y), nrow, ip(i32(row_ptr)), ip(i32(cols)),
                       dp(f64(vals)) if vals is not None else None, dp(X), R)
    return Y.reshape(nrow, R) if R > 1 else Y


def coo_mul(nrow, rows, cols, vals, x, transpose=False, ncol=None):
    rows, cols, x = i32(rows), i32(cols), f64(x)
    v = dp(f64(vals)) if vals is not None else None
    if transpose:
        y = np.zeros(ncol, np.float64)
        O.fso_coo_At_mul_B(dp(y), ncol, rows.size, ip(rows), ip(cols), v, dp(x))
    else:
        y = np.zeros(nrow, np.float64)
        O.fso_coo_A_mul_B(dp(y), nrow, rows.size, ip(rows), ip(cols), v, dp(x))
    return y


def blocked_mul(B: BlockedMatrix, X, R):
    X = f64(X)
    Y = np.zeros(B.nrow * R, np.float64)
    O.fso_blocked_A_mul_Bn(dp(Y), B.ref(), dp(X), R)
    return Y.reshape(B.nrow, R) if R > 1 else Y


def read_coo_file(path, with_vals=False):
    nrow, ncol, nnz = C.c_long(), C.c_long(), C.c_long()
    rc = O.fso_read_coo_file(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), None, None, None)
    if rc:
        raise IOError(f"fso_read_coo_file({path}) -> {rc}")
    rows = np.zeros(nnz.value, np.int32)
    cols = np.zeros(nnz.value, np.int32)
    vals = np.zeros(nnz.value, np.float64) if with_vals else None
    rc = O.fso_read_coo_file(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), ip(rows), ip(cols), dp(vals))
    if rc:
        raise IOError(f"fso_read_coo_file({path}) -> {rc}")
    return nrow.value, ncol.value, rows, cols, vals

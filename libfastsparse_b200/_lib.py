"""ctypes binding of lib/libfastsparse_b200.so (signatures = include/fsb.h)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FSB_LIB selects an experiment build (tools/sweep.py); the default is the product library
LIB_PATH = os.environ.get("FSB_LIB") or os.path.join(HERE, "lib", "libfastsparse_b200.so")

c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)
c_dbl_p = C.POINTER(C.c_double)
c_int_pp = C.POINTER(c_int_p)
c_dbl_pp = C.POINTER(c_dbl_p)
handle = C.c_void_p


class FsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfastsparse_b200 error {code}: {msg}")
        self.code = code


_LIB = None

_SIGS = {
    "fsb_version": (C.c_int, []),
    "fsb_last_error": (C.c_char_p, []),
    "fsb_last_error_code": (C.c_int, []),
    "fsb_device_count": (C.c_int, []),
    "fsb_init": (C.c_int, [C.c_int]),
    "fsb_sync": (C.c_int, []),
    "fsb_stream": (C.c_void_p, []),
    "fsb_launch_count": (C.c_long, []),
    "fsb_csr_upload": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_csr_upload_coo": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_csr_from_coo_dev": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fsb_blocked_from_coo_dev": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "fsb_cbcsr_from_coo_dev": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_int]),
    "fsb_sort_coo_hilbert_dev": (C.c_int, [C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fsb_sort_coo_hilbert": (C.c_int, [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_sort_blocked": (C.c_int, [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_pp, c_int_pp, c_dbl_pp, C.c_int]),
    "fsb_sort_blocked_auto": (C.c_int, [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_pp, c_int_pp, c_dbl_pp, C.c_int]),
    "fsb_sort_coo_hilbert_auto": (C.c_int, [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_cbcsr_upload": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p]),
    "fsb_blocked_upload": (C.c_int, [C.POINTER(handle), C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_pp, c_int_pp, c_dbl_pp]),
    "fsb_csr_load_coo_file": (C.c_int, [C.POINTER(handle), C.c_char_p, C.c_int]),
    "fsb_csr_load_bin_file": (C.c_int, [C.POINTER(handle), C.c_char_p, C.c_void_p]),
    "fsb_matrix_free": (C.c_int, [handle]),
    "fsb_matrix_info": (C.c_int, [handle, c_int_p, c_int_p, c_int_p, c_long_p, c_int_p, c_int_p]),
    "fsb_matrix_bytes": (C.c_long, [handle]),
    "fsb_matrix_tuning": (C.c_int, [handle, C.c_int, c_int_p, c_int_p, c_int_p]),
    "fsb_csr_download": (C.c_int, [handle, c_int_p, c_int_p, c_dbl_p]),
    "fsb_csr_row_slice": (C.c_int, [C.POINTER(handle), handle, C.c_int, C.c_int]),
    "fsb_spmm_dev": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fsb_spmm_host": (C.c_int, [handle, c_dbl_p, c_dbl_p, C.c_int]),
    "fsb_spmm_t_dev": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fsb_spmm_t_host": (C.c_int, [handle, c_dbl_p, c_dbl_p, C.c_int]),
    "fsb_ata_dev": (C.c_int, [handle, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p]),
    "fsb_ata_host": (C.c_int, [handle, c_dbl_p, c_dbl_p, C.c_int, C.c_double, C.c_int]),
    "fsb_ata_pair_dev": (C.c_int, [handle, handle, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "fsb_cg_host": (C.c_int, [handle, handle, c_dbl_p, c_dbl_p, C.c_int, C.c_double, C.c_double, C.c_int, c_int_p]),
    "fsb_cg_dev": (C.c_int, [handle, handle, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, c_int_p, C.c_void_p]),
    "fsb_device_malloc": (C.c_void_p, [C.c_size_t]),
    "fsb_device_free": (C.c_int, [C.c_void_p]),
    "fsb_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fsb_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fsb_randn_dev": (C.c_int, [C.c_void_p, C.c_long, C.c_ulonglong, C.c_void_p]),
    "fsb_randn_host": (C.c_int, [c_dbl_p, C.c_long, C.c_ulonglong]),
    "fsb_noise_rhs_dev": (C.c_int, [handle, handle, C.c_void_p, C.c_int, C.c_double, C.c_ulonglong, C.c_void_p]),
    "fsb_gram_dev": (C.c_int, [c_dbl_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]),
    "fsb_gram_host": (C.c_int, [c_dbl_p, c_dbl_p, c_dbl_p, C.c_long, C.c_int]),
    "fsb_rowmix_dev": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_dbl_p, C.c_long, C.c_int, C.c_void_p]),
    "fsb_dist_host": (C.c_int, [c_dbl_p, c_dbl_p, c_dbl_p, C.c_long]),
    "fsb_ata_pair_host": (C.c_int, [handle, handle, c_dbl_p, c_dbl_p, C.c_int, C.c_double, c_dbl_p]),
    "fsb_cg_shard_layout": (C.c_int, [C.c_long, C.c_int, C.c_int, c_int_p, c_long_p, c_long_p, c_long_p, c_long_p]),
    "fsb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "fsb_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "fsb_comm_finalize": (C.c_int, []),
    "fsb_comm_size": (C.c_int, []),
    "fsb_comm_rank": (C.c_int, []),
    "fsb_allreduce_sum_dev": (C.c_int, [C.c_void_p, C.c_long, C.c_void_p]),
    "fsb_matrix_set_row_sharded": (C.c_int, [handle, C.c_int]),
    "fsb_partition_rows": (C.c_int, [C.c_int, c_int_p, C.c_int, c_int_p]),
    "fsb_host_ceil_pow2": (C.c_int, [C.c_int]),
    "fsb_host_xy2d": (C.c_long, [C.c_int, C.c_int, C.c_int]),
    "fsb_host_d2xy": (None, [C.c_int, C.c_long, c_int_p, c_int_p]),
    "fsb_host_row_xy2d": (C.c_long, [C.c_int, C.c_int, C.c_int]),
    "fsb_host_row_d2xy": (None, [C.c_int, C.c_long, c_int_p, c_int_p]),
    "fsb_host_sort_keys": (None, [c_long_p, c_dbl_p, C.c_long]),
    "fsb_host_csr_from_coo": (C.c_int, [C.c_long, C.c_int, c_int_p, c_int_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p]),
    "fsb_host_cbcsr_nblocks": (C.c_int, [C.c_int, C.c_int]),
    "fsb_host_cbcsr_from_coo": (C.c_int, [C.c_int, C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]),
    "fsb_host_blocked_nblocks": (C.c_int, [C.c_int, C.c_int]),
    "fsb_host_blocked_count": (C.c_int, [C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p]),
    "fsb_host_blocked_fill": (C.c_int, [C.c_long, C.c_int, c_int_p, c_int_p, c_dbl_p, C.c_int, c_int_pp, c_int_pp, c_dbl_pp]),
    "fsb_host_sort_coo_hilbert": (C.c_int, [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_host_sort_block_hilbert": (C.c_int, [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_host_sort_block_byrow": (C.c_int, [C.c_int, C.c_long, c_int_p, c_int_p]),
    "fsb_host_read_long": (C.c_long, [C.c_void_p, c_int_p]),
    "fsb_host_read_coo": (C.c_int, [C.c_char_p, c_long_p, c_long_p, c_long_p, c_int_p, c_int_p, c_dbl_p]),
    "fsb_host_write_csr_bin": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_long, c_int_p, c_int_p]),
    "fsb_host_read_csr_bin": (C.c_int, [C.c_char_p, C.c_void_p, c_int_p, c_int_p]),
    "fsb_cache_csr": (handle, [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_cache_coo": (handle, [C.c_int, C.c_int, C.c_long, c_int_p, c_int_p, c_dbl_p]),
    "fsb_cache_cbcsr": (handle, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, c_int_p, c_int_p]),
    "fsb_cache_blocked": (handle, [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_pp, c_int_pp, c_dbl_pp]),
    "fsb_cache_drop": (None, [C.c_void_p]),
    "fsb_cache_clear": (None, []),
    "fsb_cache_settle": (C.c_int, []),
    "fsb_cache_stats": (C.c_int, [c_long_p, c_long_p]),
    "fsb_die": (None, [C.c_char_p]),
    "fsb_tune_csr_spmm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fsb_tune_csr_staged": (C.c_int, [C.c_int]),
    "fsb_tune_formats": (C.c_int, [C.c_int]),
    "fsb_tune": (C.c_int, [C.c_char_p, C.c_int]),
    "fsb_tune_cg_dist": (C.c_int, [C.c_int]),
    "fsb_tune_csr_algo": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "fsb_synth_coo_dev": (C.c_int, [C.c_ulonglong, C.c_int, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fsb_synth_coo_host": (C.c_int, [C.c_ulonglong, C.c_int, C.c_long, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p]),
}

DECLARED = tuple(sorted(_SIGS))


def lib():
    """The loaded C-ABI library.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FsbError(-1, f"{LIB_PATH} is missing: build it with `python -m libfastsparse_b200.build` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int) -> None:
    if rc != 0:
        raise FsbError(rc, lib().fsb_last_error().decode(errors="replace"))

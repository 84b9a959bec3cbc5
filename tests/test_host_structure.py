"""CPU-only checks of the product's host side (no GPU needed):
  * the C-ABI library loads and exports every symbol include/fsb.h declares;
  * compute entry points FAIL LOUDLY without a CUDA device (no CPU fallback);
  * the host constructors / Hilbert maths / sorts / file formats of the drop-in
    (fsb_host_*) reproduce the reference's structure bit for bit (golden fixtures
    generated from the unmodified reference) and agree with the oracle on random input."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import libfastsparse_b200 as fs
import oracle
from conftest import DATA, GOLDEN, ROOT, golden
from libfastsparse_b200._lib import DECLARED, LIB_PATH, lib


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fsb.h")).read()
    declared = set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = C.CDLL(LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in include/fsb.h but not exported: {missing}"
    assert declared == set(DECLARED), f"python binding out of sync: {declared ^ set(DECLARED)}"
    assert lib().fsb_version() >= 100


def test_no_cpu_fallback_without_device(have_gpu):
    if have_gpu:
        pytest.skip("a GPU is present")
    A = fs.read_sbm(os.path.join(DATA, "sbm-100-50.data"))
    B = fs.bcsr_from_sbm(A)
    y = np.zeros(B.nrow)
    with pytest.raises(fs.FsbError) as ei:
        fs.bcsr_A_mul_B(y, B, np.ones(B.ncol))
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)          # FSB_ENODEV
    with pytest.raises(fs.FsbError):
        fs.A_mul_B(y, A, np.ones(A.ncol))


def _blocked_eq(B, g, prefix):
    assert np.array_equal(B.start_row, g[prefix + "start_row"]) and np.array_equal(B.nnz, g[prefix + "blk_nnz"])
    cat = lambda arrs, dt: np.concatenate(arrs) if arrs else np.zeros(0, dt)
    assert np.array_equal(cat(B.rows, np.int32), g[prefix + "rows"]) and np.array_equal(cat(B.cols, np.int32), g[prefix + "cols"])
    if B.vals is not None:
        assert np.array_equal(cat(B.vals, np.float64), g[prefix + "vals"])


@pytest.mark.parametrize("name", ["sbm_100_50", "rand_bin_300_70"])
def test_binary_structure_matches_reference(name, tmp_path):
    g = golden(name)
    nrow, ncol, rows, cols = int(g["nrow"]), int(g["ncol"]), g["rows"], g["cols"]
    A = fs.new_sbm(nrow, ncol, rows.size, rows.copy(), cols.copy())
    B = fs.bcsr_from_sbm(A)
    assert np.array_equal(B.row_ptr, g["csr_row_ptr"]) and np.array_equal(B.cols, g["csr_cols"])
    Cb = fs.cbcsr_from_sbm(A, int(g["colblock"]))
    assert Cb.nblocks == int(g["cb_nblocks"]) and np.array_equal(Cb.row_ptr, g["cb_row_ptr"]) and np.array_equal(Cb.cols, g["cb_cols"])
    bs = int(g["bs"])
    Bl = fs.new_bsbm(A, bs); _blocked_eq(Bl, g, "blk_")
    fs.sort_bsbm(Bl); _blocked_eq(Bl, g, "blkh_")
    Br = fs.new_bsbm(A, bs); fs.sort_bsbm_byrow(Br); _blocked_eq(Br, g, "blkr_")
    fs.sort_sbm(A)
    assert np.array_equal(A.rows, g["hil_rows"]) and np.array_equal(A.cols, g["hil_cols"])
    # .csr.bin: same bytes as the reference's writer except the 16 stale-pointer bytes, and round trip
    p = str(tmp_path / "m.csr.bin")
    fs.serialize_to_file(B, p)
    mine = np.frombuffer(open(p, "rb").read(), np.uint8); ref = g["csr_bin"]
    keep = np.ones(ref.size, bool); keep[83:99] = False
    assert mine.size == ref.size and np.array_equal(mine[keep], ref[keep])
    open(p, "wb").write(ref.tobytes())
    B2 = fs.deserialize_from_file(p)
    assert (B2.nrow, B2.ncol, B2.nnz) == (nrow, ncol, rows.size)
    assert np.array_equal(B2.row_ptr, B.row_ptr) and np.array_equal(B2.cols, B.cols)


@pytest.mark.parametrize("name", ["sdm_100_50", "rand_dbl_257_129"])
def test_double_structure_matches_reference(name):
    g = golden(name)
    nrow, ncol, rows, cols, vals = int(g["nrow"]), int(g["ncol"]), g["rows"], g["cols"], g["vals"]
    A = fs.new_sdm(nrow, ncol, rows.size, rows.copy(), cols.copy(), vals.copy())
    M = fs.new_csr(A.nnz, nrow, ncol, A.rows, A.cols, A.vals)
    assert np.array_equal(M.row_ptr, g["csr_row_ptr"]) and np.array_equal(M.cols, g["csr_cols"]) and np.array_equal(M.vals, g["csr_vals"])
    Bl = fs.new_bsdm(A, int(g["bs"])); _blocked_eq(Bl, g, "blk_")
    fs.sort_bsdm(Bl); _blocked_eq(Bl, g, "blkh_")
    fs.sort_sdm(A)
    assert np.array_equal(A.rows, g["hil_rows"]) and np.array_equal(A.cols, g["hil_cols"]) and np.array_equal(A.vals, g["hil_vals"])


def test_loaders_match_reference_fixtures():     # test_sparse.c:185-193, 470-482
    A = fs.read_sbm(os.path.join(DATA, "sbm-100-50.data"))
    assert (A.nrow, A.ncol, A.nnz, A.rows[0], A.cols[0]) == (100, 50, 504, 8, 0)
    D = fs.read_sdm(os.path.join(DATA, "sdm-100-50.data"))
    assert (D.nrow, D.ncol, D.nnz, D.rows[1], D.cols[1], D.rows[469], D.cols[469]) == (100, 50, 470, 27, 0, 40, 49)
    assert abs(D.vals[1] - 0.616153) < 1e-5 and abs(D.vals[469] - 0.108172) < 1e-5
    with pytest.raises(fs.FsbError) as ei:
        fs.read_sbm("/nonexistent/file.data")
    assert ei.value.code == 5


def test_hilbert_and_sort_match_reference():     # test_sparse.c:195-265, 347-361 + golden
    for x, want in [(16, 16), (15, 16), (17, 32), (1, 1), (1 << 30, 1 << 30), ((1 << 30) - 1, 1 << 30)]:
        assert fs.ceilPower2(x) == want
    assert fs.d2xy(131072, fs.xy2d(131072, 5931, 91204)) == (5931, 91204)
    assert [fs.row_xy2d(16, *p) for p in [(0, 0), (0, 15), (0, 16), (0, 31), (1, 0)]] == [0, 255, 256, 511, 3]
    assert [fs.row_d2xy(16, d) for d in (0, 255, 256, 511, 3)] == [(0, 0), (0, 15), (0, 16), (0, 31), (1, 0)]
    g = golden("hilbert")
    assert [fs.ceilPower2(int(v)) for v in g["cp2_in"]] == list(g["cp2_out"])
    for n in (1, 2, 16, 128, 131072, 1 << 24):
        x, y, d = g[f"xy_{n}_x"], g[f"xy_{n}_y"], g[f"xy_{n}_d"]
        for j in range(x.size):
            assert fs.xy2d(n, x[j], y[j]) == d[j]
            assert fs.d2xy(n, d[j]) == (g[f"xy_{n}_bx"][j], g[f"xy_{n}_by"][j])
            assert fs.row_xy2d(n, x[j], g[f"rxy_{n}_y"][j]) == g[f"rxy_{n}_d"][j]
            assert fs.row_d2xy(n, g[f"rxy_{n}_d"][j]) == (g[f"rxy_{n}_bx"][j], g[f"rxy_{n}_by"][j])
    k, p = g["qs_keys"].copy(), g["qs_pay"].copy(); fs.quickSortD(k, p)
    assert np.array_equal(k, g["qs_keys_out"]) and np.array_equal(p, g["qs_pay_out"])
    a = np.array([7, 12, 1, -2, 0, 15, 4, 9, 11, 3, -1, 13, 5], np.int64); fs.quickSort(a)   # test_sparse.c:205-215
    assert np.all(np.diff(a) >= 0)
    a = (2000 * np.sin(np.arange(1000) * 17)).astype(np.int64); fs.quickSort(a)                 # test_sparse.c:217-233
    assert np.all(np.diff(a) >= 0)


@pytest.mark.parametrize("seed", range(4))
def test_host_constructors_vs_oracle_random(seed):
    rng = np.random.default_rng(100 + seed)
    nrow, ncol, nnz = int(rng.integers(1, 500)), int(rng.integers(1, 500)), int(rng.integers(0, 8000))
    rows = rng.integers(0, nrow, nnz, dtype=np.int32); cols = rng.integers(0, ncol, nnz, dtype=np.int32); vals = rng.random(nnz)
    M = fs.new_csr(nnz, nrow, ncol, rows, cols, vals)
    rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, vals)
    assert np.array_equal(M.row_ptr, rp) and np.array_equal(M.cols, cc) and np.array_equal(M.vals, vv)
    cbs = int(rng.integers(1, ncol + 1))
    Cb = fs.new_cbcsr(cbs, nnz, nrow, ncol, rows, cols)
    nb, crp, ccc = oracle.cbcsr_from_coo(nrow, ncol, cbs, rows, cols)
    assert Cb.nblocks == nb and np.array_equal(Cb.row_ptr, crp) and np.array_equal(Cb.cols, ccc)
    bs = int(rng.integers(1, nrow + 1))
    A = fs.new_sdm(nrow, ncol, nnz, rows.copy(), cols.copy(), vals.copy())
    Bl = fs.new_bsdm(A, bs); fs.sort_bsdm(Bl)
    Bo = oracle.blocked_from_coo(nrow, ncol, bs, rows, cols, vals); oracle.O.fso_sort_blocked_hilbert(Bo.ref())
    assert np.array_equal(np.concatenate(Bl.rows) if Bl.rows else [], Bo.rows[:nnz])
    assert np.array_equal(np.concatenate(Bl.cols) if Bl.cols else [], Bo.cols[:nnz])
    assert np.array_equal(np.concatenate(Bl.vals) if Bl.vals else [], Bo.vals[:nnz])


def test_partition_rows_balances_nnz():
    rng = np.random.default_rng(7)
    deg = rng.poisson(20, 10000); deg[100:200] = 0; deg[5000] = 40000
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    for parts in (1, 2, 3, 8):
        b = fs.partition_rows(row_ptr, parts)
        assert b[0] == 0 and b[-1] == 10000 and np.all(np.diff(b) >= 0)
        per = np.diff(row_ptr[b])
        assert per.sum() == row_ptr[-1] and per.max() <= row_ptr[-1] / parts + 40000 + 64
    assert list(fs.partition_rows(np.zeros(11, np.int32), 2)) == [0, 5, 10]


def test_synth_generator_host_is_deterministic_and_in_range():
    r1, c1, v1 = fs.synth_coo_host(0x5EED0002, 0, 5000, 1000, 300, with_vals=True)
    r2, c2, v2 = fs.synth_coo_host(0x5EED0002, 0, 5000, 1000, 300, with_vals=True)
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.array_equal(v1, v2)
    assert r1.min() >= 0 and r1.max() < 1000 and c1.min() >= 0 and c1.max() < 300 and 0 <= v1.min() and v1.max() < 1
    assert abs(np.bincount(r1, minlength=1000).mean() - 5.0) < 1e-9
    _, cz, _ = fs.synth_coo_host(0x5EED0004, 1, 200000, 1000, 4096)
    cnt = np.sort(np.bincount(cz, minlength=4096))[::-1]
    assert cz.min() >= 0 and cz.max() < 4096 and cnt[0] > 50 * max(1, cnt[2048])   # heavy head: power law


def test_read_long_matches_reference_loader(tmp_path):     # utils.h:4-12
    """read_long: one native 8-byte integer per call; a short read reports the reference's message."""
    import ctypes as C
    import numpy as np
    L = fs.lib()
    path = tmp_path / "longs.bin"
    np.array([100, 50, 504], dtype=np.int64).tofile(path)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]; libc.fclose.argtypes = [C.c_void_p]
    f = libc.fopen(str(path).encode(), b"rb")
    ok = C.c_int(0)
    got = [L.fsb_host_read_long(f, C.byref(ok)) for _ in range(3)]
    assert got == [100, 50, 504] and ok.value == 1
    L.fsb_host_read_long(f, C.byref(ok))                    # past the end
    assert ok.value == 0 and b"File is corrupt" in L.fsb_last_error()
    libc.fclose(f)

"""GPU parity tests (run with -m gpu on a B200): every product / solve goes through the
C-ABI library and is compared with the CPU oracle (oracle/fsoracle.c, pinned against the
reference by tests/test_oracle_golden.py) and with the golden outputs of the unmodified
reference.  Index/structure results must be bit-exact; fp64 products must agree to
1e-12 * max(1, sum|terms|) (summation order differs)."""
import ctypes as C
import os

import numpy as np
import pytest

import libfastsparse_b200 as fs
import oracle
from conftest import DATA, assert_close, golden, rhs_matrix, test_vec as tvec
from oracle import O, dp, f64, i32, ip

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    assert fs.device_count() >= 1
    before = fs.launch_count()
    yield
    assert fs.launch_count() > before, "no kernel of libfastsparse_b200.so was launched"


def row_scale(nrow, rows, xmax=2.0, vmax=1.0):
    return xmax * vmax * max(1, int(np.bincount(rows, minlength=nrow).max()) if rows.size else 1)


# ------------------------------------------------------------------ reference known answers
def test_known_answers_tiny():     # test_sparse.c:30-44, 114-135, 159-173, 412-468
    A = fs.new_sbm(4, 3, 5, [0, 3, 3, 1, 2], [0, 2, 0, 2, 1])
    x = np.array([0.5, -0.7, 1.9]); y = np.zeros(4)
    fs.A_mul_B(y, A, x); assert list(y) == [0.5, 1.9, -0.7, 2.4]
    Cb = fs.cbcsr_from_sbm(A, 2); y[:] = 0
    fs.cbcsr_A_mul_B(y, Cb, x); assert list(y) == [0.5, 1.9, -0.7, 2.4]
    z = np.zeros(3); fs.At_mul_B(z, A, np.array([0.2, 1.3, -0.7, -0.5])); assert np.max(np.abs(z - [-0.3, -0.7, 0.8])) < 1e-15
    D = fs.new_sdm(6, 4, 11, [1, 1, 3, 4, 1, 4, 5, 0, 1, 2, 4], [0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3],
                   [0.65, 0.84, 0.54, 0.59, 0.51, 0.27, 0.23, 0.94, 0.66, 0.31, 0.92])
    yt = np.array([2.162, 2.224, 0.713, -0.378, 2.216, 0.437]); y6 = np.zeros(6)
    fs.sdm_A_mul_B(y6, D, np.array([0.5, -0.7, 1.9, 2.3])); assert np.max(np.abs(y6 - yt)) < 1e-6
    M = fs.new_csr(D.nnz, 6, 4, D.rows, D.cols, D.vals)
    fs.csr_A_mul_B(y6, M, np.array([0.5, -0.7, 1.9, 2.3])); assert np.max(np.abs(y6 - yt)) < 1e-6
    Y = np.zeros(12); fs.csr_A_mul_Bn(Y, M, np.array([0.5, 5.0, -0.7, -7.0, 1.9, 19.0, 2.3, 23.0]), 2)
    assert np.max(np.abs(Y.reshape(6, 2) - np.stack([yt, 10 * yt], 1))) < 1e-6
    z4 = np.zeros(4); fs.sdm_At_mul_B(z4, D, np.array([0.59, 0.37, 0.14, 0.21, 0.40, 0.81]))
    assert np.max(np.abs(z4 - [0.2405, 0.6602, 0.483, 1.2102])) < 1e-6


def test_known_answers_fixture():  # test_sparse.c:46-112, 137-157
    A = fs.read_sbm(os.path.join(DATA, "sbm-100-50.data"))
    B = fs.bcsr_from_sbm(A)
    x = tvec(A.ncol); y = np.zeros(A.nrow); y2 = np.zeros(A.nrow)
    fs.A_mul_B(y2, A, x); fs.bcsr_A_mul_B(y, B, x)
    assert abs(y[0] - 1.70095) < 1e-4 and abs(y[99] + 0.174905) < 1e-4 and np.max(np.abs(y - y2)) < 1e-12
    assert abs(y[0] - 1.7009531873605335) < 1e-13 and abs(y.sum() - 42.631240828393558) < 1e-11
    tmp = np.zeros(A.nrow); z2 = np.zeros(A.ncol); z = np.zeros(A.ncol)
    fs.A_mul_B(tmp, A, x); fs.At_mul_B(z2, A, tmp)
    fs.bcsr_AA_mul_B(z, B, x); assert np.max(np.abs(z - z2)) < 1e-10
    assert abs(z[0] - 28.810541791856551) < 1e-11 and abs(z[49] - 16.518731587972916) < 1e-11
    fs.parallel_bcsr_AA_mul_B(z, B, x, None); assert np.max(np.abs(z - z2)) < 1e-10
    Cb = fs.cbcsr_from_sbm(A, 8); fs.cbcsr_A_mul_B(y, Cb, x); assert np.max(np.abs(y - y2)) < 1e-12


# ------------------------------------------------------------------ golden fixtures of the unmodified reference
@pytest.mark.parametrize("name", ["sbm_100_50", "rand_bin_300_70"])
def test_binary_products_vs_reference_golden(name):
    g = golden(name)
    nrow, ncol, rows, cols = int(g["nrow"]), int(g["ncol"]), g["rows"], g["cols"]
    x, xt = g["x"], g["xt"]
    sc = row_scale(nrow, rows); sct = row_scale(ncol, cols)
    A = fs.new_sbm(nrow, ncol, rows.size, rows.copy(), cols.copy())
    y = np.zeros(nrow); fs.A_mul_B(y, A, x); assert_close(y, g["coo_Ax"], sc, what="A_mul_B")
    z = np.zeros(ncol); fs.At_mul_B(z, A, xt); assert_close(z, g["coo_Atx"], sct, what="At_mul_B")
    B = fs.bcsr_from_sbm(A)
    fs.bcsr_A_mul_B(y, B, x); assert_close(y, g["csr_Ax"], sc, what="bcsr_A_mul_B")
    for R in map(int, g["Rs"]):
        Y = np.zeros(nrow * R); fs.bcsr_A_mul_Bn(Y, B, g[f"X{R}"], R); assert_close(Y, g[f"csr_AX{R}_Bn"], sc, what=f"bcsr_A_mul_Bn R={R}")
        if R <= 32:
            Y[:] = 0; fs.bcsr_A_mul_B32n(Y, B, g[f"X{R}"], R); assert_close(Y, g[f"csr_AX{R}_B32n"], sc, what=f"B32n R={R}")
        fixed = {2: fs.bcsr_A_mul_B2, 4: fs.bcsr_A_mul_B4, 8: fs.bcsr_A_mul_B8}.get(R)
        if fixed:
            Y[:] = 0; fixed(Y, B, g[f"X{R}"]); assert_close(Y, g[f"csr_AX{R}_fixed"], sc, what=f"fixed R={R}")
        if R == 8:
            Y[:] = 0; fs.bcsr_A_mul_B8_auto(Y, B, g["X8"]); assert_close(Y, g["csr_AX8_auto"], sc, what="B8_auto")
        # transposed CSR product (new entry point): oracle = At_mul_B column by column
        Xt = rhs_matrix(nrow, R); Z = np.zeros(ncol * R); fs.bcsr_At_mul_Bn(Z, B, Xt, R)
        want = np.stack([oracle.coo_mul(nrow, rows, cols, None, Xt[:, k], transpose=True, ncol=ncol) for k in range(R)], 1)
        assert_close(Z, want, sct, what=f"bcsr_At_mul_Bn R={R}")
    for mode in (0, 1):
        fs.bcsr_AA_mul_B(z, B, x, mode=mode); assert_close(z, g["csr_AAx"], sc * sct, what=f"AA mode {mode}")
    Cb = fs.cbcsr_from_sbm(A, int(g["colblock"]))
    fs.cbcsr_A_mul_B(y, Cb, x); assert_close(y, g["cb_Ax"], sc, what="cbcsr_A_mul_B")
    for R in map(int, g["Rs"]):
        Y = np.zeros(nrow * R); fs.cbcsr_A_mul_Bn(Y, Cb, g[f"X{R}"], R); assert_close(Y, g[f"csr_AX{R}_Bn"], sc, what=f"cbcsr_A_mul_Bn R={R}")
    bs = int(g["bs"])
    for sorter in (None, fs.sort_bsbm, fs.sort_bsbm_byrow):
        Bl = fs.new_bsbm(A, bs)
        if sorter:
            sorter(Bl)
        fs.bsbm_A_mul_B(y, Bl, x); assert_close(y, g["blkh_Ax"], sc, what="bsbm_A_mul_B")
        for R in map(int, g["Rs"]):
            Y = np.zeros(nrow * R); fs.bsbm_A_mul_Bn(Y, Bl, g[f"X{R}"], R); assert_close(Y, g[f"blkh_AX{R}_Bn"], sc, what=f"bsbm_A_mul_Bn R={R}")
        Y = np.zeros(nrow * 2); fs.bsbm_A_mul_B2(Y, Bl, g["X2"]); assert_close(Y, g["blkh_AX2_fixed"], sc, what="bsbm_B2")
        Y = np.zeros(nrow * 4); fs.bsbm_A_mul_B4(Y, Bl, g["X4"]); assert_close(Y, g["blkh_AX4_fixed"], sc, what="bsbm_B4")


@pytest.mark.parametrize("name", ["sdm_100_50", "rand_dbl_257_129"])
def test_double_products_vs_reference_golden(name):
    g = golden(name)
    nrow, ncol, rows, cols, vals = int(g["nrow"]), int(g["ncol"]), g["rows"], g["cols"], g["vals"]
    x, xt = g["x"], g["xt"]
    sc = row_scale(nrow, rows); sct = row_scale(ncol, cols)
    A = fs.new_sdm(nrow, ncol, rows.size, rows.copy(), cols.copy(), vals.copy())
    y = np.zeros(nrow); fs.sdm_A_mul_B(y, A, x); assert_close(y, g["coo_Ax"], sc, what="sdm_A_mul_B")
    z = np.zeros(ncol); fs.sdm_At_mul_B(z, A, xt); assert_close(z, g["coo_Atx"], sct, what="sdm_At_mul_B")
    M = fs.new_csr(A.nnz, nrow, ncol, A.rows, A.cols, A.vals)
    fs.csr_A_mul_B(y, M, x); assert_close(y, g["csr_Ax"], sc, what="csr_A_mul_B")
    for R in map(int, g["Rs"]):
        Y = np.zeros(nrow * R); fs.csr_A_mul_Bn(Y, M, g[f"X{R}"], R); assert_close(Y, g[f"csr_AX{R}_Bn"], sc, what=f"csr_A_mul_Bn R={R}")
    Bl = fs.new_bsdm(A, int(g["bs"]))
    fs.bsdm_A_mul_B(y, Bl, x); assert_close(y, g["blkh_Ax"], sc, what="bsdm_A_mul_B unsorted")
    fs.sort_bsdm(Bl)
    fs.bsdm_A_mul_B(y, Bl, x); assert_close(y, g["blkh_Ax"], sc, what="bsdm_A_mul_B sorted")


# ------------------------------------------------------------------ ragged / edge cases against the oracle
def _random_case(rng, nrow, ncol, nnz, long_row=0):
    rows = rng.integers(0, nrow, nnz, dtype=np.int32); cols = rng.integers(0, ncol, nnz, dtype=np.int32)
    if long_row:
        rows[:long_row] = nrow // 2          # one very long row (and, transposed, one very hot column)
    dead = rng.integers(0, nrow, max(1, nrow // 5))
    rows[np.isin(rows, dead)] = (dead[0] + 1) % nrow     # empty rows
    return rows, cols, rng.random(nnz)


@pytest.fixture(params=[1, 2, 3], ids=["team-kernel", "staged-kernel", "stream-kernel"])
def csr_algo(request):
    """Run a test once per CSR SpMM kernel (fsb_tune_csr_algo), then restore the automatic choice."""
    fs.check(fs.lib().fsb_tune_csr_algo(request.param, 0, 0))
    yield request.param
    fs.check(fs.lib().fsb_tune_csr_algo(0, 0, 0))


def test_staged_kernel_is_bit_identical_to_serial_reference_order():
    """The staged kernel sums each row in stored order: binary products must equal the oracle's
    in-order sums exactly, including rows long enough to take the split path."""
    rng = np.random.default_rng(42)
    nrow, ncol, nnz = 5000, 900, 60000
    rows = rng.integers(0, nrow, nnz, dtype=np.int32); cols = rng.integers(0, ncol, nnz, dtype=np.int32)
    rp, cc, _ = oracle.csr_from_coo(nrow, rows, cols)
    M = fs.BinaryCSR(nrow, ncol, rp, cc)
    fs.check(fs.lib().fsb_tune_csr_algo(2, 0, 0))
    try:
        for R in (1, 4, 32):
            X = f64(rng.standard_normal((ncol, R)))
            Y = np.zeros(nrow * R); fs.bcsr_A_mul_Bn(Y, M, X, R)
            assert np.array_equal(Y, oracle.csr_mul(nrow, rp, cc, None, X, R).reshape(-1)), f"R={R}"
    finally:
        fs.check(fs.lib().fsb_tune_csr_algo(0, 0, 0))


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 8, 16, 17, 31, 32, 33, 64, 100])
def test_csr_products_ragged(R, csr_algo):
    rng = np.random.default_rng(R)
    nrow, ncol, nnz = 3001, 777, 40000
    rows, cols, vals = _random_case(rng, nrow, ncol, nnz, long_row=5000)
    X = f64(rng.standard_normal((ncol, R)))
    for v in (None, vals):
        rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, v)
        M = fs.BinaryCSR(nrow, ncol, rp, cc, vv)
        Y = np.zeros(nrow * R); fs.bcsr_A_mul_Bn(Y, M, X, R)
        sc = np.abs(oracle.csr_mul(nrow, rp, cc, np.abs(vv) if vv is not None else None, np.abs(X), R)).reshape(-1)
        assert_close(Y, oracle.csr_mul(nrow, rp, cc, vv, X, R), sc, what=f"csr R={R} vals={v is not None}")
        Xt = f64(rng.standard_normal((nrow, R))); Z = np.zeros(ncol * R); fs.bcsr_At_mul_Bn(Z, M, Xt, R)
        trp, tcc, tvv = oracle.csr_from_coo(ncol, cols, rows, v)
        sct = np.abs(oracle.csr_mul(ncol, trp, tcc, np.abs(tvv) if tvv is not None else None, np.abs(Xt), R)).reshape(-1)
        assert_close(Z, oracle.csr_mul(ncol, trp, tcc, tvv, Xt, R), sct, what=f"csr^T R={R} vals={v is not None}")


@pytest.fixture(params=[0, 1], ids=["csr-view", "native-format-kernel"])
def format_mode(request):
    """Blocked / column-blocked products: default CSR view vs the format's own kernel (fsb_tune_formats)."""
    fs.check(fs.lib().fsb_tune_formats(request.param))
    yield request.param
    fs.check(fs.lib().fsb_tune_formats(0))


@pytest.mark.parametrize("R", [1, 2, 3, 8, 32, 40])
@pytest.mark.parametrize("bs", [1, 7, 64, 500, 5000])
def test_blocked_products_ragged(R, bs, format_mode):
    rng = np.random.default_rng(R * 1000 + bs)
    nrow, ncol, nnz = 2000, 333, 15000
    rows, cols, vals = _random_case(rng, nrow, ncol, nnz, long_row=600)
    X = f64(rng.standard_normal((ncol, R)))
    for v in (None, vals):
        A = fs.SparseBinaryMatrix(nrow, ncol, rows, cols) if v is None else fs.SparseDoubleMatrix(nrow, ncol, rows, cols, v)
        Bl = fs.new_bsbm(A, bs); fs.sort_bsbm(Bl)
        Bo = oracle.blocked_from_coo(nrow, ncol, bs, rows, cols, v); O.fso_sort_blocked_hilbert(Bo.ref())
        want = oracle.blocked_mul(Bo, X, R)
        Y = np.zeros(nrow * R); fs.bsbm_A_mul_Bn(Y, Bl, X, R)
        rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, v)
        sc = np.abs(oracle.csr_mul(nrow, rp, cc, vv, np.abs(X), R)).reshape(-1)
        assert_close(Y, want, sc, what=f"blocked R={R} bs={bs} vals={v is not None}")


@pytest.mark.parametrize("R", [1, 4, 32])
@pytest.mark.parametrize("colblock", [1, 5, 64, 1000])
def test_cbcsr_products_ragged(R, colblock, format_mode):
    rng = np.random.default_rng(R * 7 + colblock)
    nrow, ncol, nnz = 1500, 640, 20000
    rows, cols, _ = _random_case(rng, nrow, ncol, nnz, long_row=3000)
    X = f64(rng.standard_normal((ncol, R)))
    Cb = fs.new_cbcsr(colblock, nnz, nrow, ncol, rows, cols)
    Y = np.zeros(nrow * R); fs.cbcsr_A_mul_Bn(Y, Cb, X, R)
    rp, cc, _ = oracle.csr_from_coo(nrow, rows, cols)
    sc = np.abs(oracle.csr_mul(nrow, rp, cc, None, np.abs(X), R)).reshape(-1)
    assert_close(Y, oracle.csr_mul(nrow, rp, cc, None, X, R), sc, what=f"cbcsr R={R} colblock={colblock}")
    if R == 1:   # exact block-major association of the reference
        nb, crp, ccc = oracle.cbcsr_from_coo(nrow, ncol, colblock, rows, cols)
        y = np.zeros(nrow); O.fso_cbcsr_A_mul_B(dp(y), nrow, nb, ip(crp), ip(ccc), dp(f64(X[:, 0])))
        assert_close(Y, y, sc, what="cbcsr vs fso_cbcsr_A_mul_B")


def test_empty_and_degenerate_shapes():
    for nrow, ncol in [(1, 1), (5, 3), (64, 1)]:
        rp = np.zeros(nrow + 1, np.int32)
        M = fs.BinaryCSR(nrow, ncol, rp, np.zeros(0, np.int32))
        Y = np.full(nrow * 3, 7.0); fs.bcsr_A_mul_Bn(Y, M, np.ones(ncol * 3), 3)
        assert not Y.any()                                           # outputs are fully overwritten
        Z = np.full(ncol * 3, 7.0); fs.bcsr_At_mul_Bn(Z, M, np.ones(nrow * 3), 3); assert not Z.any()
        A = fs.SparseBinaryMatrix(nrow, ncol, np.zeros(0, np.int32), np.zeros(0, np.int32))
        Bl = fs.new_bsbm(A, 4); y = np.full(nrow, 7.0); fs.bsbm_A_mul_B(y, Bl, np.ones(ncol)); assert not y.any()
        Cb = fs.cbcsr_from_sbm(A, 2); y[:] = 7.0; fs.cbcsr_A_mul_B(y, Cb, np.ones(ncol)); assert not y.any()
    with pytest.raises(fs.FsbError):
        fs.bcsr_A_mul_Bn(np.zeros(4), fs.BinaryCSR(1, 1, np.zeros(2, np.int32), np.zeros(0, np.int32)), np.ones(4), 0)


# ------------------------------------------------------------------ device-side construction and generator
def test_device_csr_build_is_bit_exact():
    rng = np.random.default_rng(5)
    nrow, ncol, nnz = 5000, 1234, 100000
    rows, cols, vals = _random_case(rng, nrow, ncol, nnz, long_row=9000)
    rows[1000:1100] = rows[1000]; cols[1000:1100] = cols[1000]      # duplicate coordinates with different values
    for v in (None, vals):
        A = fs.SparseBinaryMatrix(nrow, ncol, rows, cols) if v is None else fs.SparseDoubleMatrix(nrow, ncol, rows, cols, v)
        rp, cc, vv = fs.DeviceMatrix.of(A).download_csr()
        orp, occ, ovv = oracle.csr_from_coo(nrow, rows, cols, v)
        assert np.array_equal(rp, orp) and np.array_equal(cc, occ) and (v is None or np.array_equal(vv, ovv))


def test_synth_generator_device_equals_host():
    import torch
    for dist, seed in [(0, 0x5EED0002), (1, 0x5EED0004)]:
        n, nrow, ncol = 200000, 50000, 7001
        hr, hc, hv = fs.synth_coo_host(seed, dist, n, nrow, ncol, with_vals=True)
        M = fs.DeviceMatrix.synth(seed, dist, n, nrow, ncol, with_vals=True, keep_coo=True)
        dr, dc, dv = (t.cpu().numpy() for t in M.coo)
        assert np.array_equal(hr, dr) and np.array_equal(hc, dc) and np.array_equal(hv, dv)
        rp, cc, vv = M.download_csr()
        orp, occ, ovv = oracle.csr_from_coo(nrow, hr, hc, hv)
        assert np.array_equal(rp, orp) and np.array_equal(cc, occ) and np.array_equal(vv, ovv)


def test_row_shards_sum_to_full_product():
    import torch
    M = fs.DeviceMatrix.synth(11, 1, 300000, 40000, 5000, with_vals=True)
    rp, _, _ = M.download_csr()
    R = 8
    X = torch.randn(M.ncol * R, dtype=torch.float64, device="cuda"); Xt = torch.randn(M.nrow * R, dtype=torch.float64, device="cuda")
    Y = M.spmm(X, R); Z = M.spmm_t(Xt, R)
    b = fs.partition_rows(rp, 3)
    Zsum = torch.zeros_like(Z)
    for p in range(3):
        S = M.row_slice(b[p], b[p + 1])
        assert torch.equal(S.spmm(X, R), Y[b[p] * R: b[p + 1] * R])          # A x needs no exchange: shards are row slabs
        Zsum += S.spmm_t(Xt[b[p] * R: b[p + 1] * R].contiguous(), R)             # A' x: partials that an allreduce would sum
    assert torch.allclose(Zsum, Z, rtol=1e-12, atol=1e-10)


# ------------------------------------------------------------------ solver
def _cg_setup(name="sbm_100_50", bs=8):
    g = golden(name)
    nrow, ncol = int(g["nrow"]), int(g["ncol"])
    A = fs.new_sbm(nrow, ncol, g["rows"].size, g["rows"].copy(), g["cols"].copy())
    fs.sort_sbm(A)
    B = fs.new_bsbm(A, bs)
    fs.transpose(A)
    Bt = fs.new_bsbm(A, bs)
    return g, nrow, ncol, B, Bt


def test_cg_matches_reference():     # test_sparse.c:560-608 + SURVEY 8c extras
    g, nrow, ncol, B, Bt = _cg_setup()
    b = tvec(ncol); x = np.zeros(ncol)
    y = np.zeros(ncol); tmp = np.zeros(nrow)
    fs.bsbm_AtA(y, B, Bt, b, tmp, 5.0); assert_close(y, g["AtA_b"], 400.0, what="bsbm_AtA")
    it = fs.bsbm_cg(x, B, Bt, b, 5.0, 1e-6)
    assert it == 15 and abs(x[0] - 0.0638578) < 1e-4 and abs(x[1] + 0.0302702) < 1e-4 and abs(x[49] + 0.0284737361861) < 1e-9
    assert np.max(np.abs(x - g["cg_x"])) < 1e-9
    rows, cols = g["rows"], g["cols"]
    resid = oracle.coo_mul(nrow, rows, cols, None, oracle.coo_mul(nrow, rows, cols, None, x), transpose=True, ncol=ncol) + 5.0 * x - b
    assert np.linalg.norm(resid) < 1e-5
    X2 = np.zeros(ncol * 2); it2 = fs.bsbm_cg2(X2, B, Bt, g["cg2_B"], 5.0, 1e-6)
    assert it2 == 13 and np.max(np.abs(X2.reshape(ncol, 2) - g["cg2_X"])) < 1e-9
    assert abs(X2[0] - 0.0638578) < 1e-4 and abs(X2[2] + 0.0302702) < 1e-4
    with pytest.raises(fs.FsbError):       # cg.h:32-36 shape check
        fs.bsbm_cg(x, B, B, b, 5.0, 1e-6)


@pytest.mark.parametrize("R", [2, 5, 32])
def test_block_cg_n_rhs(R):
    """R-RHS block CG (no reference counterpart): oracle = bsbm_cg per column at tight tolerance + residual."""
    g, nrow, ncol, B, Bt = _cg_setup("rand_bin_300_70", bs=32)
    rng = np.random.default_rng(R)
    Bm = f64(rng.standard_normal((ncol, R))); lam = 3.0
    X = np.zeros(ncol * R); it = fs.bsbm_cgn(X, B, Bt, Bm, R, lam, 1e-9)
    X = X.reshape(ncol, R)
    assert 0 < it <= ncol
    hr, hc = g["hil_rows"], g["hil_cols"]
    Ao = oracle.blocked_from_coo(nrow, ncol, 32, hr, hc); Ato = oracle.blocked_from_coo(ncol, nrow, 32, hc, hr)
    for k in range(R):
        xo = np.zeros(ncol); O.fso_blocked_cg(dp(xo), Ao.ref(), Ato.ref(), dp(f64(Bm[:, k])), lam, 1e-11)
        assert np.max(np.abs(X[:, k] - xo)) <= 1e-7 * max(1.0, np.max(np.abs(xo))), f"column {k}"
        r = oracle.coo_mul(nrow, hr, hc, None, oracle.coo_mul(nrow, hr, hc, None, X[:, k]), transpose=True, ncol=ncol) + lam * X[:, k] - Bm[:, k]
        assert np.linalg.norm(r) <= 1e-7 * np.linalg.norm(Bm[:, k])


def test_cg_on_csr_handle_with_cached_transpose():
    import torch
    M = fs.DeviceMatrix.synth(3, 0, 200000, 20000, 2000)
    R = 8
    Bm = torch.randn(M.ncol * R, dtype=torch.float64, device="cuda")
    X, it = M.cg(Bm, R, lam=15.0, tol=1e-8)
    KX = M.ata(X, R, lam=15.0)
    rel = (KX - Bm).reshape(M.ncol, R).norm(dim=0) / Bm.reshape(M.ncol, R).norm(dim=0)
    assert it > 0 and float(rel.max()) < 1e-7
    assert torch.allclose(M.ata(X, R, lam=15.0, mode=1), KX, rtol=1e-11, atol=1e-9)      # fused scatter mode agrees


def test_linalg_reductions():        # test_sparse.c:511-544
    x = np.array([0.12, -0.82, 1.3, 0.5]); y = np.array([6.12, 0.19, 3.4, -4.1])
    assert abs(fs.pnormsq(x, 4) - 2.6268) < 1e-8 and abs(fs.pnormsq(y, 4) - 65.8605) < 1e-8 and abs(fs.pdot(x, y, 4) - 2.9486) < 1e-8
    o = np.zeros(3); fs.pouter2(o, x, 2)
    assert np.max(np.abs(o - [0.12 * 0.12 + 1.3 * 1.3, 0.82 * 0.82 + 0.5 * 0.5, 0.12 * -0.82 + 1.3 * 0.5])) < 1e-12
    X = np.array([0.95, 0.9, 0.16, 0.46, 0.86, 0.29]); Y = np.array([0.9695, 0.6678, 0.277, 0.1908, 1.108, 0.7632])
    d = np.zeros(3); fs.pdot2sym(d, X, Y, 3); assert np.max(np.abs(d - [1.918225, 0.910116, 1.32129])) < 1e-8
    S = np.zeros(4); fs.solve2sym(S, [0.59, 1.34, 0.86], [-1.21, 1.91, -0.82, 0.03])
    assert np.max(np.abs(S - [-64.0, 42.5, -22.05098039, 14.1745098])) < 1e-8
    gl = golden("linalg"); n = gl["x"].size
    assert abs(fs.pdot(gl["x"], gl["y"], n) - gl["dot"]) < 1e-10 and abs(fs.dist(gl["x"], gl["y"], n) - gl["dist"]) < 1e-10
    d = np.zeros(3); fs.pdot2sym(d, gl["X"], gl["Y"], n); assert np.max(np.abs(d - gl["dot2sym"])) < 1e-10


@pytest.mark.parametrize("R", [1, 2, 5, 8, 13, 16, 17, 24, 31, 32])
@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000, 40003])
def test_dense_gram_and_row_mix(R, n):
    """The tensor-core (DMMA) Gram and row-mix passes of the CG loop (cg.h:148-170, linalg.h:15-73)
    against numpy fp64, including ragged row counts, every R <= 32 and an unaligned R = 32 operand."""
    import torch
    rng = np.random.default_rng(1000 * R + n)
    L = fs.lib()
    for shift in ((0, 1) if R == 32 else (0,)):     # shift 1: 8-byte aligned only -> generic (scalar-load) path
        def dev(a):
            buf = torch.zeros(a.size + 2, dtype=torch.float64, device="cuda")
            v = buf[shift:shift + a.size]; v.copy_(torch.from_numpy(a.reshape(-1))); return buf, v
        I = rng.standard_normal((n, R)); O0 = rng.standard_normal((n, R)); Add = rng.standard_normal((n, R)); M = rng.standard_normal((R, R))
        _, dM = dev(M)
        scale = 4.0 * (R + 1)
        # Gram
        (_, dI), (_, dA) = dev(I), dev(Add)
        G = np.zeros((R, R))
        fs.check(L.fsb_gram_dev(dp(G), dI.data_ptr(), dA.data_ptr(), n, R, None))
        assert np.max(np.abs(G - I.T @ Add)) <= 1e-12 * max(n, 1) * 16
        fs.check(L.fsb_gram_dev(dp(G), dI.data_ptr(), dI.data_ptr(), n, R, None))
        assert np.max(np.abs(G - I.T @ I)) <= 1e-12 * max(n, 1) * 16 and np.array_equal(G, G.T)
        # mode 0 / 1 / 2
        _, dO = dev(O0)
        fs.check(L.fsb_rowmix_dev(0, dO.data_ptr(), dI.data_ptr(), None, dM.data_ptr(), None, n, R, None))
        assert np.max(np.abs(dO.cpu().numpy().reshape(n, R) - (O0 + I @ M))) <= 1e-13 * scale
        _, dO = dev(O0)
        fs.check(L.fsb_rowmix_dev(1, dO.data_ptr(), dI.data_ptr(), None, dM.data_ptr(), dp(G), n, R, None))
        want = O0 - I @ M
        assert np.max(np.abs(dO.cpu().numpy().reshape(n, R) - want)) <= 1e-13 * scale
        assert np.max(np.abs(G - want.T @ want)) <= 1e-12 * max(n, 1) * scale * scale and np.array_equal(G, G.T)
        _, dP = dev(I)                                # in place: P = Add + P M
        fs.check(L.fsb_rowmix_dev(2, dP.data_ptr(), dP.data_ptr(), dA.data_ptr(), dM.data_ptr(), None, n, R, None))
        assert np.max(np.abs(dP.cpu().numpy().reshape(n, R) - (Add + I @ M))) <= 1e-13 * scale


@pytest.mark.parametrize("with_vals", [False, True])
def test_staged_builds_and_column_passes_agree_bitwise(with_vals):
    """The per-handle autotune of the staged SpMM picks between column passes (1 | 2) and the lean / deep
    build of the kernel by timing: every candidate must give the same bits (each column's sum is taken in
    stored order), and the choice it records must be one of them."""
    import torch
    L = fs.lib()
    nrow, ncol, nnz, R = 400_000, 120_000, 6_000_000, 32       # >= 4M entries: the autotune engages
    A = fs.DeviceMatrix.synth(0xABCD, 1, nnz, nrow, ncol, with_vals=with_vals)
    X = torch.randn(ncol * R, dtype=torch.float64, device="cuda")
    ref = None
    try:
        for slabs in (1, 2):
            for deep in (0, 1):
                fs.check(L.fsb_tune_csr_algo(2, 0, 0)); fs.check(L.fsb_tune_csr_spmm(0, 0, 0, slabs)); fs.check(L.fsb_tune_csr_staged(deep))
                Y = A.spmm(X, R)
                ref = Y if ref is None else ref
                assert torch.equal(Y, ref), f"slabs={slabs} deep={deep} differs"
    finally:
        fs.check(L.fsb_tune_csr_algo(0, 0, 0)); fs.check(L.fsb_tune_csr_spmm(0, 0, 0, 0)); fs.check(L.fsb_tune_csr_staged(-1))
    assert A.tuning()[0] == 0                                   # explicit settings do not touch the recorded choice
    Y = A.spmm(X, R)                                            # first automatic product: times the candidates
    tR, passes, deep = A.tuning()
    assert tR == R and passes in (1, 2) and torch.equal(Y, ref)
    assert torch.equal(A.spmm(X, R), ref)                       # steady state uses the recorded choice
    # the fused epilogue Y = A X + lambda Z agrees with the two-step form to rounding of one fma
    K = A.ata(X, R, lam=0.75)                                   # A'(A X) + 0.75 X through the fused epilogue
    K2 = A.spmm_t(ref, R) + 0.75 * X
    assert float((K - K2).abs().max()) <= 1e-12 * float(K2.abs().max())


def test_noise_rhs_and_sampling_step():
    """One Macau-style sampling step with the matrix resident (bench_a_mul_b.c:334-360): device noise,
    B = A'N + sqrt(lambda) E fused into the A' product, block-CG solve."""
    import torch
    L = fs.lib()
    n = 100_003
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    fs.check(L.fsb_randn_dev(d.data_ptr(), n, 12345, None))
    h = np.zeros(n); fs.check(L.fsb_randn_host(dp(h), n, 12345))
    g = d.cpu().numpy()
    assert np.max(np.abs(g - h)) < 1e-12                       # same counter-based stream on both sides
    assert abs(g.mean()) < 0.02 and abs(g.std() - 1.0) < 0.02 and abs((g ** 3).mean()) < 0.05 and abs((g ** 4).mean() - 3.0) < 0.15
    nrow, ncol, nnz, R, lam, seed = 50_000, 4_000, 600_000, 8, 15.0, 99
    rows, cols, _ = fs.synth_coo_host(31337, 1, nnz, nrow, ncol)
    A = fs.DeviceMatrix.of(fs.new_bcsr(nnz, nrow, ncol, rows, cols))
    Bm = A.noise_rhs(R, lam, seed)
    Nh = np.zeros(nrow * R); fs.check(L.fsb_randn_host(dp(Nh), nrow * R, seed ^ 0x9E3779B97F4A7C15))
    Eh = np.zeros(ncol * R); fs.check(L.fsb_randn_host(dp(Eh), ncol * R, seed + 0x5bd1e995))
    orp, occ, _ = oracle.csr_from_coo(ncol, cols, rows)        # the oracle's CSR of A'
    want = oracle.csr_mul(ncol, orp, occ, None, Nh.reshape(nrow, R), R).reshape(-1) + np.sqrt(lam) * Eh
    assert_close(Bm.cpu().numpy(), want, scale=4.0 * row_scale(ncol, cols) + 8.0, what="noise rhs")
    X, it = A.cg(Bm, R, lam=lam, tol=1e-8)
    res = (A.ata(X, R, lam=lam) - Bm).reshape(ncol, R).norm(dim=0) / Bm.reshape(ncol, R).norm(dim=0)
    assert it > 0 and float(res.max()) < 1e-7
    assert not torch.equal(A.noise_rhs(R, lam, seed + 1), Bm)    # a new seed is a new sample


def test_invalid_indices_are_rejected_at_upload(tmp_path):
    """The reference trusts its indices; on the GPU an out-of-range one would be an illegal address that kills the
    context, so every upload / load path validates once on the device and fails with an error instead."""
    nrow, ncol, nnz = 50, 20, 200
    rng = np.random.default_rng(3)
    rows = np.sort(rng.integers(0, nrow, nnz)).astype(np.int32); cols = rng.integers(0, ncol, nnz).astype(np.int32)
    good = fs.new_bcsr(nnz, nrow, ncol, rows, cols)
    x = tvec(ncol); y = np.zeros(nrow)

    def expect_rejected(fn):
        with pytest.raises(fs.FsbError) as e:
            fn()
        assert "out of range" in str(e.value) or "offset array" in str(e.value), str(e.value)

    for bad_col in (ncol, -1, 2 ** 30):
        c = good.cols.copy(); c[17] = bad_col
        expect_rejected(lambda: fs.bcsr_A_mul_B(y, fs.BinaryCSR(nrow, ncol, good.row_ptr, c), x))
    rp = good.row_ptr.copy(); rp[10] = rp[12] + 3              # a decreasing step
    expect_rejected(lambda: fs.bcsr_A_mul_B(y, fs.BinaryCSR(nrow, ncol, rp, good.cols), x))
    r = rows.copy(); r[5] = nrow                              # COO with a row index one past the end
    expect_rejected(lambda: fs.A_mul_B(y, fs.SparseBinaryMatrix(nrow, ncol, r, cols), x))
    # a .csr.bin whose column array was corrupted on disk
    path = str(tmp_path / "m.csr.bin")
    fs.serialize_to_file(good, path)
    raw = bytearray(open(path, "rb").read())
    raw[-4:] = np.array([ncol + 5], dtype=np.int32).tobytes()
    open(path, "wb").write(bytes(raw))
    expect_rejected(lambda: fs.DeviceMatrix.load_csr_bin(path))
    # the context is still healthy
    fs.bcsr_A_mul_B(y, good, x)
    want = np.zeros(nrow); np.add.at(want, rows, x[cols])
    assert np.max(np.abs(y - want)) < 1e-12


def _write_coo_file(path, nrow, ncol, rows, cols, vals=None):
    """The reference's raw COO format (sparse.h:112-139, dsparse.h:64-93): 3 x int64, 1-based int32 indices."""
    with open(path, "wb") as f:
        np.array([nrow, ncol, rows.size], dtype=np.int64).tofile(f)
        (rows + 1).astype(np.int32).tofile(f)
        (cols + 1).astype(np.int32).tofile(f)
        if vals is not None:
            vals.astype(np.float64).tofile(f)


def test_file_loaders_straight_to_hbm(tmp_path):
    """read_sbm / read_sdm / .csr.bin files loaded directly into HBM (chunked, pinned staging) give the
    same structure, bit for bit, as the host loaders + new_bcsr / new_csr (csr.h:30-67, 375-422)."""
    for name, with_vals in (("sbm-100-50.data", False), ("sdm-100-50.data", True)):
        path = os.path.join(DATA, name)
        M = fs.DeviceMatrix.load_coo_file(path, with_vals)
        H = fs.read_sdm(path) if with_vals else fs.read_sbm(path)
        want = fs.new_csr(H.nnz, H.nrow, H.ncol, H.rows, H.cols, H.vals) if with_vals else fs.new_bcsr(H.nnz, H.nrow, H.ncol, H.rows, H.cols)
        rp, cc, vv = M.download_csr()
        assert (M.nrow, M.ncol, M.nnz) == (H.nrow, H.ncol, H.nnz)
        assert np.array_equal(rp, want.row_ptr) and np.array_equal(cc, want.cols)
        assert (vv is None) == (not with_vals) and (vv is None or np.array_equal(vv, want.vals))
    # several staging chunks per array (5M entries = 20 MB of indices, 40 MB of values), duplicates, empty rows
    rng = np.random.default_rng(77)
    nrow, ncol, nnz = 300_000, 70_000, 5_000_000
    rows = rng.integers(0, nrow // 2, nnz).astype(np.int32) * 2     # odd rows stay empty
    cols = rng.integers(0, ncol, nnz).astype(np.int32)
    vals = rng.standard_normal(nnz)
    for with_vals in (False, True):
        path = str(tmp_path / ("big.sdm" if with_vals else "big.sbm"))
        _write_coo_file(path, nrow, ncol, rows, cols, vals if with_vals else None)
        M = fs.DeviceMatrix.load_coo_file(path, with_vals)
        orp, occ, ovv = oracle.csr_from_coo(nrow, rows, cols, vals if with_vals else None)
        rp, cc, vv = M.download_csr()
        assert np.array_equal(rp, orp) and np.array_equal(cc, occ) and (not with_vals or np.array_equal(vv, ovv))
    # .csr.bin written by serialize_to_file, read back into HBM
    B = fs.new_bcsr(nnz, nrow, ncol, rows, cols)
    path = str(tmp_path / "big.csr.bin")
    fs.serialize_to_file(B, path)
    M = fs.DeviceMatrix.load_csr_bin(path)
    rp, cc, _ = M.download_csr()
    assert (M.nrow, M.ncol, M.nnz) == (nrow, ncol, nnz) and np.array_equal(rp, B.row_ptr) and np.array_equal(cc, B.cols)
    x = tvec(ncol); y = np.zeros(nrow); fs.bcsr_A_mul_B(y, B, x)
    import torch
    yd = M.spmm(torch.from_numpy(x).cuda(), 1).cpu().numpy()
    assert np.max(np.abs(yd - y)) <= 1e-12 * row_scale(nrow, rows)
    # error behaviour: truncated and foreign files are rejected
    bad = str(tmp_path / "bad.csr.bin")
    open(bad, "wb").write(open(path, "rb").read()[:1000])
    with pytest.raises(fs.FsbError):
        fs.DeviceMatrix.load_csr_bin(bad)
    with pytest.raises(fs.FsbError):
        fs.DeviceMatrix.load_csr_bin(os.path.join(DATA, "sbm-100-50.data"))
    with pytest.raises(fs.FsbError):
        fs.DeviceMatrix.load_coo_file(str(tmp_path / "missing.data"))


# ------------------------------------------------------------------ BASELINE.json full size: size-independent properties
def test_full_size_c2_properties():
    """C2: binary CSR 10M x 1M, 200M nnz, R = 32.  Checks: (i) a slab of rows against the oracle,
    (ii) checksum of checksums sum_r Y[r,:] == sum_c count[c] X[c,:], (iii) linearity,
    (iv) A'(A X) two-pass == fused scatter == <AX, AX> identity."""
    import torch
    N, F, NNZ, R = 10_000_000, 1_000_000, 200_000_000, 32
    M = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F, keep_coo=True)
    cols_t = M.coo[1]
    cnt = torch.bincount(cols_t.long(), minlength=F).double()
    del M.coo
    c = torch.arange(F, device="cuda", dtype=torch.float64)[:, None]; k = torch.arange(R, device="cuda", dtype=torch.float64)[None, :]
    X = torch.sin(7.0 * c + 17.0 * k + 0.3).reshape(-1).contiguous()
    Y = M.spmm(X, R)
    # (i) first 3000 rows against the oracle
    S = M.row_slice(0, 3000); rp, cc, _ = S.download_csr()
    want = oracle.csr_mul(3000, rp, cc, None, X.cpu().numpy(), R)
    assert_close(Y[: 3000 * R].cpu().numpy(), want, 64.0, what="C2 row slab")
    # (ii) column sums
    colsum = Y.reshape(N, R).sum(0); ref = (cnt[:, None] * X.reshape(F, R)).sum(0)
    assert torch.allclose(colsum, ref, rtol=1e-9, atol=1e-3)
    # (iii) linearity
    X2 = torch.cos(3.0 * c - 5.0 * k).reshape(-1).contiguous()
    Y2 = M.spmm(X2, R); Y12 = M.spmm(X + 2.0 * X2, R)
    assert torch.allclose(Y12, Y + 2.0 * Y2, rtol=1e-12, atol=1e-10)
    del Y2, Y12
    # (iv) <X, A'A X> == <AX, AX>, per column
    Z = M.ata(X, R)
    lhs = (X.reshape(F, R) * Z.reshape(F, R)).sum(0); rhs = (Y.reshape(N, R) ** 2).sum(0)
    assert torch.allclose(lhs, rhs, rtol=1e-10)


def test_full_size_c3_c4_c5_properties():
    """The other BASELINE.json configs at full size through size-independent properties.
    C3 (double CSR 10M x 1M, 200M nnz): entry-sum identity sum_r y[r] == sum_j vals[j] x[cols[j]], a row slab against
    the oracle, the adjoint identity <A x, u> == <x, A'u>.  C4 (power-law columns, R = 32): the blocked (Hilbert) and
    column-blocked formats against CSR on the same entries.  C5: the block-CG solve on the C2 matrix reaches its tolerance
    (residual recomputed with the product kernels)."""
    import torch
    N, F, NNZ = 10_000_000, 1_000_000, 200_000_000
    # ---- C3
    A = fs.DeviceMatrix.synth(0x5EED0003, 0, NNZ, N, F, with_vals=True, keep_coo=True)
    _, cols_t, vals_t = A.coo
    x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    y = A.spmm(x, 1)
    entry_sum = float((vals_t * x[cols_t.long()]).sum())
    del A.coo, cols_t, vals_t
    assert abs(float(y.sum()) - entry_sum) <= 1e-11 * NNZ * 0.1            # |terms| <= 0.1
    S = A.row_slice(0, 3000); rp, cc, vv = S.download_csr()
    want = oracle.csr_mul(3000, rp, cc, vv, x.cpu().numpy().reshape(F, 1), 1)
    assert_close(y[:3000].cpu().numpy(), want, 8.0, what="C3 row slab")
    u = torch.cos(3.0 * torch.arange(N, device="cuda", dtype=torch.float64)).contiguous()
    z = A.spmm_t(u, 1)
    lhs, rhs = float((y * u).sum()), float((x * z).sum())
    assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs), float(y.abs().sum()))
    A.free(); del A, y, z, u, S
    # ---- C4
    R = 32
    M = fs.DeviceMatrix.synth(0x5EED0004, 1, NNZ, N, F, keep_coo=True)
    rows_t, cols_t, _ = M.coo
    X = torch.randn(F * R, dtype=torch.float64, device="cuda")
    Y = M.spmm(X, R)
    Bk = fs.DeviceMatrix.blocked_from_coo_tensors(N, F, rows_t, cols_t, None, 512, order=1)
    assert float((Bk.spmm(X, R) - Y).abs().max()) <= 1e-12 * 4.0 * 4096
    Bk.free(); del Bk
    Cb = fs.DeviceMatrix.cbcsr_from_coo_tensors(N, F, rows_t, cols_t, 65536)
    assert float((Cb.spmm(X, R) - Y).abs().max()) <= 1e-12 * 4.0 * 4096
    Cb.free(); del Cb, M.coo, rows_t, cols_t
    M.free(); del M, X, Y
    # ---- C5
    K = fs.DeviceMatrix.synth(0x5EED0002, 0, NNZ, N, F)
    Bm = K.noise_rhs(R, 15.0, 7)
    Xs, it = K.cg(Bm, R, lam=15.0, tol=1e-6)
    res = (K.ata(Xs, R, lam=15.0) - Bm).reshape(F, R).norm(dim=0) / Bm.reshape(F, R).norm(dim=0)
    assert 5 <= it <= 30 and float(res.max()) < 2e-6


# ------------------------------------------------------------------ device-side builders of the blocked formats
@pytest.mark.parametrize("bs", [64, 512])
@pytest.mark.parametrize("with_vals", [False, True])
def test_device_blocked_builder_matches_host_sort_bsbm(bs, with_vals):
    import torch
    nrow, ncol, nnz, R = 3000, 5000, 40000, 8
    rows, cols, vals = fs.synth_coo_host(77, 1, nnz, nrow, ncol, with_vals=True)
    keep = np.unique(rows.astype(np.int64) * ncol + cols, return_index=True)[1]      # unique coordinates: order is unique
    rows, cols, vals = rows[keep], cols[keep], (vals[keep] if with_vals else None)
    A = fs.SparseBinaryMatrix(nrow, ncol, rows, cols) if vals is None else fs.SparseDoubleMatrix(nrow, ncol, rows, cols, vals)
    Bl = fs.new_bsbm(A, bs); fs.sort_bsbm(Bl)
    X = torch.randn(ncol * R, dtype=torch.float64, device="cuda")
    Yh = fs.DeviceMatrix.of(Bl).spmm(X, R)
    tr, tc = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
    tv = torch.from_numpy(vals).cuda() if vals is not None else None
    for order in (0, 1, 2):
        D = fs.DeviceMatrix.blocked_from_coo_tensors(nrow, ncol, tr, tc, tv, bs, order=order)
        Yd = D.spmm(X, R)
        if order == 1:
            assert torch.equal(Yd, Yh), "device Hilbert order differs from host sort_bsbm order"
        assert torch.allclose(Yd, Yh, rtol=1e-13, atol=1e-12)


def test_device_cbcsr_builder_matches_host():
    import torch
    nrow, ncol, nnz, R = 2500, 3000, 50000, 4
    rows, cols, _ = fs.synth_coo_host(78, 1, nnz, nrow, ncol)
    Cb = fs.new_cbcsr(256, nnz, nrow, ncol, rows, cols)
    X = torch.randn(ncol * R, dtype=torch.float64, device="cuda")
    Yh = fs.DeviceMatrix.of(Cb).spmm(X, R)
    D = fs.DeviceMatrix.cbcsr_from_coo_tensors(nrow, ncol, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), 256)
    assert D.nblocks == Cb.nblocks and torch.equal(D.spmm(X, R), Yh)


# ------------------------------------------------------------------ residency cache (fsb_dropin.cpp)
@pytest.mark.parametrize("n", [2000, 700000])        # inline full hash / background-worker hash (> 4 MB of arrays)
def test_residency_cache_sees_in_place_edits(n):
    """The host arrays are the source of truth (the reference reads them on every call): edits between two sampled
    positions of the round-1 fingerprint, in vals[], cols[] and row_ptr[], must all be seen by the next product."""
    m = 97
    rows = np.arange(n, dtype=np.int32); cols = (np.arange(n) % m).astype(np.int32); vals = 1.0 + (np.arange(n) % 5)
    M = fs.new_csr(n, n, m, rows, cols, vals)
    x = 1.0 + np.arange(m, dtype=np.float64); y = np.zeros(n)
    fs.csr_A_mul_B(y, M, x)
    k = n // 2 + 1
    assert y[k] == vals[k] * x[cols[k]]
    M.vals[k] = 4096.0
    fs.csr_A_mul_B(y, M, x)
    assert y[k] == 4096.0 * x[cols[k]], "stale device copy after an in-place edit of vals[]"
    M.cols[k] = (cols[k] + 3) % m
    fs.csr_A_mul_B(y, M, x)
    assert y[k] == 4096.0 * x[(cols[k] + 3) % m], "stale device copy after an in-place edit of cols[]"
    # move one entry from row k to row k+1 (row_ptr edit)
    M.row_ptr[k + 1] -= 1
    fs.csr_A_mul_B(y, M, x)
    assert y[k] == 0.0 and y[k + 1] == 4096.0 * x[(cols[k] + 3) % m] + vals[k + 1] * x[cols[k + 1]]
    # unchanged arrays: served from the cache (no growth in entries)
    e0 = C.c_long(); fs.lib().fsb_cache_stats(C.byref(e0), None)
    fs.csr_A_mul_B(y, M, x)
    e1 = C.c_long(); fs.lib().fsb_cache_stats(C.byref(e1), None)
    assert e1.value == e0.value


def test_residency_cache_evicts_least_recently_used(tmp_path):
    """FSB_CACHE_MAX_MB caps the HBM the cache may pin: matrices whose host arrays were dropped with plain free()
    (never fsb_cache_drop) do not accumulate."""
    import subprocess, sys
    code = r'''
import ctypes as C, numpy as np, libfastsparse_b200 as fs
keep = []
for i in range(6):
    n = 300000
    M = fs.new_csr(n, n, 64, np.arange(n, dtype=np.int32), (np.arange(n) % 64).astype(np.int32), np.ones(n))
    y = np.zeros(n); fs.csr_A_mul_B(y, M, np.ones(64)); assert y[5] == 1.0
    keep.append(M)
e, b = C.c_long(), C.c_long(); fs.lib().fsb_cache_stats(C.byref(e), C.byref(b))
assert e.value <= 3 and b.value <= 12 << 20, (e.value, b.value)
y = np.zeros(300000); fs.csr_A_mul_B(y, keep[0], np.ones(64)); assert y[7] == 1.0      # evicted entries are simply re-uploaded
print("EVICT OK", e.value, b.value)
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, FSB_CACHE_MAX_MB="12", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    assert r.returncode == 0 and "EVICT OK" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------ round-1 review items
def test_device_builders_validate_their_indices():
    """fsb_blocked_from_coo_dev / fsb_cbcsr_from_coo_dev compute keys from unchecked indices: an out-of-range row used
    to index past the class offsets.  Every builder now checks both index arrays first."""
    import torch
    nrow, ncol, nnz = 1000, 300, 5000
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rows = torch.randint(0, nrow, (nnz,), device="cuda", generator=g, dtype=torch.int32)
    cols = torch.randint(0, ncol, (nnz,), device="cuda", generator=g, dtype=torch.int32)
    for bad_r, bad_c in ((nrow, 0), (-1, 0), (0, ncol), (0, -7)):
        r = rows.clone(); c = cols.clone(); r[123] = bad_r if bad_r else r[123]; c[123] = bad_c if bad_c else c[123]
        for build in (lambda: fs.DeviceMatrix.blocked_from_coo_tensors(nrow, ncol, r, c, None, 64, order=1),
                      lambda: fs.DeviceMatrix.cbcsr_from_coo_tensors(nrow, ncol, r, c, 128)):
            with pytest.raises(fs.FsbError) as e:
                build()
            assert e.value.code == 1 and "out of range" in str(e.value)
    # in-block key guard (ADVICE low): row_xy2d keys reach n*n when ncol < n
    one = torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(fs.FsbError) as e:
        fs.DeviceMatrix.blocked_from_coo_tensors(1 << 21, 1 << 10, one, one, None, 1 << 21, order=1)
    assert "40 bits" in str(e.value)
    with pytest.raises(fs.FsbError) as e:
        fs.DeviceMatrix.blocked_from_coo_tensors(1 << 21, 1 << 20, one, one, None, 1 << 21, order=2)
    assert "40 bits" in str(e.value)
    fs.DeviceMatrix.blocked_from_coo_tensors(nrow, ncol, rows, cols, None, 64, order=1)       # still healthy


def test_blocked_upload_validates_block_metadata():
    nrow, ncol = 40, 10
    A = fs.new_sbm(nrow, ncol, 6, [0, 9, 10, 25, 39, 39], [1, 2, 3, 4, 5, 6])
    B = fs.new_bsbm(A, 10)
    x = tvec(ncol); y = np.zeros(nrow)
    fs.bsbm_A_mul_B(y, B, x)
    bad = fs.new_bsbm(A, 10); bad.start_row[2] = 5           # decreasing start_row
    with pytest.raises(fs.FsbError) as e:
        fs.bsbm_A_mul_B(y, bad, x)
    assert "start_row" in str(e.value)
    bad = fs.new_bsbm(A, 10); bad.rows[0][0] = 15            # a row outside its block
    with pytest.raises(fs.FsbError) as e:
        fs.bsbm_A_mul_B(y, bad, x)
    assert "outside its block" in str(e.value)
    fs.bsbm_A_mul_B(y, B, x)


def test_row_range_aliases_stay_on_row_local_kernels():
    """The chunked host product passes row-range aliases (absolute row_ptr offsets) to the dispatcher; with the stream
    kernel forced (fsb_tune_csr_algo(3), R = 2) they used to reach the merge-path kernel, which assumes row_ptr[0] == 0."""
    nrow, ncol, nnz, R = 4_500_000, 1000, 9_000_000, 2
    M = fs.DeviceMatrix.synth(99, 0, nnz, nrow, ncol)
    rp, cc, _ = M.download_csr()
    X = rhs_matrix(ncol, R).reshape(-1)
    want = np.zeros((nrow, R)); np.add.at(want, np.repeat(np.arange(nrow), np.diff(rp)), X.reshape(ncol, R)[cc])
    fs.check(fs.lib().fsb_tune_csr_algo(3, 0, 0))
    try:
        Y = np.zeros(nrow * R)
        fs.check(fs.lib().fsb_spmm_host(M.h, Y.ctypes.data_as(C.POINTER(C.c_double)), X.ctypes.data_as(C.POINTER(C.c_double)), R))
    finally:
        fs.check(fs.lib().fsb_tune_csr_algo(0, 0, 0))
    assert np.max(np.abs(Y.reshape(nrow, R) - want)) < 1e-10


def test_autotune_remembers_every_width():
    """A handle alternating between operand widths keeps one tuning slot per width (no re-tuning at 8x product cost)."""
    import torch
    M = fs.DeviceMatrix.synth(5, 0, 6_000_000, 300_000, 40_000)
    launches = {}
    for visit, R in enumerate((32, 8, 32, 8)):
        X = torch.randn(40_000 * R, dtype=torch.float64, device="cuda")
        before = fs.launch_count()
        M.spmm(X, R); M.spmm(X, R)
        torch.cuda.synchronize()
        launches[(R, visit >= 2)] = fs.launch_count() - before
        assert M.tuning()[0] == R
    assert launches[(32, True)] <= 4 and launches[(8, True)] <= 4, f"a width was re-tuned: {launches}"
    assert launches[(32, False)] > launches[(32, True)], f"the first visit should have timed its candidates: {launches}"


@pytest.mark.parametrize("with_vals", [False, True])
def test_xblocked_transposed_spmv_matches_the_plain_transpose(with_vals):
    """y = A'x with one right-hand side switches to the x-blocked transpose when x outgrows L2 (fsb_capi.cu
    spmv_t_xblocked).  Forced here on a small matrix with many blocks (ragged last block, empty cells, a 3000-entry
    column): same result as the plain cached transpose and as the oracle's COO scatter (sparse.h:68-75)."""
    import torch
    nrow, ncol, nnz = 100_003, 2_000, 900_000
    rows, cols, vals = fs.synth_coo_host(77, 1, nnz, nrow, ncol, with_vals=with_vals)
    M = fs.DeviceMatrix.from_coo_tensors(nrow, ncol, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(),
                                         torch.from_numpy(vals).cuda() if with_vals else None)
    x = tvec(nrow)
    xd = torch.from_numpy(x).cuda()
    L = fs.lib()
    fs.check(L.fsb_tune(b"t_xblock", 0))
    plain = M.spmm_t(xd, 1).cpu().numpy()
    fs.check(L.fsb_tune(b"t_xblock", 1)); fs.check(L.fsb_tune(b"t_xblock_min_kb", 1)); fs.check(L.fsb_tune(b"t_xblock_kb", 16))
    try:
        blocked = M.spmm_t(xd, 1).cpu().numpy()
        z = M.ata(torch.from_numpy(tvec(ncol)).cuda(), 1, lam=0.5).cpu().numpy()
    finally:
        fs.check(L.fsb_tune(b"t_xblock_min_kb", 0)); fs.check(L.fsb_tune(b"t_xblock_kb", 32 << 10))   # 0 = the built-in thresholds
    want = oracle.coo_mul(nrow, rows, cols, vals, x, transpose=True, ncol=ncol)
    scale = 2.0 * float(np.bincount(cols, minlength=ncol).max())
    assert_close(blocked, want, scale=scale, what="x-blocked A'x vs oracle")
    assert_close(blocked, plain, scale=scale, what="x-blocked A'x vs plain transpose")
    # A'(A x) + lambda x through the same path
    xc = tvec(ncol)
    ax = oracle.coo_mul(nrow, rows, cols, vals, xc)
    want2 = oracle.coo_mul(nrow, rows, cols, vals, ax, transpose=True, ncol=ncol) + 0.5 * xc
    assert_close(z, want2, scale=scale * scale, what="A'(A x) + lambda x via the x-blocked transpose")


@pytest.mark.parametrize("with_vals", [False, True])
def test_device_global_hilbert_sort_matches_host(with_vals):
    """sort_sbm / sort_sdm on the device: bit-exact against the host routine (itself pinned to the
    reference's goldens) on unique coordinates, values carried; duplicates keep their input order; the drop-in
    entry point picks the device path above FSB_SORT_DEVICE_MIN entries."""
    import torch
    rng = np.random.default_rng(11)
    for nrow, ncol, nnz in [(100, 50, 504), (70_000, 131_072, 400_000), (5, 3, 15), (1, 1, 1)]:
        flat = rng.choice(nrow * ncol, size=nnz, replace=False)            # unique coordinates
        rows = (flat // ncol).astype(np.int32); cols = (flat % ncol).astype(np.int32)
        vals = rng.random(nnz) if with_vals else None
        hr, hc = rows.copy(), cols.copy(); hv = vals.copy() if with_vals else None
        fs.check(fs.lib().fsb_host_sort_coo_hilbert(nrow, ncol, nnz, fs.api._ip(hr), fs.api._ip(hc), fs.api._dp(hv)))
        dr, dc = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda()
        dv = torch.from_numpy(vals).cuda() if with_vals else None
        fs.check(fs.lib().fsb_sort_coo_hilbert_dev(nrow, ncol, nnz, dr.data_ptr(), dc.data_ptr(), dv.data_ptr() if with_vals else None))
        assert np.array_equal(dr.cpu().numpy(), hr) and np.array_equal(dc.cpu().numpy(), hc)
        if with_vals:
            assert np.array_equal(dv.cpu().numpy(), hv)
        # host-array entry point (upload / sort / download)
        ar, ac = rows.copy(), cols.copy(); av = vals.copy() if with_vals else None
        fs.check(fs.lib().fsb_sort_coo_hilbert(nrow, ncol, nnz, fs.api._ip(ar), fs.api._ip(ac), fs.api._dp(av)))
        assert np.array_equal(ar, hr) and np.array_equal(ac, hc) and (not with_vals or np.array_equal(av, hv))
        # keys strictly increasing (test_sparse.c:283-289)
        n = fs.ceilPower2(max(nrow, ncol))
        keys = np.array([fs.xy2d(n, int(r), int(c)) for r, c in zip(hr[:2000], hc[:2000])])
        assert np.all(np.diff(keys) > 0)
    # duplicates: binary entries are interchangeable; valued duplicates keep their input order (stable radix sort)
    rows = np.array([3, 1, 3, 3, 0], np.int32); cols = np.array([2, 1, 2, 2, 0], np.int32); vals = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    dr, dc, dv = torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), torch.from_numpy(vals).cuda()
    fs.check(fs.lib().fsb_sort_coo_hilbert_dev(4, 3, 5, dr.data_ptr(), dc.data_ptr(), dv.data_ptr()))
    r, c, v = dr.cpu().numpy(), dc.cpu().numpy(), dv.cpu().numpy()
    dup = (r == 3) & (c == 2)
    assert dup.sum() == 3 and list(v[dup]) == [1.0, 3.0, 4.0]
    with pytest.raises(fs.FsbError):
        bad = torch.tensor([0, 9], dtype=torch.int32, device="cuda")
        fs.check(fs.lib().fsb_sort_coo_hilbert_dev(4, 3, 2, bad.data_ptr(), bad.data_ptr(), None))


@pytest.mark.parametrize("with_vals", [False, True])
@pytest.mark.parametrize("order", [1, 2])
def test_device_sort_of_host_blocked_structure_matches_host(with_vals, order):
    """sort_bsbm / sort_bsbm_byrow / sort_bsdm on a HOST BlockedSBM / BlockedSDM through the device (one keyed radix sort
    over all blocks, fsb_sort_blocked): the same per-block order as the host routines, block boundaries respected, a
    ragged last block and empty blocks included, values carried."""
    if with_vals and order == 2:
        pytest.skip("the reference has no by-row sort for the double-valued format")
    rng = np.random.default_rng(21 + order)
    nrow, ncol, nnz, bs = 10_007, 5_000, 300_000, 512
    flat = rng.choice(nrow * ncol, size=nnz, replace=False)
    rows = (flat // ncol).astype(np.int32); cols = (flat % ncol).astype(np.int32)
    rows[rows // bs == 3] = 4 * bs                      # block 3 becomes empty
    flat2 = np.unique(rows.astype(np.int64) * ncol + cols)        # keep coordinates unique after the move
    rows = (flat2 // ncol).astype(np.int32); cols = (flat2 % ncol).astype(np.int32)
    perm = rng.permutation(rows.size); rows, cols = rows[perm], cols[perm]
    vals = rng.random(rows.size) if with_vals else None
    mk = (lambda: fs.new_bsdm(fs.new_sdm(nrow, ncol, rows.size, rows.copy(), cols.copy(), vals.copy()), bs)) if with_vals else \
         (lambda: fs.new_bsbm(fs.new_sbm(nrow, ncol, rows.size, rows.copy(), cols.copy()), bs))
    Bh, Bd = mk(), mk()
    sort = fs.sort_bsbm if order == 1 else fs.sort_bsbm_byrow
    sort(Bh, device=False)
    sort(Bd, device=True)
    assert Bh.nblocks == Bd.nblocks and int(Bh.nnz[3]) == 0
    for b in range(Bh.nblocks):
        assert np.array_equal(Bh.rows[b], Bd.rows[b]) and np.array_equal(Bh.cols[b], Bd.cols[b]), f"block {b}"
        if with_vals:
            assert np.array_equal(Bh.vals[b], Bd.vals[b]), f"values of block {b}"
    # the product is unchanged by the re-ordering (test_sparse.c:372-377)
    x = tvec(ncol); y = np.zeros(nrow); y2 = np.zeros(nrow)
    (fs.bsdm_A_mul_B if with_vals else fs.bsbm_A_mul_B)(y, Bd, x)
    (fs.bsdm_A_mul_B if with_vals else fs.bsbm_A_mul_B)(y2, mk(), x)
    assert_close(y, y2, scale=2.0 * 64, what="product after the device sort")


def test_round2_paths_on_degenerate_inputs():
    """Empty and minimal inputs through the paths added in round 2: device sorts with no entries, the x-blocked
    transpose of a matrix with a single column / no entries, a one-block operand."""
    import torch
    L = fs.lib()
    z = torch.zeros(1, dtype=torch.int32, device="cuda")
    fs.check(L.fsb_sort_coo_hilbert_dev(5, 7, 0, z.data_ptr(), z.data_ptr(), None))            # nothing to sort
    e = np.zeros(0, np.int32)
    fs.check(L.fsb_sort_coo_hilbert(5, 7, 0, fs.api._ip(e), fs.api._ip(e), None))
    B = fs.new_bsbm(fs.new_sbm(6, 4, 0, e.copy(), e.copy()), 4)                                 # blocked structure without entries
    fs.sort_bsbm(B, device=True)
    y = np.ones(6); fs.bsbm_A_mul_B(y, B, np.ones(4)); assert not y.any()
    fs.check(L.fsb_tune(b"t_xblock_min_kb", 1)); fs.check(L.fsb_tune(b"t_xblock_kb", 1))
    try:
        # one column: every cell row of the blocked transpose belongs to column 0
        nrow = 3000
        rows = np.arange(0, nrow, 3, dtype=np.int32); cols = np.zeros(rows.size, np.int32); vals = 1.0 + np.arange(rows.size)
        M = fs.DeviceMatrix.from_coo_tensors(nrow, 1, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), torch.from_numpy(vals).cuda())
        x = torch.from_numpy(tvec(nrow)).cuda()
        got = M.spmm_t(x, 1).cpu().numpy()
        assert_close(got, [float(np.dot(vals, tvec(nrow)[rows]))], scale=2.0 * vals.sum(), what="x-blocked A'x, one column")
        # no entries at all
        M0 = fs.DeviceMatrix.from_coo_tensors(nrow, 5, z[:0], z[:0], None)
        assert not M0.spmm_t(x, 1).cpu().numpy().any()
    finally:
        fs.check(L.fsb_tune(b"t_xblock_min_kb", 0)); fs.check(L.fsb_tune(b"t_xblock_kb", 32 << 10))   # 0 = the built-in thresholds


@pytest.mark.parametrize("with_vals", [False, True])
def test_stream_kernel_tma_fed_form_is_bit_identical_to_the_per_thread_load_form(with_vals):
    """R = 1 merge-path kernel: the default brings a tile's run of indices / values into shared memory by TMA bulk copies
    that start at the enclosing 16-byte boundary (kernels_csr_stream.cu csr_stream_tma_kernel); knob stream_tma = 0 is the
    earlier per-thread-load form.  Both must give the same bits -- and the oracle's product -- on a matrix whose tiles
    start at every alignment (ragged rows, empty rows, a 9000-entry row spanning several tiles, a last tile of a few
    entries), through A x and the cached transpose."""
    import torch
    rng = np.random.default_rng(5 + with_vals)
    nrow, ncol, nnz = 7001, 1913, 70_003
    rows, cols, vals = _random_case(rng, nrow, ncol, nnz, long_row=9000)
    v = vals if with_vals else None
    rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, v)
    M = fs.DeviceMatrix.from_coo_tensors(nrow, ncol, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(),
                                         torch.from_numpy(vals).cuda() if with_vals else None)
    x = torch.from_numpy(tvec(ncol)).cuda(); xt = torch.from_numpy(tvec(nrow)).cuda()
    L = fs.lib()
    fs.check(L.fsb_tune_csr_algo(3, 0, 0))
    try:
        got = {}
        for form in (1, 0, 2):       # 2 = the TMA-fed form with LDG gathers instead of texture fetches (knob stream_tex)
            fs.check(L.fsb_tune(b"stream_tma", 1 if form else 0)); fs.check(L.fsb_tune(b"stream_tex", 0 if form == 2 else 1))
            got[form] = (M.spmm(x, 1).cpu().numpy(), M.spmm_t(xt, 1).cpu().numpy())
    finally:
        fs.check(L.fsb_tune(b"stream_tma", 1)); fs.check(L.fsb_tune(b"stream_tex", 1)); fs.check(L.fsb_tune_csr_algo(0, 0, 0))
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    assert np.array_equal(got[2][0], got[1][0]) and np.array_equal(got[2][1], got[1][1])
    want = oracle.csr_mul(nrow, rp, cc, vv, tvec(ncol).reshape(-1, 1), 1).reshape(-1)
    sc = np.abs(oracle.csr_mul(nrow, rp, cc, np.abs(vv) if vv is not None else None, np.abs(tvec(ncol)).reshape(-1, 1), 1)).reshape(-1)
    assert_close(got[1][0], want, sc, what="TMA-fed stream kernel vs oracle")
    want_t = oracle.coo_mul(nrow, rows, cols, v, tvec(nrow), transpose=True, ncol=ncol)
    assert_close(got[1][1], want_t, scale=2.0 * float(np.bincount(cols, minlength=ncol).max()), what="TMA-fed stream kernel, transpose, vs oracle")


@pytest.mark.parametrize("R", [1, 2, 3, 4, 8, 32, 100])
def test_staged_kernel_texture_gathers_give_the_same_bits(R):
    """Knob staged_tex: the staged kernel fetches the dense operand through a linear texture (8-byte texels for one
    double per lane, 16-byte texels for two or four) instead of LDG -- the default for binary matrices with R <= 4
    (kernels_csr_staged.cu staged_tex_auto).  Same bits as the LDG form on a ragged matrix with empty rows and a row
    long enough for the split path, with and without values, A x and the cached transpose."""
    import torch
    rng = np.random.default_rng(100 + R)
    nrow, ncol, nnz = 3001, 777, 40000
    rows, cols, vals = _random_case(rng, nrow, ncol, nnz, long_row=5000)
    L = fs.lib()
    X = torch.from_numpy(f64(rng.standard_normal((ncol, R)))).cuda().reshape(-1)
    Xt = torch.from_numpy(f64(rng.standard_normal((nrow, R)))).cuda().reshape(-1)
    for v in (None, vals):
        M = fs.DeviceMatrix.from_coo_tensors(nrow, ncol, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(),
                                             torch.from_numpy(v).cuda() if v is not None else None)
        fs.check(L.fsb_tune_csr_algo(2, 0, 0))
        try:
            got = {}
            for tex in (0, 1):
                fs.check(L.fsb_tune(b"staged_tex", tex))
                got[tex] = (M.spmm(X, R).cpu().numpy(), M.spmm_t(Xt, R).cpu().numpy())
        finally:
            fs.check(L.fsb_tune(b"staged_tex", -1)); fs.check(L.fsb_tune_csr_algo(0, 0, 0))
        assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1]), f"R={R} vals={v is not None}"
        rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, v)
        Xh = X.cpu().numpy().reshape(ncol, R)
        sc = np.abs(oracle.csr_mul(nrow, rp, cc, np.abs(vv) if vv is not None else None, np.abs(Xh), R)).reshape(-1)
        assert_close(got[1][0], oracle.csr_mul(nrow, rp, cc, vv, Xh, R), sc, what=f"texture gathers R={R} vals={v is not None}")


@pytest.mark.parametrize("offset", [1, 2, 37])
def test_texture_gathers_from_operands_off_the_512_byte_boundary(offset):
    """A linear texture must start on a 512-byte boundary; operands that do not (a row shard's slice of a larger vector)
    get a texture from the boundary below plus a texel offset (fsb_capi.cu fsb_linear_texture).  Products with the dense
    operand `offset` doubles into an allocation: texture and LDG forms agree to the bit, R = 1 (merge-path kernel,
    matrix with values) and R = 1, 2, 4 (staged kernel, binary)."""
    import torch
    rng = np.random.default_rng(offset)
    nrow, ncol, nnz = 4001, 1500, 50000
    rows, cols, vals = _random_case(rng, nrow, ncol, nnz, long_row=3000)
    L = fs.lib()
    Md = fs.DeviceMatrix.from_coo_tensors(nrow, ncol, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), torch.from_numpy(vals).cuda())
    Mb = fs.DeviceMatrix.from_coo_tensors(nrow, ncol, torch.from_numpy(rows).cuda(), torch.from_numpy(cols).cuda(), None)
    for M, R, knob in ((Md, 1, b"stream_tex"), (Mb, 1, b"staged_tex"), (Mb, 2, b"staged_tex"), (Mb, 4, b"staged_tex")):
        big = torch.from_numpy(f64(rng.standard_normal(ncol * R + 64))).cuda()
        X = big[offset:offset + ncol * R]
        assert X.data_ptr() % 512 != 0
        got = {}
        try:
            for tex in (1, 0):
                fs.check(L.fsb_tune(knob, tex))
                got[tex] = M.spmm(X, R).cpu().numpy()
        finally:
            fs.check(L.fsb_tune(b"stream_tex", 1)); fs.check(L.fsb_tune(b"staged_tex", -1))
        assert np.array_equal(got[0], got[1]), f"R={R} offset={offset}"
        rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, vals if M is Md else None)
        Xh = X.cpu().numpy().reshape(ncol, R)
        sc = np.abs(oracle.csr_mul(nrow, rp, cc, np.abs(vv) if vv is not None else None, np.abs(Xh), R)).reshape(-1)
        assert_close(got[1], oracle.csr_mul(nrow, rp, cc, vv, Xh, R), sc, what=f"offset operand R={R}")

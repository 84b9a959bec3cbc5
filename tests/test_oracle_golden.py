"""Pins the CPU oracle (oracle/fsoracle.c) against
  (1) the known answers in the reference's own test_sparse.c,
  (2) tests/golden/*.npz = outputs of the UNMODIFIED reference (oracle/gen_golden.py),
  (3) the compiled reference itself (oracle/_ref) on fresh random inputs, when present.
Structure/index work is compared bit-exactly; fp64 products to 1e-12 (the reference
is built with -ffast-math, so its own summation order is not fixed)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import O, REF, BlockedMatrix, dp, f64, i32, ip, lp
from conftest import DATA, assert_close, golden, rhs_matrix, test_vec as tvec


def make_sbm():  # test_sparse.c:16-28
    return 4, 3, i32([0, 3, 3, 1, 2]), i32([0, 2, 0, 2, 1])


def make_sdm():  # test_sparse.c:395-410
    return (6, 4, i32([1, 1, 3, 4, 1, 4, 5, 0, 1, 2, 4]), i32([0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 3]),
            f64([0.65, 0.84, 0.54, 0.59, 0.51, 0.27, 0.23, 0.94, 0.66, 0.31, 0.92]))


# ---------------------------------------------------------------- (1) known answers
def test_known_A_mul_B_4x3():  # test_sparse.c:30-44, 159-173
    nrow, ncol, r, c = make_sbm()
    y = oracle.coo_mul(nrow, r, c, None, [0.5, -0.7, 1.9])
    assert list(y) == [0.5, 1.9, -0.7, 2.4]
    z = oracle.coo_mul(nrow, r, c, None, [0.2, 1.3, -0.7, -0.5], transpose=True, ncol=ncol)
    assert list(z) == [-0.3, -0.7, 0.8]


def test_known_fixture_loader_and_products():  # test_sparse.c:46-77, 185-193
    nrow, ncol, r, c, _ = oracle.read_coo_file(os.path.join(DATA, "sbm-100-50.data"))
    assert (nrow, ncol, r.size, r[0], c[0]) == (100, 50, 504, 8, 0)
    x = tvec(ncol)
    y = oracle.coo_mul(nrow, r, c, None, x)
    assert abs(y[0] - 1.70095) < 1e-4 and abs(y[99] + 0.174905) < 1e-4
    rp, cc, _ = oracle.csr_from_coo(nrow, r, c)
    y2 = oracle.csr_mul(nrow, rp, cc, None, x, 1)
    assert np.max(np.abs(y - y2)) < 1e-12
    # SURVEY.md 8c extra known answers
    assert list(rp[:6]) == [0, 5, 7, 13, 19, 24] and list(cc[:8]) == [9, 27, 34, 41, 45, 1, 43, 1]
    assert abs(y2[0] - 1.7009531873605335) < 1e-14 and abs(y2.sum() - 42.631240828393558) < 1e-12


def test_known_sdm():  # test_sparse.c:412-482
    nrow, ncol, r, c, v = make_sdm()
    yt = [2.162, 2.224, 0.713, -0.378, 2.216, 0.437]
    y = oracle.coo_mul(nrow, r, c, v, [0.5, -0.7, 1.9, 2.3])
    assert np.max(np.abs(y - yt)) < 1e-6
    rp, cc, vv = oracle.csr_from_coo(nrow, r, c, v)
    assert np.max(np.abs(oracle.csr_mul(nrow, rp, cc, vv, [0.5, -0.7, 1.9, 2.3], 1) - yt)) < 1e-6
    Y = oracle.csr_mul(nrow, rp, cc, vv, [0.5, 5.0, -0.7, -7.0, 1.9, 19.0, 2.3, 23.0], 2)
    assert np.max(np.abs(Y[:, 0] - yt)) < 1e-6 and np.max(np.abs(Y[:, 1] - 10 * np.array(yt))) < 1e-6
    z = oracle.coo_mul(nrow, r, c, v, [0.59, 0.37, 0.14, 0.21, 0.40, 0.81], transpose=True, ncol=ncol)
    assert np.max(np.abs(z - [0.2405, 0.6602, 0.483, 1.2102])) < 1e-6
    n2, c2, r2, cc2, v2 = oracle.read_coo_file(os.path.join(DATA, "sdm-100-50.data"), with_vals=True)
    assert (n2, c2, r2.size, r2[1], cc2[1], r2[469], cc2[469]) == (100, 50, 470, 27, 0, 40, 49)
    assert abs(v2[1] - 0.616153) < 1e-5 and abs(v2[469] - 0.108172) < 1e-5


def test_known_cbcsr_and_blocking():  # test_sparse.c:114-157, 293-301
    nrow, ncol, r, c = make_sbm()
    nb, rp, cc = oracle.cbcsr_from_coo(nrow, ncol, 2, r, c)
    assert nb == 2
    y = np.zeros(nrow); O.fso_cbcsr_A_mul_B(dp(y), nrow, nb, ip(rp), ip(cc), dp(f64([0.5, -0.7, 1.9])))
    assert list(y) == [0.5, 1.9, -0.7, 2.4]
    nrow, ncol, r, c, _ = oracle.read_coo_file(os.path.join(DATA, "sbm-100-50.data"))
    B = oracle.blocked_from_coo(nrow, ncol, 8, r, c)
    assert (B.nblocks, B.start_row[0], B.start_row[1], B.start_row[13]) == (13, 0, 8, 100)


def test_known_hilbert():  # test_sparse.c:195-203, 250-265, 347-361
    for x, want in [(16, 16), (15, 16), (17, 32), (1, 1), (1 << 30, 1 << 30), ((1 << 30) - 1, 1 << 30)]:
        assert O.fso_ceil_pow2(x) == want
    h = O.fso_xy2d(131072, 5931, 91204)
    a, b = C.c_int(), C.c_int(); O.fso_d2xy(131072, h, C.byref(a), C.byref(b))
    assert (a.value, b.value) == (5931, 91204)
    assert [O.fso_row_xy2d(16, *p) for p in [(0, 0), (0, 15), (0, 16), (0, 31), (1, 0)]] == [0, 255, 256, 511, 3]
    for d, want in [(0, (0, 0)), (255, (0, 15)), (256, (0, 16)), (511, (0, 31)), (3, (1, 0))]:
        O.fso_row_d2xy(16, d, C.byref(a), C.byref(b)); assert (a.value, b.value) == want


def test_known_linalg():  # test_sparse.c:511-558
    x = f64([0.12, -0.82, 1.3, 0.5]); y = f64([6.12, 0.19, 3.4, -4.1])
    assert abs(O.fso_normsq(dp(x), 4) - 2.6268) < 1e-8 and abs(O.fso_normsq(dp(y), 4) - 65.8605) < 1e-8
    assert abs(O.fso_dot(dp(x), dp(y), 4) - 2.9486) < 1e-8
    X = f64([0.95, 0.9, 0.16, 0.46, 0.86, 0.29]); Y = f64([0.9695, 0.6678, 0.277, 0.1908, 1.108, 0.7632])
    o = np.zeros(3); O.fso_dot2sym(dp(o), dp(X), dp(Y), 3)
    assert np.max(np.abs(o - [1.918225, 0.910116, 1.32129])) < 1e-8
    S = np.zeros(4); O.fso_solve2sym(dp(S), dp(f64([0.59, 1.34, 0.86])), dp(f64([-1.21, 1.91, -0.82, 0.03])))
    assert np.max(np.abs(S - [-64.0, 42.5, -22.05098039, 14.1745098])) < 1e-8


def test_known_cg():  # test_sparse.c:560-608 + SURVEY.md 8c
    nrow, ncol, r, c, _ = oracle.read_coo_file(os.path.join(DATA, "sbm-100-50.data"))
    r0, c0 = r.copy(), c.copy()
    O.fso_sort_coo_hilbert(nrow, ncol, r.size, ip(r), ip(c), None)
    A = oracle.blocked_from_coo(nrow, ncol, 8, r, c)
    At = oracle.blocked_from_coo(ncol, nrow, 8, c, r)
    b = tvec(ncol); x = np.zeros(ncol)
    it = O.fso_blocked_cg(dp(x), A.ref(), At.ref(), dp(b), 5.0, 1e-6)
    assert it == 15 and abs(x[0] - 0.0638578) < 1e-4 and abs(x[1] + 0.0302702) < 1e-4
    assert abs(x[49] + 0.0284737361861) < 1e-10
    resid = oracle.coo_mul(nrow, r0, c0, None, oracle.coo_mul(nrow, r0, c0, None, x), transpose=True, ncol=ncol) + 5.0 * x - b
    assert np.linalg.norm(resid) < 1e-5
    i = np.arange(ncol, dtype=np.int64)
    B2 = f64(np.stack([b, np.cos(i * 23 + 0.7) + np.sin(i * i * 7)], 1)); X2 = np.zeros((ncol, 2))
    it2 = O.fso_blocked_cg2(dp(X2), A.ref(), At.ref(), dp(B2), 5.0, 1e-6)
    assert it2 == 13 and abs(X2[0, 0] - 0.0638578) < 1e-4 and abs(X2[1, 0] + 0.0302702) < 1e-4
    assert abs(X2[0, 1] - 0.106690812806) < 1e-9 and abs(X2[1, 1] - 0.121147589909) < 1e-9


# ---------------------------------------------------------------- (2) golden fixtures
def _blocked_eq(B, g, prefix):
    assert np.array_equal(B.start_row, g[prefix + "start_row"])
    assert np.array_equal(B.blk_nnz[:B.nblocks], g[prefix + "blk_nnz"])
    assert np.array_equal(B.rows[:B.nnz], g[prefix + "rows"]) and np.array_equal(B.cols[:B.nnz], g[prefix + "cols"])
    if B.vals is not None:
        assert np.array_equal(B.vals[:B.nnz], g[prefix + "vals"])


@pytest.mark.parametrize("name", ["sbm_100_50", "rand_bin_300_70"])
def test_golden_binary(name, tmp_path):
    g = golden(name)
    nrow, ncol, rows, cols = int(g["nrow"]), int(g["ncol"]), g["rows"], g["cols"]
    x, xt = g["x"], g["xt"]
    sc = 2.0 * max(1, np.bincount(rows, minlength=nrow).max())
    assert_close(oracle.coo_mul(nrow, rows, cols, None, x), g["coo_Ax"], sc, what="coo_Ax")
    sct = 2.0 * np.bincount(cols, minlength=ncol).max()
    assert_close(oracle.coo_mul(nrow, rows, cols, None, xt, transpose=True, ncol=ncol), g["coo_Atx"], sct, what="coo_Atx")
    rp, cc, _ = oracle.csr_from_coo(nrow, rows, cols)
    assert np.array_equal(rp, g["csr_row_ptr"]) and np.array_equal(cc, g["csr_cols"])
    assert_close(oracle.csr_mul(nrow, rp, cc, None, x, 1), g["csr_Ax"], sc, what="csr_Ax")
    for R in g["Rs"]:
        R = int(R)
        Y = oracle.csr_mul(nrow, rp, cc, None, g[f"X{R}"], R)
        for key in (f"csr_AX{R}_Bn", f"csr_AX{R}_B32n", f"csr_AX{R}_fixed", "csr_AX8_auto" if R == 8 else "-"):
            if key in g.files:
                assert_close(Y, g[key], sc, what=key)
    z = np.zeros(ncol); O.fso_bcsr_AA_mul_B(dp(z), nrow, ncol, ip(rp), ip(cc), dp(x))
    assert_close(z, g["csr_AAx"], sc * sct, what="AAx"); assert_close(z, g["csr_AAx_par"], sc * sct, what="AAx_par")
    nb, crp, ccc = oracle.cbcsr_from_coo(nrow, ncol, int(g["colblock"]), rows, cols)
    assert nb == int(g["cb_nblocks"]) and np.array_equal(crp, g["cb_row_ptr"]) and np.array_equal(ccc, g["cb_cols"])
    y = np.zeros(nrow); O.fso_cbcsr_A_mul_B(dp(y), nrow, nb, ip(crp), ip(ccc), dp(x)); assert_close(y, g["cb_Ax"], sc, what="cb_Ax")
    hr, hc = rows.copy(), cols.copy(); O.fso_sort_coo_hilbert(nrow, ncol, rows.size, ip(hr), ip(hc), None)
    assert np.array_equal(hr, g["hil_rows"]) and np.array_equal(hc, g["hil_cols"])
    bs = int(g["bs"])
    B = oracle.blocked_from_coo(nrow, ncol, bs, rows, cols); _blocked_eq(B, g, "blk_")
    Bh = B.copy(); O.fso_sort_blocked_hilbert(Bh.ref()); _blocked_eq(Bh, g, "blkh_")
    Br = B.copy(); O.fso_sort_blocked_byrow(Br.ref()); _blocked_eq(Br, g, "blkr_")
    assert_close(oracle.blocked_mul(Bh, x, 1), g["blkh_Ax"], sc, what="blkh_Ax")
    for R in g["Rs"]:
        R = int(R)
        Y = oracle.blocked_mul(Bh, g[f"X{R}"], R)
        for key in (f"blkh_AX{R}_Bn", f"blkh_AX{R}_fixed"):
            if key in g.files:
                assert_close(Y, g[key], sc, what=key)
    # solver
    lam = float(g["cg_lambda"])
    A = oracle.blocked_from_coo(nrow, ncol, bs, hr, hc); At = oracle.blocked_from_coo(ncol, nrow, bs, hc, hr)
    yy = np.zeros(ncol); tmp = np.zeros(nrow)
    O.fso_blocked_AtA(dp(yy), A.ref(), At.ref(), dp(f64(g["cg_b"])), dp(tmp), lam)
    assert_close(yy, g["AtA_b"], sc * sct, what="AtA")
    xs = np.zeros(ncol); it = O.fso_blocked_cg(dp(xs), A.ref(), At.ref(), dp(f64(g["cg_b"])), lam, 1e-6)
    # iteration count within +-1 of the reference (its reductions run under -ffast-math/OpenMP,
    # so a borderline convergence test may flip), solution within 10*tol relative
    assert abs(it - int(g["cg_iter"])) <= 1; assert np.max(np.abs(xs - g["cg_x"])) <= 1e-5 * np.max(np.abs(g["cg_x"]))
    X2 = np.zeros((ncol, 2)); it2 = O.fso_blocked_cg2(dp(X2), A.ref(), At.ref(), dp(f64(g["cg2_B"])), lam, 1e-6)
    assert abs(it2 - int(g["cg2_iter"])) <= 1; assert np.max(np.abs(X2 - g["cg2_X"])) <= 1e-5 * np.max(np.abs(g["cg2_X"]))
    if name == "sbm_100_50":                      # the reference's own fixture: exact iteration counts
        assert (it, it2) == (15, 13) and np.max(np.abs(xs - g["cg_x"])) < 1e-10 and np.max(np.abs(X2 - g["cg2_X"])) < 1e-10
    # .csr.bin: identical bytes except the 16 bytes of stale host pointers inside the struct image
    p = str(tmp_path / "m.csr.bin")
    assert O.fso_write_csr_bin(p.encode(), nrow, ncol, rows.size, ip(rp), ip(cc)) == 0
    mine = np.frombuffer(open(p, "rb").read(), dtype=np.uint8); ref = g["csr_bin"]
    assert mine.size == ref.size
    ptr0 = 50 + 17 + 16
    keep = np.ones(mine.size, bool); keep[ptr0:ptr0 + 16] = False
    assert np.array_equal(mine[keep], ref[keep])
    if name == "sbm_100_50":
        assert mine.size == 2537
    n_, c_, z_ = C.c_int(), C.c_int(), C.c_long()
    rp2 = np.zeros_like(rp); cc2 = np.zeros_like(cc)
    open(p, "wb").write(ref.tobytes())      # the reference-written file must load through the oracle reader
    assert O.fso_read_csr_bin(p.encode(), C.byref(n_), C.byref(c_), C.byref(z_), ip(rp2), ip(cc2)) == 0
    assert (n_.value, c_.value, z_.value) == (nrow, ncol, rows.size) and np.array_equal(rp2, rp) and np.array_equal(cc2, cc)


@pytest.mark.parametrize("name", ["sdm_100_50", "rand_dbl_257_129"])
def test_golden_double(name):
    g = golden(name)
    nrow, ncol, rows, cols, vals = int(g["nrow"]), int(g["ncol"]), g["rows"], g["cols"], g["vals"]
    x, xt = g["x"], g["xt"]
    sc = 2.0 * max(1, np.bincount(rows, minlength=nrow).max())
    assert_close(oracle.coo_mul(nrow, rows, cols, vals, x), g["coo_Ax"], sc, what="coo_Ax")
    assert_close(oracle.coo_mul(nrow, rows, cols, vals, xt, transpose=True, ncol=ncol), g["coo_Atx"],
                 2.0 * np.bincount(cols, minlength=ncol).max(), what="coo_Atx")
    rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, vals)
    assert np.array_equal(rp, g["csr_row_ptr"]) and np.array_equal(cc, g["csr_cols"]) and np.array_equal(vv, g["csr_vals"])
    assert_close(oracle.csr_mul(nrow, rp, cc, vv, x, 1), g["csr_Ax"], sc, what="csr_Ax")
    for R in g["Rs"]:
        R = int(R)
        assert_close(oracle.csr_mul(nrow, rp, cc, vv, g[f"X{R}"], R), g[f"csr_AX{R}_Bn"], sc, what=f"AX{R}")
    hr, hc, hv = rows.copy(), cols.copy(), vals.copy()
    O.fso_sort_coo_hilbert(nrow, ncol, rows.size, ip(hr), ip(hc), dp(hv))
    assert np.array_equal(hr, g["hil_rows"]) and np.array_equal(hc, g["hil_cols"]) and np.array_equal(hv, g["hil_vals"])
    B = oracle.blocked_from_coo(nrow, ncol, int(g["bs"]), rows, cols, vals); _blocked_eq(B, g, "blk_")
    Bh = B.copy(); O.fso_sort_blocked_hilbert(Bh.ref()); _blocked_eq(Bh, g, "blkh_")
    assert_close(oracle.blocked_mul(Bh, x, 1), g["blkh_Ax"], sc, what="blkh_Ax")


def test_golden_hilbert_and_sort():
    g = golden("hilbert")
    assert [O.fso_ceil_pow2(int(v)) for v in g["cp2_in"]] == list(g["cp2_out"])
    a, b = C.c_int(), C.c_int()
    for n in (1, 2, 16, 128, 131072, 1 << 24):
        x, y, d = g[f"xy_{n}_x"], g[f"xy_{n}_y"], g[f"xy_{n}_d"]
        for j in range(x.size):
            assert O.fso_xy2d(n, int(x[j]), int(y[j])) == d[j]
            O.fso_d2xy(n, int(d[j]), C.byref(a), C.byref(b)); assert (a.value, b.value) == (g[f"xy_{n}_bx"][j], g[f"xy_{n}_by"][j])
            assert O.fso_row_xy2d(n, int(x[j]), int(g[f"rxy_{n}_y"][j])) == g[f"rxy_{n}_d"][j]
            O.fso_row_d2xy(n, int(g[f"rxy_{n}_d"][j]), C.byref(a), C.byref(b))
            assert (a.value, b.value) == (g[f"rxy_{n}_bx"][j], g[f"rxy_{n}_by"][j])
    k, p = g["qs_keys"].copy(), g["qs_pay"].copy(); O.fso_sort_keys_vals(lp(k), dp(p), k.size)
    assert np.array_equal(k, g["qs_keys_out"]) and np.array_equal(p, g["qs_pay_out"])


def test_golden_linalg():
    g = golden("linalg"); n = g["x"].size
    x, y, X, Y = f64(g["x"]), f64(g["y"]), f64(g["X"]), f64(g["Y"])
    assert abs(O.fso_dist(dp(x), dp(y), n) - g["dist"]) < 1e-11 and abs(O.fso_normsq(dp(x), n) - g["normsq"]) < 1e-10
    assert abs(O.fso_dot(dp(x), dp(y), n) - g["dot"]) < 1e-10
    o = np.zeros(2); O.fso_normsq2(dp(o), dp(X), n); assert np.max(np.abs(o - g["normsq2"])) < 1e-10
    o = np.zeros(3); O.fso_outer2(dp(o), dp(X), n); assert np.max(np.abs(o - g["outer2"])) < 1e-10
    o = np.zeros(3); O.fso_dot2sym(dp(o), dp(X), dp(Y), n); assert np.max(np.abs(o - g["dot2sym"])) < 1e-10
    S = np.zeros(4); O.fso_solve2sym(dp(S), dp(f64(g["s2_A"])), dp(f64(g["s2_RHS"]))); assert np.max(np.abs(S - g["s2_X"])) < 1e-12


# ---------------------------------------------------------------- (3) live reference
@pytest.mark.skipif(REF is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_vs_live_reference(seed):
    rng = np.random.default_rng(seed)
    nrow, ncol, nnz = int(rng.integers(1, 400)), int(rng.integers(1, 300)), int(rng.integers(0, 5000))
    rows = rng.integers(0, nrow, nnz, dtype=np.int32); cols = rng.integers(0, ncol, nnz, dtype=np.int32)
    vals = rng.random(nnz)
    rp, cc, vv = oracle.csr_from_coo(nrow, rows, cols, vals)
    rrp = np.zeros(nrow + 1, np.int32); rcc = np.zeros(max(nnz, 1), np.int32); rvv = np.zeros(max(nnz, 1))
    REF.ref_new_csr(nnz, nrow, ncol, ip(rows), ip(cols), dp(vals), ip(rrp), ip(rcc), dp(rvv))
    assert np.array_equal(rp, rrp) and np.array_equal(cc, rcc[:nnz]) and np.array_equal(vv, rvv[:nnz])
    R = int(rng.integers(1, 33)); X = f64(rng.standard_normal((ncol, R)))
    Y = np.zeros((nrow, R)); REF.ref_bcsr_mul(32, dp(Y), nrow, ncol, nnz, ip(rrp), ip(rcc), dp(X), R)
    assert_close(oracle.csr_mul(nrow, rp, cc, None, X, R), Y, 50.0, what="B32n")
    Y = np.zeros((nrow, R)); REF.ref_csr_mul(0, dp(Y), nrow, ncol, nnz, ip(rrp), ip(rcc), dp(rvv), dp(X), R)
    assert_close(oracle.csr_mul(nrow, rp, cc, vv, X, R), Y, 50.0, what="csr_Bn")
    bs = int(rng.integers(1, 64))
    B = oracle.blocked_from_coo(nrow, ncol, bs, rows, cols); O.fso_sort_blocked_hilbert(B.ref())
    Bref = BlockedMatrix(nrow, ncol, nnz, bs, False)
    REF.ref_new_bsbm(nnz, nrow, ncol, ip(rows), ip(cols), bs, Bref.ref()); REF.ref_sort_bsbm(Bref.ref())
    assert np.array_equal(B.rows, Bref.rows) and np.array_equal(B.cols, Bref.cols) and np.array_equal(B.blk_nnz, Bref.blk_nnz)


def test_oracle_synth_generator_matches_the_product_generator_bit_for_bit():
    """bench.py's reference arm draws its inputs from oracle/ (fso_synth_coo) so that it never loads the product
    library; both generators must emit the identical COO stream, for both column distributions and any offset."""
    import libfastsparse_b200 as fs
    for dist, nrow, ncol in [(0, 100000, 7919), (1, 5000, 100003), (1, 77, 1), (0, 1, 1)]:
        n = 200000
        r1, c1, v1 = fs.synth_coo_host(0x5EED0002 + dist, dist, n, nrow, ncol, with_vals=True)
        r2, c2, v2 = oracle.synth_coo(0x5EED0002 + dist, dist, n, nrow, ncol, with_vals=True)
        assert np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.array_equal(v1, v2)
        r3, c3, _ = oracle.synth_coo(0x5EED0002 + dist, dist, 1000, nrow, ncol, j0=12345)
        assert np.array_equal(r3, r1[12345:13345]) and np.array_equal(c3, c1[12345:13345])

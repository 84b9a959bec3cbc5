"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists):
    python -m oracle.gen_golden
Every array below is an output of the reference's own code (via
oracle/ref_wrap.c) on the stated inputs; inputs are stored next to the outputs
so the fixtures are self-contained on the GPU box.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

import oracle
from oracle import REF, BlockedMatrix, dp, f64, i32, ip, lp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(OUT, "data")


def test_vec(n):
    """x[i] = sin(i*19 + 0.4) + cos(i*i*3)  (test_sparse.c:51)"""
    i = np.arange(n, dtype=np.int64)
    return np.sin(i * 19 + 0.4) + np.cos(i * i * 3)


def ref_read_sbm(path):
    nrow, ncol, nnz = C.c_long(), C.c_long(), C.c_long()
    REF.ref_read_sbm(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), None, None)
    rows = np.zeros(nnz.value, np.int32); cols = np.zeros(nnz.value, np.int32)
    REF.ref_read_sbm(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), ip(rows), ip(cols))
    return nrow.value, ncol.value, rows, cols


def ref_read_sdm(path):
    nrow, ncol, nnz = C.c_long(), C.c_long(), C.c_long()
    REF.ref_read_sdm(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), None, None, None)
    rows = np.zeros(nnz.value, np.int32); cols = np.zeros(nnz.value, np.int32); vals = np.zeros(nnz.value)
    REF.ref_read_sdm(path.encode(), C.byref(nrow), C.byref(ncol), C.byref(nnz), ip(rows), ip(cols), dp(vals))
    return nrow.value, ncol.value, rows, cols, vals


def ref_blocked(nrow, ncol, rows, cols, vals, bs):
    B = BlockedMatrix(nrow, ncol, rows.size, bs, vals is not None)
    if vals is None:
        REF.ref_new_bsbm(rows.size, nrow, ncol, ip(rows), ip(cols), bs, B.ref())
    else:
        REF.ref_new_bsdm(rows.size, nrow, ncol, ip(rows), ip(cols), dp(vals), bs, B.ref())
    return B


def pack_blocked(prefix, B, d):
    d[prefix + "start_row"] = B.start_row.copy(); d[prefix + "blk_nnz"] = B.blk_nnz[:B.nblocks].copy()
    d[prefix + "rows"] = B.rows[:B.nnz].copy(); d[prefix + "cols"] = B.cols[:B.nnz].copy()
    if B.vals is not None:
        d[prefix + "vals"] = B.vals[:B.nnz].copy()


def binary_case(nrow, ncol, rows, cols, colblock, bs, Rs, lam=None, rng=None):
    """All binary-matrix reference outputs for one COO input."""
    rows, cols = i32(rows), i32(cols)
    nnz = rows.size
    d = dict(nrow=nrow, ncol=ncol, rows=rows, cols=cols, colblock=colblock, bs=bs, Rs=np.array(Rs))
    x = test_vec(ncol); xt = test_vec(nrow)
    d["x"] = x; d["xt"] = xt
    # COO products
    y = np.zeros(nrow); REF.ref_A_mul_B(dp(y), nrow, ncol, nnz, ip(rows), ip(cols), dp(x)); d["coo_Ax"] = y
    z = np.zeros(ncol); REF.ref_At_mul_B(dp(z), nrow, ncol, nnz, ip(rows), ip(cols), dp(xt)); d["coo_Atx"] = z
    # CSR build + products
    rp = np.zeros(nrow + 1, np.int32); cc = np.zeros(max(nnz, 1), np.int32)
    REF.ref_new_bcsr(nnz, nrow, ncol, ip(rows), ip(cols), ip(rp), ip(cc))
    d["csr_row_ptr"] = rp; d["csr_cols"] = cc[:nnz]
    y = np.zeros(nrow); REF.ref_bcsr_mul(1, dp(y), nrow, ncol, nnz, ip(rp), ip(cc), dp(x), 1); d["csr_Ax"] = y
    for R in Rs:
        k = np.arange(R)
        X = np.sin(7.0 * np.arange(ncol)[:, None] + 17.0 * k[None, :] + 0.3)  # bench_a_mul_b.c:149-152
        X = f64(X)
        d[f"X{R}"] = X
        Y = np.zeros((nrow, R)); REF.ref_bcsr_mul(0, dp(Y), nrow, ncol, nnz, ip(rp), ip(cc), dp(X), R); d[f"csr_AX{R}_Bn"] = Y
        if R <= 32:
            Y = np.zeros((nrow, R)); REF.ref_bcsr_mul(32, dp(Y), nrow, ncol, nnz, ip(rp), ip(cc), dp(X), R); d[f"csr_AX{R}_B32n"] = Y
        if R in (2, 4, 8):
            Y = np.zeros((nrow, R)); REF.ref_bcsr_mul(R, dp(Y), nrow, ncol, nnz, ip(rp), ip(cc), dp(X), R); d[f"csr_AX{R}_fixed"] = Y
        if R == 8:
            Y = np.zeros((nrow, R)); REF.ref_bcsr_mul(80, dp(Y), nrow, ncol, nnz, ip(rp), ip(cc), dp(X), R); d["csr_AX8_auto"] = Y
    z = np.zeros(ncol); REF.ref_bcsr_AA_mul_B(dp(z), nrow, ncol, nnz, ip(rp), ip(cc), dp(x)); d["csr_AAx"] = z
    z = np.zeros(ncol); REF.ref_parallel_bcsr_AA_mul_B(dp(z), nrow, ncol, nnz, ip(rp), ip(cc), dp(x)); d["csr_AAx_par"] = z
    # column-blocked CSR
    nb = REF.ref_new_cbcsr(colblock, nnz, nrow, ncol, ip(rows), ip(cols), None, None)
    crp = np.zeros(nb * nrow + 1, np.int32); ccc = np.zeros(max(nnz, 1), np.int32)
    REF.ref_new_cbcsr(colblock, nnz, nrow, ncol, ip(rows), ip(cols), ip(crp), ip(ccc))
    d["cb_nblocks"] = nb; d["cb_row_ptr"] = crp; d["cb_cols"] = ccc[:nnz]
    y = np.zeros(nrow); REF.ref_cbcsr_A_mul_B(dp(y), nrow, ncol, nb, colblock, nnz, ip(crp), ip(ccc), dp(x)); d["cb_Ax"] = y
    # Hilbert sort of the COO
    hr, hc = rows.copy(), cols.copy()
    REF.ref_sort_sbm(nrow, ncol, nnz, ip(hr), ip(hc)); d["hil_rows"] = hr; d["hil_cols"] = hc
    # row-blocked COO (unsorted, Hilbert-sorted, row-sorted) + products
    B = ref_blocked(nrow, ncol, rows, cols, None, bs); pack_blocked("blk_", B, d)
    Bh = B.copy(); REF.ref_sort_bsbm(Bh.ref()); pack_blocked("blkh_", Bh, d)
    Br = B.copy(); REF.ref_sort_bsbm_byrow(Br.ref()); pack_blocked("blkr_", Br, d)
    y = np.zeros(nrow); REF.ref_bsbm_mul(1, dp(y), Bh.ref(), dp(x), 1); d["blkh_Ax"] = y
    for R in Rs:
        X = d[f"X{R}"]
        Y = np.zeros((nrow, R)); REF.ref_bsbm_mul(0, dp(Y), Bh.ref(), dp(X), R); d[f"blkh_AX{R}_Bn"] = Y
        if R in (2, 4):
            Y = np.zeros((nrow, R)); REF.ref_bsbm_mul(R, dp(Y), Bh.ref(), dp(X), R); d[f"blkh_AX{R}_fixed"] = Y
    # solver on (Hilbert-sorted COO -> blocked, transpose -> blocked), as test_sparse.c:560-608
    if lam is not None:
        A = ref_blocked(nrow, ncol, hr, hc, None, bs)
        At = ref_blocked(ncol, nrow, hc, hr, None, bs)
        b = test_vec(ncol)
        tmp = np.zeros(nrow); yy = np.zeros(ncol)
        REF.ref_bsbm_AtA(dp(yy), A.ref(), At.ref(), dp(b), dp(tmp), lam); d["AtA_b"] = yy
        xs = np.zeros(ncol); it = REF.ref_bsbm_cg(dp(xs), A.ref(), At.ref(), dp(b), lam, 1e-6)
        d["cg_b"] = b; d["cg_x"] = xs; d["cg_iter"] = it; d["cg_lambda"] = lam
        i = np.arange(ncol, dtype=np.int64)
        b2 = np.stack([np.sin(i * 19 + 0.4) + np.cos(i * i * 3), np.cos(i * 23 + 0.7) + np.sin(i * i * 7)], 1)
        b2 = f64(b2)
        X2 = np.zeros((ncol, 2)); it2 = REF.ref_bsbm_cg2(dp(X2), A.ref(), At.ref(), dp(b2), lam, 1e-6)
        d["cg2_B"] = b2; d["cg2_X"] = X2; d["cg2_iter"] = it2
    # .csr.bin bytes
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "m.csr.bin")
        REF.ref_serialize_to_file(p.encode(), nrow, ncol, nnz, ip(rp), ip(cc))
        d["csr_bin"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
    return d


def double_case(nrow, ncol, rows, cols, vals, bs, Rs):
    rows, cols, vals = i32(rows), i32(cols), f64(vals)
    nnz = rows.size
    d = dict(nrow=nrow, ncol=ncol, rows=rows, cols=cols, vals=vals, bs=bs, Rs=np.array(Rs))
    x = test_vec(ncol); xt = test_vec(nrow); d["x"] = x; d["xt"] = xt
    y = np.zeros(nrow); REF.ref_sdm_A_mul_B(dp(y), nrow, ncol, nnz, ip(rows), ip(cols), dp(vals), dp(x)); d["coo_Ax"] = y
    z = np.zeros(ncol); REF.ref_sdm_At_mul_B(dp(z), nrow, ncol, nnz, ip(rows), ip(cols), dp(vals), dp(xt)); d["coo_Atx"] = z
    rp = np.zeros(nrow + 1, np.int32); cc = np.zeros(max(nnz, 1), np.int32); vv = np.zeros(max(nnz, 1))
    REF.ref_new_csr(nnz, nrow, ncol, ip(rows), ip(cols), dp(vals), ip(rp), ip(cc), dp(vv))
    d["csr_row_ptr"] = rp; d["csr_cols"] = cc[:nnz]; d["csr_vals"] = vv[:nnz]
    y = np.zeros(nrow); REF.ref_csr_mul(1, dp(y), nrow, ncol, nnz, ip(rp), ip(cc), dp(vv), dp(x), 1); d["csr_Ax"] = y
    for R in Rs:
        k = np.arange(R)
        X = f64(np.sin(7.0 * np.arange(ncol)[:, None] + 17.0 * k[None, :] + 0.3))
        d[f"X{R}"] = X
        Y = np.zeros((nrow, R)); REF.ref_csr_mul(0, dp(Y), nrow, ncol, nnz, ip(rp), ip(cc), dp(vv), dp(X), R); d[f"csr_AX{R}_Bn"] = Y
    hr, hc, hv = rows.copy(), cols.copy(), vals.copy()
    REF.ref_sort_sdm(nrow, ncol, nnz, ip(hr), ip(hc), dp(hv)); d["hil_rows"] = hr; d["hil_cols"] = hc; d["hil_vals"] = hv
    B = ref_blocked(nrow, ncol, rows, cols, vals, bs); pack_blocked("blk_", B, d)
    Bh = B.copy(); REF.ref_sort_bsdm(Bh.ref()); pack_blocked("blkh_", Bh, d)
    y = np.zeros(nrow); REF.ref_bsdm_A_mul_B(dp(y), Bh.ref(), dp(x)); d["blkh_Ax"] = y
    return d


def random_coo(rng, nrow, ncol, nnz, empty_frac=0.2, dup_frac=0.1):
    live = rng.permutation(nrow)[: max(1, int(nrow * (1 - empty_frac)))]
    rows = rng.choice(live, size=nnz).astype(np.int32)
    cols = rng.integers(0, ncol, size=nnz, dtype=np.int32)
    ndup = int(nnz * dup_frac)
    src = rng.integers(0, nnz, size=ndup)
    dst = rng.integers(0, nnz, size=ndup)
    rows[dst] = rows[src]; cols[dst] = cols[src]          # exact duplicate coordinates
    return rows, cols


def hilbert_case(rng):
    d = {}
    xs = np.array([1, 2, 3, 15, 16, 17, 100, 1000, 65535, 65536, 65537, (1 << 24) - 1, (1 << 24) + 1,
                   (1 << 30) - 1, 1 << 30], np.int32)
    d["cp2_in"] = xs; d["cp2_out"] = np.array([REF.ref_ceilPower2(int(v)) for v in xs], np.int32)
    for n in (1, 2, 16, 128, 131072, 1 << 24):
        m = 257
        x = rng.integers(0, n, size=m, dtype=np.int32); y = rng.integers(0, n, size=m, dtype=np.int32)
        dd = np.array([REF.ref_xy2d(n, int(a), int(b)) for a, b in zip(x, y)], np.int64)
        bx = np.zeros(m, np.int32); by = np.zeros(m, np.int32)
        for j in range(m):
            a, b = C.c_int(), C.c_int()
            REF.ref_d2xy(n, int(dd[j]), C.byref(a), C.byref(b)); bx[j] = a.value; by[j] = b.value
        d[f"xy_{n}_x"] = x; d[f"xy_{n}_y"] = y; d[f"xy_{n}_d"] = dd; d[f"xy_{n}_bx"] = bx; d[f"xy_{n}_by"] = by
        # row-strip variant: x in [0,n), y anywhere in [0, 2^20)
        yy = rng.integers(0, 1 << 20, size=m, dtype=np.int32)
        rd = np.array([REF.ref_row_xy2d(n, int(a), int(b)) for a, b in zip(x, yy)], np.int64)
        rx = np.zeros(m, np.int32); ry = np.zeros(m, np.int32)
        for j in range(m):
            a, b = C.c_int(), C.c_int()
            REF.ref_row_d2xy(n, int(rd[j]), C.byref(a), C.byref(b)); rx[j] = a.value; ry[j] = b.value
        d[f"rxy_{n}_y"] = yy; d[f"rxy_{n}_d"] = rd; d[f"rxy_{n}_bx"] = rx; d[f"rxy_{n}_by"] = ry
    # quicksort with payload, duplicates included (placement of payloads under equal keys)
    keys = rng.integers(0, 50, size=500).astype(np.int64); pay = rng.random(500)
    k2, p2 = keys.copy(), pay.copy(); REF.ref_quickSortD(lp(k2), dp(p2), 500)
    d["qs_keys"] = keys; d["qs_pay"] = pay; d["qs_keys_out"] = k2; d["qs_pay_out"] = p2
    return d


def linalg_case(rng):
    d = {}
    n = 1001
    x = rng.standard_normal(n); y = rng.standard_normal(n); X = rng.standard_normal((n, 2)); Y = rng.standard_normal((n, 2))
    d.update(x=x, y=y, X=X, Y=Y)
    d["dist"] = REF.ref_dist(dp(x), dp(y), n); d["normsq"] = REF.ref_pnormsq(dp(x), n); d["dot"] = REF.ref_pdot(dp(x), dp(y), n)
    o = np.zeros(2); REF.ref_pnormsq2(dp(o), dp(X), n); d["normsq2"] = o
    o = np.zeros(3); REF.ref_pouter2(dp(o), dp(X), n); d["outer2"] = o
    o = np.zeros(3); REF.ref_pdot2sym(dp(o), dp(X), dp(Y), n); d["dot2sym"] = o
    A = np.array([0.59, 1.34, 0.86]); RHS = np.array([-1.21, 1.91, -0.82, 0.03]); S = np.zeros(4)
    REF.ref_solve2sym(dp(S), dp(A), dp(RHS)); d["s2_A"] = A; d["s2_RHS"] = RHS; d["s2_X"] = S
    return d


def main():
    assert REF is not None, "oracle/_ref not built (need /root/reference): make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    nrow, ncol, rows, cols = ref_read_sbm(os.path.join(DATA, "sbm-100-50.data"))
    np.savez_compressed(os.path.join(OUT, "sbm_100_50.npz"),
                        **binary_case(nrow, ncol, rows, cols, colblock=8, bs=8, Rs=[2, 3, 4, 8, 32], lam=5.0))
    nrow, ncol, rows, cols, vals = ref_read_sdm(os.path.join(DATA, "sdm-100-50.data"))
    np.savez_compressed(os.path.join(OUT, "sdm_100_50.npz"), **double_case(nrow, ncol, rows, cols, vals, bs=8, Rs=[2, 5, 32]))
    r, c = random_coo(rng, 300, 70, 2000)
    np.savez_compressed(os.path.join(OUT, "rand_bin_300_70.npz"),
                        **binary_case(300, 70, r, c, colblock=16, bs=32, Rs=[2, 4, 7, 8, 32, 40], lam=3.0))
    r, c = random_coo(rng, 257, 129, 1500, dup_frac=0.0)   # no duplicate coordinates: payload order is unique
    v = rng.random(1500)
    np.savez_compressed(os.path.join(OUT, "rand_dbl_257_129.npz"), **double_case(257, 129, r, c, v, bs=64, Rs=[2, 8, 32]))
    np.savez_compressed(os.path.join(OUT, "hilbert.npz"), **hilbert_case(rng))
    np.savez_compressed(os.path.join(OUT, "linalg.npz"), **linalg_case(rng))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

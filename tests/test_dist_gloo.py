"""Multi-rank host logic on CPU (gloo, world_size 2): the row partition, the shard
extraction and the exchange pattern of SURVEY 8e.  No GPU here, so each rank's local
product is computed by the oracle (the checker) -- what is under test is that
  * fsb_partition_rows gives every rank a contiguous, nnz-balanced row range,
  * A x needs no collective (the shards' Y slabs concatenate to the full product),
  * A' x and A'(A x) are the SUM-allreduce of the per-shard partials,
  * the CG Gram reduction is an allreduce of R x R partials."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import libfastsparse_b200 as fs
import oracle


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nrow, ncol, nnz, R = 4000, 300, 50000, 4
        rows, cols, vals = fs.synth_coo_host(99, 1, nnz, nrow, ncol, with_vals=True)
        rows[:8000] = 17                                           # a heavy row: balance by nnz, not by rows
        M = fs.new_csr(nnz, nrow, ncol, rows, cols, vals)
        b = fs.partition_rows(M.row_ptr, world)
        r0, r1 = int(b[rank]), int(b[rank + 1])
        lo, hi = int(M.row_ptr[r0]), int(M.row_ptr[r1])
        rp = (M.row_ptr[r0:r1 + 1] - lo).astype(np.int32); cc = M.cols[lo:hi]; vv = M.vals[lo:hi]
        rng = np.random.default_rng(5)
        X = rng.standard_normal((ncol, R)); Xt = rng.standard_normal((nrow, R))
        # A x: local slab only
        Yloc = oracle.csr_mul(r1 - r0, rp, cc, vv, X, R)
        slabs = [None] * world
        dist.all_gather_object(slabs, (r0, r1, Yloc))
        # A' x: partial + allreduce
        trp, tcc, tvv = oracle.csr_from_coo(ncol, cc, np.repeat(np.arange(r1 - r0, dtype=np.int32), np.diff(rp)), vv)
        Zpart = torch.from_numpy(oracle.csr_mul(ncol, trp, tcc, tvv, Xt[r0:r1], R).copy())
        dist.all_reduce(Zpart)
        # A'(A x) partial + allreduce, and a Gram allreduce on a row-sharded tall matrix
        Kpart = torch.from_numpy(oracle.csr_mul(ncol, trp, tcc, tvv, Yloc, R).copy())
        dist.all_reduce(Kpart)
        Gpart = torch.from_numpy(Yloc.T @ Yloc)
        dist.all_reduce(Gpart)
        if rank == 0:
            Yfull = oracle.csr_mul(nrow, M.row_ptr, M.cols, M.vals, X, R)
            Ycat = np.concatenate([s[2] for s in sorted(slabs, key=lambda s: s[0])], 0)
            frp, fcc, fvv = oracle.csr_from_coo(ncol, cols, rows, vals)
            Zfull = oracle.csr_mul(ncol, frp, fcc, fvv, Xt, R)
            Kfull = oracle.csr_mul(ncol, frp, fcc, fvv, Yfull, R)
            per = np.diff(M.row_ptr[b])
            q.put(dict(cover=[(s[0], s[1]) for s in sorted(slabs, key=lambda s: s[0])], nrow=nrow,
                       y=float(np.max(np.abs(Ycat - Yfull))), z=float(np.max(np.abs(Zpart.numpy() - Zfull))),
                       k=float(np.max(np.abs(Kpart.numpy() - Kfull)) / np.max(np.abs(Kfull))),
                       g=float(np.max(np.abs(Gpart.numpy() - Yfull.T @ Yfull)) / np.max(np.abs(Yfull.T @ Yfull))),
                       imbalance=float(per.max() / (nnz / world))))
    finally:
        dist.destroy_process_group()


def test_row_sharding_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["cover"][0][0] == 0 and res["cover"][-1][1] == res["nrow"]
    assert all(a[1] == b[0] for a, b in zip(res["cover"], res["cover"][1:]))       # contiguous, disjoint
    assert res["y"] == 0.0                                                         # row slabs: bit-identical, no collective
    assert res["z"] < 1e-11 and res["k"] < 1e-13 and res["g"] < 1e-13
    assert res["imbalance"] < 1.2                                                  # nnz-balanced despite the heavy row


# ---------------------------------------------------------------------------------------------
# The exchange pattern of the multi-GPU block CG (fsb_cg.cu, cg_run_sharded) on CPU: vectors
# sharded over the unknowns in the library's chunk / slice geometry (fsb_cg_shard_layout),
# reduce-scatter of the A'(A P) partial chunk by chunk, all-gather of P, allreduce of the R x R
# Gram matrices.  Local products come from the oracle; what is under test is that the geometry
# (padding included) and the sequence of collectives reproduce the unsharded solve.
def _block_cg(op, gram, B, tol, max_iter=200):
    """bsbm_cg2's recurrence (cg.h:85-187) for R columns on whatever rows the caller holds."""
    norm = np.sqrt(np.maximum(np.diag(gram(B, B)), 1e-300))
    Rm = B / norm; P = Rm.copy(); X = np.zeros_like(B)
    G1 = gram(Rm, Rm)
    it = 0
    for it in range(max_iter):
        KP = op(P)
        alpha = np.linalg.solve(gram(P, KP), G1)
        X += P @ alpha; Rm -= KP @ alpha
        G2 = gram(Rm, Rm)
        if np.all(np.diag(G2) <= tol * tol):
            break
        P = Rm + P @ np.linalg.solve(G1, G2); G1 = G2
    return X * norm, it


def _cg_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nrow, F, nnz, R, lam, tol = 30000, 131_100, 400_000, 32, 15.0, 1e-8      # F*R*8 >= 32 MB: four chunks; F % (4*world) != 0: padding
        rows, cols, _ = fs.synth_coo_host(4711, 0, nnz, nrow, F)
        M = fs.new_bcsr(nnz, nrow, F, rows, cols)
        b = fs.partition_rows(M.row_ptr, world)
        r0, r1 = int(b[rank]), int(b[rank + 1])
        lo, hi = int(M.row_ptr[r0]), int(M.row_ptr[r1])
        rp = (M.row_ptr[r0:r1 + 1] - lo).astype(np.int32); cc = M.cols[lo:hi]
        trp, tcc, _ = oracle.csr_from_coo(F, cc, np.repeat(np.arange(r1 - r0, dtype=np.int32), np.diff(rp)))
        C, s, Fc, Fp, nloc = fs.cg_shard_layout(F, R, world)
        assert C == 4 and Fp >= F and Fp - F < C * world * 2 and nloc * world == Fp and Fc == s * world

        def to_local(full):          # rows [c*Fc + rank*s, +s) of every chunk, zero-padded past F
            pad = np.zeros((Fp, R)); pad[:F] = full
            return np.concatenate([pad[c * Fc + rank * s: c * Fc + (rank + 1) * s] for c in range(C)], 0)

        def allgather(loc):          # chunk by chunk, like shard_allgather
            full = np.zeros((Fp, R))
            for c in range(C):
                parts = [torch.zeros(s, R, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(loc[c * s:(c + 1) * s])))
                full[c * Fc:(c + 1) * Fc] = torch.cat(parts, 0).numpy()
            return full

        def op(Ploc):
            Pfull = allgather(Ploc)
            tmp = oracle.csr_mul(r1 - r0, rp, cc, None, np.ascontiguousarray(Pfull[:F]), R)
            part = np.zeros((Fp, R)); part[:F] = oracle.csr_mul(F, trp, tcc, None, tmp, R)
            KP = np.zeros((nloc, R))
            for c in range(C):       # reduce-scatter of chunk c (gloo has no reduce_scatter: allreduce + own slice)
                t = torch.from_numpy(part[c * Fc:(c + 1) * Fc].copy()); dist.all_reduce(t)
                KP[c * s:(c + 1) * s] = t.numpy()[rank * s:(rank + 1) * s]
            return KP + lam * Ploc

        def gram(Xa, Xb):
            t = torch.from_numpy(Xa.T @ Xb); dist.all_reduce(t); return t.numpy()

        rng = np.random.default_rng(11)
        B = rng.standard_normal((F, R))
        Xloc, it = _block_cg(op, gram, to_local(B), tol)
        Xfull = allgather(Xloc)[:F]
        if rank == 0:
            frp, fcc, _ = oracle.csr_from_coo(F, cols, rows)
            full_op = lambda P: oracle.csr_mul(F, frp, fcc, None, oracle.csr_mul(nrow, M.row_ptr, M.cols, None, np.ascontiguousarray(P), R), R) + lam * P
            Xref, itref = _block_cg(full_op, lambda a, b: a.T @ b, B, tol)
            res = np.linalg.norm(full_op(Xfull) - B, axis=0) / np.linalg.norm(B, axis=0)
            q.put(dict(err=float(np.max(np.abs(Xfull - Xref)) / np.max(np.abs(Xref))), it=it, itref=itref, res=float(res.max())))
    finally:
        dist.destroy_process_group()


def test_sharded_block_cg_exchange_pattern_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["err"] < 1e-9 and abs(res["it"] - res["itref"]) <= 1 and res["res"] < 1e-7

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

"""Row-sharded configs of BASELINE.json on N GPUs (run under torchrun, one rank per GPU):

  C3  double CSR 10M x 1M, 200M nnz: SpMV (no collective) and At_mul_B (NCCL allreduce of the
      [F] partial), STRONG scaling: the one matrix is cut into N nnz-balanced row shards.
  C5  block CG (lambda I + A'A) X = B, R = 32, on the C2 matrix, row-sharded, allreduce of the
      [F][32] partial of A'(A P) every iteration (vectors replicated).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_dist.py [--small]

Timing: CUDA events on the launching stream after a barrier, max over ranks."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import libfastsparse_b200 as fs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="", help="c3 | c5")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fs.comm_init_from_torch()
    N, F, NNZ = (1_000_000, 100_000, 20_000_000) if args.small else (10_000_000, 1_000_000, 200_000_000)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def shard_of(seed, with_vals):
        full = fs.DeviceMatrix.synth(seed, 0, NNZ, N, F, with_vals=with_vals)
        rp, _, _ = full.download_csr()
        b = fs.partition_rows(rp, world)
        r0, r1 = int(b[rank]), int(b[rank + 1])
        sh = full.row_slice(r0, r1)
        sh.set_row_sharded(True)
        full.free()
        return sh, r0, r1

    out = []
    if args.only in ("", "c3"):
        bench_c3(args, world, NNZ, N, F, shard_of, timed, out)
    if args.only in ("", "c5"):
        bench_c5(args, world, rank, NNZ, N, F, shard_of, timed, out)
    if rank == 0:
        for o in out:
            print(json.dumps(o), flush=True)
    fs.comm_finalize()
    dist.destroy_process_group()


def bench_c3(args, world, NNZ, N, F, shard_of, timed, out):
    A, r0, r1 = shard_of(0x5EED0003, True)
    x = (torch.sin(7.0 * torch.arange(F, device="cuda", dtype=torch.float64) + 0.3) / 10).contiguous()
    y = torch.empty(r1 - r0, dtype=torch.float64, device="cuda")
    z = torch.empty(F, dtype=torch.float64, device="cuda")
    ms = timed(lambda: A.spmm(x, 1, out=y), args.reps)
    out.append(dict(config="C3 double CSR SpMV, row-sharded, no collective", n_gpus=world, ms=ms, nnz_per_s=NNZ / ms * 1e3, scaling="strong"))
    A.spmm_t(y, 1, out=z)
    ms = timed(lambda: A.spmm_t(y, 1, out=z), args.reps)
    out.append(dict(config="C3 double CSR At_mul_B, row-sharded, NCCL allreduce of [F]", n_gpus=world, ms=ms, nnz_per_s=NNZ / ms * 1e3, scaling="strong"))
    A.free(); del A, x, y, z


def bench_c5(args, world, rank, NNZ, N, F, shard_of, timed, out):
    R = 32
    A, r0, r1 = shard_of(0x5EED0002, False)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    Nn = torch.randn(N * R, dtype=torch.float64, device="cuda", generator=g)
    E = torch.randn(F * R, dtype=torch.float64, device="cuda", generator=g)
    Bm = A.spmm_t(Nn[r0 * R: r1 * R].contiguous(), R) + (15.0 ** 0.5) * E     # allreduced inside: B = A'N + sqrt(lambda) E
    del Nn, E
    for mode, name in ((0, "vectors sharded over F: reduce-scatter (overlapped) + all-gather of P + R x R Gram allreduce"),
                       (1, "vectors replicated: allreduce of the [F][32] partial")):
        fs.check(fs.lib().fsb_tune_cg_dist(mode))
        Xs, it = A.cg(Bm, R, lam=15.0, tol=1e-6)          # warm-up (builds the cached transpose)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        Xs, it = A.cg(Bm, R, lam=15.0, tol=1e-6)
        dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        res = (A.ata(Xs, R, lam=15.0) - Bm).reshape(F, R).norm(dim=0) / Bm.reshape(F, R).norm(dim=0)
        per_it = dt / (it + 1)
        out.append(dict(config="C5 block CG R=32 lambda=15 tol=1e-6, row-sharded A; " + name, n_gpus=world,
                        iterations=it, seconds=dt, ms_per_iteration=per_it * 1e3, nnz_rhs_per_s=2 * NNZ * R / per_it, max_rel_residual=float(res.max()),
                        scaling="strong"))
    fs.check(fs.lib().fsb_tune_cg_dist(0))
    ms = timed(lambda: A.ata(Xs, R, lam=15.0), 5)
    out.append(dict(config="C5 operator A'(A X)+lambda X, R=32, row-sharded + allreduce", n_gpus=world, ms=ms, nnz_rhs_per_s=2 * NNZ * R / ms * 1e3, scaling="strong"))


if __name__ == "__main__":
    main()

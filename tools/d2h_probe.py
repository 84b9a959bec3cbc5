"""Concurrent host <-> device copy ceiling of the box (run under torchrun, one rank per GPU): every rank copies a
pinned buffer to / from its GPU at the same time; the aggregate is what the end-to-end (host-buffer) product can
reach at N GPUs, whatever the kernels do.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 27                                        # 1 GiB of fp64
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    out = {"n_gpus": world, "bytes_per_rank": n * 8}
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_ms_max_over_ranks"] = float(t.item())
        out[name + "_gbs_per_rank"] = n * 8 / (float(t.item()) * 1e-3) / 1e9
        out[name + "_gbs_aggregate"] = world * n * 8 / (float(t.item()) * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

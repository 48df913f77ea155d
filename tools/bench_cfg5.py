#!/usr/bin/env python3
"""BASELINE.json configs[4]: MNT4-753 Fr forward + inverse NTT (omega^-1, n^-1) at 2^26 across the ranks
(fused four-step).  torchrun ... tools/bench_cfg5.py [LOGN] [STEPS].  Checks inverse(forward(x)) == x bit for
bit on every rank, then times forward+inverse pairs with CUDA events (max over ranks); rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpusnarks_b200 as g  # noqa: E402
from gpusnarks_b200 import field as F  # noqa: E402
from gpusnarks_b200 import fourstep  # noqa: E402


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = g.Context(local)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    n = 1 << logn
    plan = fourstep.FusedFourStepNTT768(ctx, dev, logn, F.root_of_unity768(n))
    gen = torch.Generator(device=dev)
    gen.manual_seed(100 + rank)
    x0 = torch.randint(-(1 << 31), (1 << 31) - 1, plan.column_block_shape(), dtype=torch.int32, device=dev, generator=gen)
    x0[..., 23] &= 0xFFFF
    plan.x.copy_(x0)
    y = plan.forward()
    checksum = int(y.to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF
    back = plan.inverse()
    ok = bool((back == x0).all())
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for _ in range(2):
        plan.forward(); plan.inverse()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        plan.forward()
        plan.inverse()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        bf = 2 * (n // 2) * logn
        print(json.dumps({"workload": f"MNT4-753 Fr forward + inverse NTT n=2^{logn} on {world}xB200 (fused four-step)", "n_gpus": world,
                          "roundtrip_bit_exact_all_ranks": bool(flag.item()), "ms_per_forward_plus_inverse": float(ms.item()),
                          "butterflies_per_s": bf / (float(ms.item()) * 1e-3), "forward_checksum_rank0": checksum, "steps": steps}))
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

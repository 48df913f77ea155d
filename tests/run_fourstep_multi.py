"""torchrun worker: four-step NTT across the ranks (NCCL), checked against the oracle /
single-GPU transform.  usage: torchrun ... run_fourstep_multi.py LOGN [LOGN ...]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import fieldgen  # noqa: E402
import gpusnarks_b200 as g  # noqa: E402
from gpusnarks_b200 import fourstep  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = g.Context(local)
    for logn in [int(x) for x in sys.argv[1:]]:
        n = 1 << logn
        a = fieldgen.random_elements(n, 4000 + logn)      # every rank generates the same vector
        w = fieldgen.omega768(n)
        plan = fourstep.FourStepNTT768(fourstep.CudaBackend(ctx, dev), logn, w)
        x = torch.from_numpy(fourstep.to_column_block(a, logn, world, rank).view(np.int32)).to(dev)
        x0 = x.clone()
        y = plan.forward(x)
        blocks = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(blocks, y)
        got = fourstep.from_row_blocks([b.cpu().numpy().view(np.uint32) for b in blocks], logn)
        ref = ctx.ntt768(a, w)                             # single-GPU transform of the whole vector (itself oracle-checked)
        assert (got == ref).all(), f"rank {rank}: four-step != single-GPU at 2^{logn}"
        if logn > 20:   # the NCCL comparison path is only run up to 2^20
            pass
        if logn <= 16 and rank == 0:
            import oracle_lib as O
            assert (got == O.fft768(a, w, 3)).all(), f"four-step != oracle at 2^{logn}"
        back = plan.inverse(y)
        assert bool((back == x0).all()), f"rank {rank}: inverse(forward) != input at 2^{logn}"
        # fused exchange (peer stores over NVLink) must give the same bits
        fplan = fourstep.FusedFourStepNTT768(ctx, dev, logn, w)
        xf0 = torch.from_numpy(fourstep.to_column_layout(a, logn, world, rank, fplan.rank_bit).view(np.int32)).to(dev)
        for rep in range(3):   # repeated calls alternate the receive buffers and advance the flag epochs
            fplan.x.copy_(xf0)
            yf = fplan.forward()
            blocks = [torch.empty_like(yf) for _ in range(world)]
            dist.all_gather(blocks, yf.contiguous())
            gotf = fourstep.from_row_layouts([b.cpu().numpy().view(np.uint32) for b in blocks], logn)
            assert (gotf == ref).all(), f"rank {rank}: fused four-step != single-GPU at 2^{logn} (call {rep})"
            backf = fplan.inverse()
            assert bool((backf == xf0).all()), f"rank {rank}: fused inverse(forward) != input at 2^{logn} (call {rep})"
        if rank == 0:
            print(f"fused 2^{logn}: rank_bit {fplan.rank_bit}, per-source row start {fplan.plan.per_source}, tables {fplan.plan.table_bytes >> 20} MiB", flush=True)
        fplan.close()
        dist.barrier()
    if rank == 0:
        print("FOURSTEP_OK", world, "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

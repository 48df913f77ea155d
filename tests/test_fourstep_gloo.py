"""Multi-rank host logic of the four-step driver (gpusnarks_b200/fourstep.py) on CPU:
world_size-2 and -4 gloo groups, the all-to-all and the two layouts are the product code, the
local transforms are done by a CPU stand-in backend built on the oracle (tests only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


class OracleBackend:
    """CPU stand-in for CudaBackend: same two operations, computed by the oracle."""

    def __init__(self):
        import oracle_lib as O
        import pyref
        self.O, self.pyref = O, pyref

    def ntt(self, t, n, batch, log_r, omega, inverse_root=False, no_scale=False, pre_table=None):
        O, pyref = self.O, self.pyref
        R = 1 << log_r
        v = t.numpy().view(np.uint32).reshape(batch, n, R, 24)
        if pre_table is not None:
            tab = pre_table.numpy().view(np.uint32).reshape(batch, n, R, 24)
            v[...] = O.fp768_binop("mul", v.reshape(-1, 24), tab.reshape(-1, 24)).reshape(v.shape)
        w = np.ascontiguousarray(omega, dtype=np.uint32)
        if inverse_root:
            w = O.fp768_pow(w, n - 1)
        assert no_scale or not inverse_root
        for b in range(batch):
            for r in range(R):
                v[b, :, r] = O.fft768(np.ascontiguousarray(v[b, :, r]), w, -1)

    def table(self, rows, cols, row0, col0, n_total, omega, inverse_root=False, scale=False):
        pyref = self.pyref
        p = pyref.FR
        w = pyref.from_limbs(omega) * pow(pyref.RMONT, -1, p) % p
        if inverse_root:
            w = pow(w, -1, p)
        s = pow(n_total, -1, p) if scale else 1
        vals = [pow(w, (row0 + r) * (col0 + c), p) * s * pyref.RMONT % p for r in range(rows) for c in range(cols)]
        return torch.from_numpy(pyref.ints_to_array(vals).view(np.int32).reshape(rows, cols, 24))


def _worker(rank, world, logn, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fieldgen
        import oracle_lib as O
        from gpusnarks_b200 import fourstep
        n = 1 << logn
        a = fieldgen.random_elements(n, 900 + logn)
        w = fieldgen.omega768(n)
        plan = fourstep.FourStepNTT768(OracleBackend(), logn, w)
        x = torch.from_numpy(fourstep.to_column_block(a, logn, world, rank).view(np.int32))
        x0 = x.clone()
        y = plan.forward(x)
        blocks = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(blocks, y)
        got = fourstep.from_row_blocks([b.numpy().view(np.uint32) for b in blocks], logn)
        ok_fwd = bool((got == O.fft768(a, w, -1)).all())
        back = plan.inverse(y)
        ok_inv = bool((back == x0).all())
        q.put((rank, ok_fwd, ok_inv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,logn", [(2, 4), (2, 7), (4, 6)])
def test_fourstep_exchange_logic(world, logn):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + logn
    procs = [ctx.Process(target=_worker, args=(r, world, logn, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True, True) for r in range(world)], res

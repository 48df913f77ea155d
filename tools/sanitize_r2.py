#!/usr/bin/env python3
"""workload for compute-sanitizer (memcheck / racecheck) over the round-2 kernels: the warp-owned large-tile kernel
(phase A only, phase A + B, pre- and post-twiddles, two-level tables), the four-step plan with one rank (scatter +
flags), the bucket-method multiexp and Fq2.  Every result is still checked (against the CTA-wide kernel / round trips).
usage: compute-sanitizer --tool racecheck python tools/sanitize_r2.py"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fieldgen  # noqa: E402
import g1ref  # noqa: E402
import pyref  # noqa: E402
import gpusnarks_b200 as g  # noqa: E402


def run(ctx, n, batch, variant, limit=None):
    ctx.set_option("kernel_variant", variant)
    ctx.set_option("flat_table_limit", limit if limit is not None else 4 << 30)
    a = fieldgen.random_elements(n * batch, 9000 + n)
    w = fieldgen.omega768(n)
    d = ctx.device_alloc(a.nbytes)
    ctx.h2d(d, a)
    ctx.ntt768_device(d, n, w, batch=batch)
    out = np.empty_like(a)
    ctx.d2h(out, d)
    ctx.ntt768_device(d, n, w, inverse=True, batch=batch)
    back = np.empty_like(a)
    ctx.d2h(back, d)
    ctx.device_free(d)
    assert (back == a).all()
    return out


def main():
    ctx = g.Context(0)
    for n, batch in ((1 << 10, 512), (1 << 8, 2048), (1 << 19, 1)):
        ref = run(ctx, n, batch, 4)
        for variant, limit in ((1, None), (5, None), (5, 1 << 12)):
            assert (run(ctx, n, batch, variant, limit) == ref).all(), (n, variant, limit)
    ctx.set_option("kernel_variant", 1)
    ctx.trim()
    n = 1 << 19
    a = fieldgen.random_elements(n, 9100)
    w = fieldgen.omega768(n)
    shift = pyref.ints_to_array([pyref.mont(17)])[0]
    assert (ctx.coset_ntt768(ctx.coset_ntt768(a, w, shift), w, shift, inverse=True) == a).all()
    # four-step plan, one rank: scatter stores + flags + [k2][r] rows
    import torch
    from gpusnarks_b200 import fourstep
    dev = torch.device("cuda", 0)
    plan = fourstep.FusedFourStepNTT768(ctx, dev, 19, w)
    x0 = torch.from_numpy(fourstep.to_column_layout(a, 19, 1, 0).view(np.int32)).to(dev)
    plan.x.copy_(x0)
    y = plan.forward()
    torch.cuda.synchronize()
    got = fourstep.from_row_layouts([y.cpu().numpy().view(np.uint32)], 19)
    assert (got == ctx.ntt768(a, w)).all()
    back = plan.inverse()
    torch.cuda.synchronize()
    assert bool((back == x0).all())
    plan.close()
    # bucket-method multiexp (incl. a heavy bucket) and Fq2
    rng = random.Random(3)
    pts = [g1ref.random_point(rng) for _ in range(8)]
    P = np.zeros((200, 3, 24), dtype=np.uint32)
    for i in range(200):
        for c, v in enumerate(g1ref.to_projective_mont(pts[i % 8])):
            P[i, c] = pyref.to_limbs(v)
    ks = [rng.randrange(pyref.FR) for _ in range(200)]
    ks[:150] = [5] * 150
    out = ctx.g1_multiexp(P, pyref.ints_to_array(ks), method="bucket", window_bits=4)
    aff = g1ref.from_projective_mont(*[pyref.from_limbs(out[c]) for c in range(3)])
    assert aff == g1ref.multiexp(pts, [sum(ks[j::8]) for j in range(8)])
    # split buckets: 9 000 points in one bucket (three parts + the combine kernel), and the split window reduction (c = 13)
    P2 = P[np.arange(9000) % 8]
    out = ctx.g1_multiexp(P2, pyref.ints_to_array([7] * 9000), method="bucket", window_bits=13)
    aff = g1ref.from_projective_mont(*[pyref.from_limbs(out[c]) for c in range(3)])
    assert aff == g1ref.multiexp(pts, [7 * 1125] * 8)
    # pageable host path: column blocks through the pinned bounce buffers (numpy memory is pageable), 2^16 = two passes
    n = 1 << 16
    a = fieldgen.random_elements(n, 9200)
    w = fieldgen.omega768(n)
    assert (ctx.ntt768(ctx.ntt768(a, w), w, inverse=True) == a).all()
    q = g1ref.Q
    fa = np.stack([fieldgen.random_elements(64, 1, q), fieldgen.random_elements(64, 2, q)], axis=1)
    ctx.fp2_binop("mul", fa, fa)
    ctx.close()
    print("sanitize_r2 workload ok")


if __name__ == "__main__":
    main()

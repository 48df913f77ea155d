#!/usr/bin/env python3
"""G1 multi-exponentiation timing on one GPU: the reference's README shape (2^16 points, full 753-bit scalars), both
algorithms.  usage: python tools/gpu_probe_g1.py [LOGN ...]"""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import g1ref  # noqa: E402
import pyref  # noqa: E402
import gpusnarks_b200 as g  # noqa: E402


def main():
    ctx = g.Context(0)
    rng = random.Random(7)
    base = [g1ref.random_point(rng) for _ in range(64)]
    packed = np.zeros((64, 3, 24), dtype=np.uint32)
    for i, P in enumerate(base):
        for c, v in enumerate(g1ref.to_projective_mont(P)):
            packed[i, c] = pyref.to_limbs(v)
    for logn in [int(x) for x in sys.argv[1:]] or [12, 16]:
        n = 1 << logn
        pts = packed[np.arange(n) % 64]
        nrng = np.random.Generator(np.random.PCG64(logn))
        ks = nrng.integers(0, 1 << 32, size=(n, 24), dtype=np.uint64).astype(np.uint32)
        ks[:, 23] &= 0xFFFF   # < 2^752 < r
        res = {}
        for method in (["naive", "bucket"] if logn <= 16 else ["bucket"]):
            ctx.g1_multiexp(pts[:64], ks[:64], method=method)   # warm-up
            t0 = time.perf_counter()
            out = ctx.g1_multiexp(pts, ks, method=method)
            res[method] = {"ms_host_call": (time.perf_counter() - t0) * 1e3}
            res[method]["out0"] = int(out[0, 0])
        print(json.dumps({"log_n": logn, "points": n, **res}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

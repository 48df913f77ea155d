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
        dp, ds, do = ctx.device_alloc(n * 288), ctx.device_alloc(n * 96), ctx.device_alloc(288)
        ctx.h2d(dp, pts)
        ctx.h2d(ds, ks)
        affine = {}
        for method in ("naive", "bucket"):
            ms = []
            for rep in range(4 if method == "bucket" or logn <= 16 else 1):   # the first call grows the scratch arena
                ctx.synchronize()
                t0 = time.perf_counter()
                ctx.g1_multiexp_device(do, dp, ds, n, method=method)
                ctx.synchronize()
                ms.append((time.perf_counter() - t0) * 1e3)
            res[method + "_device_resident_ms"] = min(ms[1:]) if len(ms) > 1 else ms[0]
            got = np.empty((3, 24), dtype=np.uint32)
            ctx.d2h(got, do)
            affine[method] = g1ref.from_projective_mont(*[pyref.from_limbs(got[c]) for c in range(3)])
        res["bucket_equals_reference_algorithm"] = affine["naive"] == affine["bucket"]
        for p in (dp, ds, do):
            ctx.device_free(p)
        ms = []
        for rep in range(3):
            t0 = time.perf_counter()
            out = ctx.g1_multiexp(pts, ks)      # gsn_g1_multiexp_host: pageable numpy arrays in, 288-byte point out
            ms.append((time.perf_counter() - t0) * 1e3)
        res["host_call_ms"] = min(ms)
        res["host_call_all_ms"] = ms
        res["points_per_s_host_call"] = n / (min(ms) * 1e-3)
        if g.device_count() > 1 and logn >= 16:
            from gpusnarks_b200.ntt import g1_multiexp_multi
            ms, ok = [], True
            for rep in range(3):
                t0 = time.perf_counter()
                got = g1_multiexp_multi(pts, ks)     # one slice per visible device, partial sums added on the host
                ms.append((time.perf_counter() - t0) * 1e3)
            res["devices"] = g.device_count()
            res["multi_host_call_ms"] = min(ms)
            res["multi_equals_single"] = g1ref.from_projective_mont(*[pyref.from_limbs(got[c]) for c in range(3)]) == affine["bucket"]
        print(json.dumps({"log_n": logn, "points": n, **res}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

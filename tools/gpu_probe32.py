#!/usr/bin/env python3
"""GPU-box probe for the 32-bit path: device-resident timings at 2^22 (cfg2) and others."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fieldgen  # noqa: E402
import gpusnarks_b200 as g  # noqa: E402
import pyref  # noqa: E402


def main():
    ctx = g.Context(0)
    out = []
    for logn, batch in [(16, 1), (20, 1), (22, 1), (22, 16), (24, 4), (26, 1)]:
        n = 1 << logn
        a = fieldgen.random_u32(n * batch, 1, pyref.P32)
        w = fieldgen.omega32(n)
        d = ctx.device_alloc(a.nbytes)
        ctx.h2d(d, a)
        ms = ctx.time_ntt32(d, n, w, pyref.P32, batch=batch, reps=12)
        ctx.device_free(d)
        med = float(np.median(ms[2:])) / batch
        bf = (n // 2) * logn
        passes = -(-logn // 11)
        out.append({"logn": logn, "batch": batch, "ms_per_transform": med, "bf_per_s": bf / (med * 1e-3),
                    "algorithmic_gbs": (2 * passes) * n * 4 / (med * 1e-3) / 1e9})
        ctx.trim()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

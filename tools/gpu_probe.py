#!/usr/bin/env python3
"""GPU-box probe: INT32 issue rates + device-resident NTT timings.  Writes JSON to stdout."""
import json
import sys
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fieldgen  # noqa: E402
import gpusnarks_b200 as g  # noqa: E402


def main():
    logns = [int(x) for x in sys.argv[1:]] or [16, 20, 22]
    ctx = g.Context(0)
    out = {"int32": ctx.int32_issue_rates(), "ntt768": []}
    for logn in logns:
        n = 1 << logn
        a = fieldgen.random_elements(n, 1)
        w = fieldgen.omega768(n)
        d = ctx.device_alloc(a.nbytes)
        ctx.h2d(d, a)
        t0 = time.time()
        ctx.prepare768(n, w)
        prep = time.time() - t0
        ms = ctx.time_ntt768(d, n, w, reps=12)
        ctx.device_free(d)
        best, med = min(ms[2:]), float(np.median(ms[2:]))
        bf = (n // 2) * logn
        out["ntt768"].append({"logn": logn, "prepare_s": prep, "ms_all": ms, "ms_best": best, "ms_median": med,
                              "bf_per_s": bf / (med * 1e-3), "wide_mac_per_s": bf * 1176 / (med * 1e-3)})
        ctx.trim()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

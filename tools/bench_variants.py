#!/usr/bin/env python3
"""A/B timing of the 768-bit pass-kernel variants on one GPU (device resident, CUDA events inside the library):
variant 4 = CTA-wide kernel (round-1 structure: one barrier per stage, strict ranges), 1 = warp-owned tiles with wide
lazy ranges, 5 = the same with stages 3-4 enumerated CTA-wide (unit-twiddle skips).  Also flat vs two-level boundary tables.
usage: python tools/bench_variants.py [LOGN ...]   -> one JSON line per (size, variant) on stdout"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpusnarks_b200 as g  # noqa: E402
from gpusnarks_b200 import field as F  # noqa: E402


def main():
    logns = [int(x) for x in sys.argv[1:]] or [20, 24]
    for logn in logns:
        n = 1 << logn
        w = F.root_of_unity768(n)
        rng = np.random.Generator(np.random.PCG64(1))
        a = rng.integers(0, 1 << 32, size=(n, 24), dtype=np.uint64).astype(np.uint32)
        a[:, 23] &= 0xFFFF
        ref = None
        for variant, limit in [(4, None), (1, None), (5, None), (4, 1 << 16), (5, 1 << 16)]:
            ctx = g.Context(0)
            ctx.set_option("kernel_variant", variant)
            if limit:
                ctx.set_option("flat_table_limit", limit)
            d = ctx.device_alloc(a.nbytes)
            try:
                ctx.h2d(d, a)
                ctx.ntt768_device(d, n, w)
                ctx.synchronize()
                out = np.empty_like(a)
                ctx.d2h(out, d)
                if ref is None:
                    ref = out
                same = bool((out == ref).all())
                ms = ctx.time_ntt768(d, n, w, reps=24)
                info = ctx.plan_info768(n, w)
                print(json.dumps({"log_n": logn, "variant": variant, "flat_table_limit": limit, "ms_median": float(np.median(ms[4:])),
                                  "ms_min": float(np.min(ms[4:])), "same_bits_as_first_variant": same, "table_bytes": info["table_bytes"],
                                  "two_level_boundaries": info["two_level_boundaries"],
                                  "butterflies_per_s": (n // 2 * logn) / (float(np.median(ms[4:])) * 1e-3)}), flush=True)
            finally:
                ctx.device_free(d)
                ctx.close()


if __name__ == "__main__":
    main()

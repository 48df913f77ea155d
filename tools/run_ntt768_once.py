#!/usr/bin/env python3
"""one warm-up + REPS device-resident 768-bit transforms of 2^LOGN elements, for ncu captures.
usage: python tools/run_ntt768_once.py LOGN VARIANT [REPS] [FLAT_TABLE_LIMIT]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusnarks_b200 as g  # noqa: E402
from gpusnarks_b200 import field as F  # noqa: E402

logn, variant = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 1 << logn
ctx = g.Context(0)
ctx.set_option("kernel_variant", variant)
if len(sys.argv) > 4:
    ctx.set_option("flat_table_limit", int(sys.argv[4]))
rng = np.random.Generator(np.random.PCG64(1))
a = rng.integers(0, 1 << 32, size=(n, 24), dtype=np.uint64).astype(np.uint32)
a[:, 23] &= 0xFFFF
d = ctx.device_alloc(a.nbytes)
ctx.h2d(d, a)
w = F.root_of_unity768(n)
for _ in range(1 + reps):
    ctx.ntt768_device(d, n, w)
ctx.synchronize()
print("done", ctx.launch_count())

#!/usr/bin/env python3
"""a few 32-bit transforms of 2^LOGN elements for ncu, over BATCH distinct buffers per launch pair (16 x 16 MiB = 256 MiB
exceeds the L2, so the profile shows the HBM-resident configuration bench.py times):
python tools/run_ntt32_once.py [logn] [reps] [batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fieldgen, pyref
import gpusnarks_b200 as g
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ctx = g.Context(0)
n = 1 << logn
a = fieldgen.random_u32(n * batch, 1, pyref.P32)
d = ctx.device_alloc(a.nbytes)
ctx.h2d(d, a)
for _ in range(reps):
    ctx.ntt32_device(d, n, fieldgen.omega32(n), pyref.P32, batch=batch)
ctx.synchronize()

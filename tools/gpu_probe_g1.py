#!/usr/bin/env python3
"""timing of the G1 multiexp at the reference's README sizes (2^16 points, full-width scalars)"""
import os, sys, time, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import g1ref, pyref, fieldgen
import gpusnarks_b200 as g
ctx = g.Context(0)
ctx.set_field768(g.FIELD_FQ)
rng = random.Random(1)
base = [g1ref.random_point(rng) for _ in range(16)]
bp = np.zeros((16, 3, 24), dtype=np.uint32)
for i, P in enumerate(base):
    for c, v in enumerate(g1ref.to_projective_mont(P)):
        bp[i, c] = pyref.to_limbs(v)
for logn in (12, 16):
    n = 1 << logn
    pts = np.ascontiguousarray(bp[np.arange(n) % 16])
    ks = fieldgen.random_elements(n, 5, pyref.FR)
    ctx.g1_multiexp(pts[:128], ks[:128])
    t0 = time.time(); out = ctx.g1_multiexp(pts, ks); dt = time.time() - t0
    print(f"G1 multiexp n=2^{logn}: {dt*1e3:.1f} ms wall (host call incl. copies), {n/dt:.3e} points/s")

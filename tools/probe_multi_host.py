#!/usr/bin/env python3
"""best_fft's large-vector path on a multi-GPU box: gsn_multi_ntt768_host (all devices, one process) against
gsn_ntt768_host on one device, pageable and pinned host vectors.  usage: python tools/probe_multi_host.py [LOGN]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fieldgen  # noqa: E402
import gpusnarks_b200 as g  # noqa: E402
from gpusnarks_b200.ntt import MultiGpu  # noqa: E402


def timed(fn, reps=3):
    ms = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ms.append((time.perf_counter() - t0) * 1e3)
    return min(ms)


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = 1 << logn
    G = 1
    while G * 2 <= min(g.device_count(), 8):
        G *= 2
    w = fieldgen.omega768(n)
    a = fieldgen.random_elements(n, 1)
    pinned = torch.from_numpy(a.view(np.int32)).clone().pin_memory().numpy().view(np.uint32)
    ctx = g.Context(0)
    want = ctx.ntt768(a, w)
    res = {"log_n": logn, "devices": G}
    v = a.copy()
    res["one_gpu_pageable_ms"] = timed(lambda: ctx.best_fft768(v, w))
    res["one_gpu_pinned_ms"] = timed(lambda: ctx.best_fft768(pinned, w))
    ctx.close()
    m = MultiGpu(list(range(G)), n, w, directions=("forward",))
    v = a.copy()
    m.ntt_host(v)
    res["multi_equals_one_gpu"] = bool((v == want).all())
    res["multi_pageable_ms"] = timed(lambda: m.ntt_host(v))
    res["multi_pinned_ms"] = timed(lambda: m.ntt_host(pinned))
    m.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()

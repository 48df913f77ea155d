#!/usr/bin/env python3
"""bench.py -- the hot-path benchmark of gpusnarks_b200 (contract: see the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log-n L]

Metric (BASELINE.json): 768-bit NTT butterflies/s, butterflies = (n/2) * log2(n) per transform.
  N = 1   workload = BASELINE.json configs[2]: MNT4-753 Fr forward NTT, n = 2^20, one B200.
  N > 1   workload = configs[3]: MNT4-753 Fr forward NTT, n = 2^24, four-step sharded across
          the N GPUs (strong scaling: the same transform on more GPUs); the exchange is fused into
          the transform kernels (peer stores over NVLink, arrival flags), gsn_fourstep_*.
          Outside the timed region the sharded result is compared BIT FOR BIT with the single-GPU
          transform of the same input on rank 0 and spot-checked against the CPU oracle ("parity");
          at N = 8 a secondary runs configs[4] (2^26 forward + inverse) the same way.
A "step" is one whole transform of synthetic random field elements.
  value   device-resident: inputs already in HBM, K steps between barrier + synchronize,
          CUDA events on the launching stream, max over ranks.
  e2e     the same transform through the host-pointer C ABI call (gsn_ntt768_host): pinned
          host buffer -> H2D -> transform -> D2H inside the timed region, every step.
  roofline  the dominant kernel (ntt768_pass) against the INT32 multiply issue peak measured
          in this same process (gsn_int32_issue_rates, IMAD.WIDE.U32 accumulate form), algorithmic work
          = 1176 wide MACs per butterfly; HBM figures given alongside.
  cpu_baseline  the reference's own host FFT (oracle/_ref/libref_verbatim.so, compiled from
          /root/reference, test/fft_host.h via the host half of test/main.cpp:64-76) timed on
          this box's cores on a bounded sample.
--impl reference times only that CPU reference (rank 0) on the same config and metric.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MACS_PER_BUTTERFLY = 1176  # 2*24^2 + 24


def butterflies(logn):
    return (1 << logn) // 2 * logn


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, v in zip(names, parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU reference
def _ref_lib():
    """the reference compiled from /root/reference (kind 'reference'), else the oracle port"""
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    path = os.path.join(ROOT, "oracle", "_ref", "libref_verbatim.so")
    if os.path.exists(path):
        L = C.CDLL(path)
        L.ref_fft_scalar.argtypes = [u32p, C.c_size_t, u32p, C.c_int, C.c_int]
        L.ref_fft_scalar.restype = C.c_double
        return "reference", lambda a, n, w, lc, thr: L.ref_fft_scalar(a, n, w, lc, thr)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    lib = O.lib()

    def run(a, n, w, lc, thr):
        lib.oracle_set_threads(thr)
        return lib.oracle_time_fft768(a, n, w, lc)
    return "port", run


def synth_host(n, seed):
    """seeded canonical elements: 24 random limbs, top limb masked to 16 bits (< 2^752 < r)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 32, size=(n, 24), dtype=np.uint64).astype(np.uint32)
    a[:, 23] &= 0xFFFF
    return a


def cpu_reference_run(sample_logn, steps, warmup):
    from gpusnarks_b200 import field as F
    kind, run = _ref_lib()
    cores = os.cpu_count() or 1
    log_cpus = int(math.log2(cores))
    n = 1 << sample_logn
    a = synth_host(n, 1)
    w = F.root_of_unity768(n)
    times = []
    for i in range(warmup + steps):
        buf = a.copy()
        t = run(buf, n, w, log_cpus, cores)
        if i >= warmup:
            times.append(t)
    sec = float(np.mean(times))
    sample = (f"one forward host FFT (reference test/fft_host.h _basic_parallel_radix2_FFT_inner<fields::Scalar>, log_cpus={log_cpus}, "
              f"{cores} OpenMP threads, g++ -O2) of n=2^{sample_logn} random elements per step; mean of {steps} step(s)")
    return {"value": butterflies(sample_logn) / sec, "unit": "butterflies/s", "cores": cores, "kind": kind, "sample": sample,
            "seconds_per_step": sec}


def spot_check(a_host, omega, got_rows, ks, inverse_of=None):
    """checker leg: A[k] by Horner on the CPU oracle (all host threads) for the indices ks, against the GPU rows"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    exp = O.dft_points768_mt(a_host, omega, ks)
    return bool((exp == got_rows).all())


def ncu_traffic_bytes(name, per="launch", divide=1.0):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed summary of one
    `ncu --set full` capture (profiles/<name>.raw.csv, made by tools/summarize_ncu.py): the mean per launch, or
    (per="sum") the sum over the captured launches divided by `divide`"""
    import csv
    path = os.path.join(ROOT, "profiles", name + ".raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:]]
        return sum(tot) / divide if per == "sum" else sum(tot) / len(tot)
    except Exception:
        return None


def executed_utilisation(logn, ms_step, p_mac):
    """wide-multiply issue slots the single-GPU transform really executes (fixed-operand product: 876 IMAD.WIDE + 48 IMAD
    = 900 wide-equivalents; per pass of l stages a tile runs l - 1.9375 stages' worth of products -- stage 1 and the unit
    butterflies of stages 2-5 are skipped -- plus one product per element at every pass boundary) / time / peak"""
    n = 1 << logn
    passes = max(1, -(-logn // 10))
    base, extra = divmod(logn, passes)
    digits = [base + (1 if i < extra else 0) for i in range(passes)]
    products = sum((max(l - 1.9375, 0)) * n / 2 for l in digits) + (passes - 1) * n
    slots = products * 900.0
    return {"products_per_transform": products, "wide_equiv_slots_per_product": 900, "slots_per_s": slots / (ms_step * 1e-3),
            "frac_of_peak": slots / (ms_step * 1e-3) / p_mac}


def bench_g1(ctx, logn=16, reps=3):
    """MNT4-753 G1 multi-exponentiation, the reference README's shape (2^16 points, full-width
    scalars), device resident.  Curve points come from tests/g1ref.py (input generation only; parity is tests/test_gpu_g1.py)."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import g1ref
    import pyref
    n = 1 << logn
    rng = random.Random(7)
    packed = np.zeros((64, 3, 24), dtype=np.uint32)
    for i in range(64):
        for c, v in enumerate(g1ref.to_projective_mont(g1ref.random_point(rng))):
            packed[i, c] = pyref.to_limbs(v)
    pts = packed[np.arange(n) % 64]
    ks = np.random.Generator(np.random.PCG64(logn)).integers(0, 1 << 32, size=(n, 24), dtype=np.uint64).astype(np.uint32)
    ks[:, 23] &= 0xFFFF   # < 2^752
    dp, ds, do = ctx.device_alloc(n * 288), ctx.device_alloc(n * 96), ctx.device_alloc(288)
    try:
        ctx.h2d(dp, pts)
        ctx.h2d(ds, ks)
        out = {}
        for method, r in (("bucket", reps + 1), ("naive", 1)):
            ms = []
            for _ in range(r):
                ctx.synchronize()
                t0 = time.perf_counter()
                ctx.g1_multiexp_device(do, dp, ds, n, method=method)
                ctx.synchronize()
                ms.append((time.perf_counter() - t0) * 1e3)
            out[method] = min(ms[1:]) if len(ms) > 1 else ms[0]
    finally:
        for p in (dp, ds, do):
            ctx.device_free(p)
    return {"workload": f"MNT4-753 G1 multi-exponentiation, 2^{logn} points (64 distinct, affine), 752-bit scalars, device resident, 1xB200",
            "ms": out["bucket"], "points_per_s": n / (out["bucket"] * 1e-3), "algorithm": "bucket method (gsn_g1_multiexp_device_ex, automatic window)",
            "reference_algorithm_ms": out["naive"],
            "reference_algorithm": "one double-and-add per point + tree reduction (cuda/multi_exp.cu:86-137) on the same field arithmetic"}


def bench_ntt32(ctx, hbm_peak_gbs, logn=22, batch=16, reps=12):
    """BASELINE.json configs[1]: 32-bit prime-field NTT 2^22 on one B200 (HBM-bound).  `batch` distinct
    16 MiB buffers (256 MiB > L2) are transformed per launch pair so that inputs come from HBM; the
    roofline uses the algorithmic bytes of SURVEY.md section 8d: passes x 2 x n x 4 B (two passes)."""
    from gpusnarks_b200 import field as F
    n = 1 << logn
    rng = np.random.Generator(np.random.PCG64(3))
    a = rng.integers(0, F.P32, size=n * batch, dtype=np.uint64).astype(np.uint32)
    w = F.root_of_unity32(n)
    d = ctx.device_alloc(a.nbytes)
    try:
        ctx.h2d(d, a)
        ms = ctx.time_ntt32(d, n, w, F.P32, batch=batch, reps=reps)
    finally:
        ctx.device_free(d)
    t = float(np.median(ms[2:])) / batch * 1e-3
    alg = 2 * 2 * n * 4
    return {"workload": f"32-bit prime-field forward NTT n=2^{logn}, p=2013265921, {batch} distinct buffers per launch pair, 1xB200",
            "us_per_transform": t * 1e6, "value": butterflies(logn) / t, "unit": "butterflies/s",
            "roofline": {"kernel": "gsn::ntt32_fast_pass<4,4,3,...>", "bound": "hbm", "achieved": alg / t / 1e9, "peak": hbm_peak_gbs, "unit": "GB/s",
                         "frac": alg / t / 1e9 / hbm_peak_gbs, "algorithmic_bytes_per_transform": alg,
                         # the capture is the two passes of ONE 16-buffer batch (tools/run_ntt32_once.py 22 4 16): per transform = sum / 16
                         "traffic": ncu_traffic_bytes("ntt32_fast_pass_r02_batch16", per="sum", divide=16.0),
                         "traffic_note": "dram bytes of both passes of a 16-transform batch / 16 (profiles/ntt32_fast_pass_r02_batch16.md): the HBM-resident configuration timed here"}}


# ------------------------------------------------------------------------------ main
def emit(line):
    """the ONE JSON line goes to the real stdout; everything else this process (or NCCL) prints goes to stderr"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)  # libraries (e.g. NCCL's version banner) write to fd 1: keep the bench output a single line

DTYPE = "u32x24 (768-bit Montgomery, integer)"
FIRST_CALL_SCRIPT = r"""
import json, sys, time
import numpy as np
sys.path.insert(0, %r)
t0 = time.perf_counter()
import gpusnarks_b200 as g
from gpusnarks_b200 import field as F
n = 1 << %d
a = np.zeros((n, 24), dtype=np.uint32); a[:, 0] = 1234
w = F.root_of_unity768(n)
t1 = time.perf_counter()
ctx = g.Context(0)
t2 = time.perf_counter()
ctx.best_fft768(a, w)
t3 = time.perf_counter()
ctx.best_fft768(a, w)
t4 = time.perf_counter()
print(json.dumps({"import_ms": (t1 - t0) * 1e3, "ctx_create_ms": (t2 - t1) * 1e3, "first_call_ms": (t3 - t2) * 1e3, "second_call_ms": (t4 - t3) * 1e3}))
"""


_ALL_CPUS = os.sched_getaffinity(0)


def unbind_cpus():
    """the CPU legs (reference arm, oracle spot checks) run on every core of the box again"""
    try:
        os.sched_setaffinity(0, _ALL_CPUS)
    except Exception:
        pass


def bind_to_gpu_numa_node(gpu_index):
    """pin this process to the CPU cores NVML reports as local to its GPU BEFORE any pinned host memory is allocated
    (first touch then places the staging buffers on the GPU's NUMA node); with eight ranks copying at once the
    end-to-end path is bound by host memory placement, not by the transforms.  Returns the number of cores, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def peaks_file():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def roofline_block(value, N, logn, ms_step, rates, single, kernels_per_step, kernel_name, table_bytes=None):
    # roofline denominator: the HIGHEST wide-MAC issue rate the probes reach (distinct-operand, shared-operand
    # and carry-chain forms all issue at ~32 lanes/clk/SM; taking the maximum is the conservative choice)
    p_mac = max(rates["rates"][k] for k in ("imad_wide", "imad_wide_shared_operands", "imad_wide_x_chain"))
    achieved = value * MACS_PER_BUTTERFLY / N  # per GPU
    n = 1 << logn
    passes = max(1, -(-logn // 10)) if single else None
    peaks = peaks_file()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # algorithmic bytes per transform: each pass reads and writes every element once (data only)
    alg_bytes = 2 * passes * n * 96 if single else None
    return {
        "kernel": kernel_name,
        "bound": "int32_mul",
        "achieved": achieved / 1e12, "peak": p_mac / 1e12, "unit": "T wide-MAC/s (32x32+64 IMAD.WIDE.U32)",
        "frac": achieved / p_mac,
        "algorithmic_macs_per_butterfly": MACS_PER_BUTTERFLY,
        # frac can exceed 1: the 1176-MAC figure is the CIOS product of SURVEY.md 8d, while the kernel multiplies
        # by table twiddles with a fixed-operand product of 876 wide + 48 low multiplies (= 900 wide-equivalent
        # issue slots) and skips unit twiddles.  `executed` is the multiplier-pipe utilisation of what really runs
        # (analytic); the hardware counter (sm__pipe_fmaheavy_cycles_active) is in profiles/ntt768_pass_r02.md.
        "executed": executed_utilisation(logn, ms_step, p_mac) if single else None,
        "peak_source": "gsn_int32_issue_rates: max over the IMAD.WIDE.U32 probes (accumulate form with distinct / shared multiplicands, .X carry chains), 8 independent accumulators, measured in this process (SASS-verified loops)",
        "int32_issue_rates_per_s": rates["rates"],
        "launches_per_step": kernels_per_step,
        "avg_launch_ms": (ms_step / kernels_per_step) if kernels_per_step else None,
        "traffic": ncu_traffic_bytes("ntt768_pass_r02_cta_wide") or ncu_traffic_bytes("ntt768_pass_r01"),
        "table_bytes": table_bytes,
        "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": (alg_bytes / (ms_step * 1e-3) / 1e9) if alg_bytes else None,
                "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650",
                "note": "data only (2 x passes x n x 96 B); the pass-boundary twiddle table adds table_bytes of reads per transform"},
    }


def bench_pipeline(ctx, torch, dev, stream, logn, reps=5):
    """the prover's quotient pipeline, device resident -- iFFT(a), iFFT(b), coset FFT(a), coset
    FFT(b), pointwise product, coset iFFT: five transforms and one element-wise product per pipeline"""
    from gpusnarks_b200 import field as F
    n = 1 << logn
    w = F.root_of_unity768(n)
    shift = F.to_mont768(17)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    a = torch.randint(-(1 << 31), (1 << 31) - 1, (n, 24), dtype=torch.int32, device=dev, generator=gen)
    b = torch.randint(-(1 << 31), (1 << 31) - 1, (n, 24), dtype=torch.int32, device=dev, generator=gen)
    a[:, 23] &= 0xFFFF
    b[:, 23] &= 0xFFFF
    s = stream.cuda_stream

    def once():
        for t in (a, b):
            ctx.ntt768_device(t.data_ptr(), n, w, inverse=True, stream=s)
            ctx.coset_ntt768_device(t.data_ptr(), n, w, shift, stream=s)
        ctx.fp768_binop_device("mul", a.data_ptr(), a.data_ptr(), b.data_ptr(), n, stream=s)
        ctx.coset_ntt768_device(a.data_ptr(), n, w, shift, inverse=True, stream=s)
    once()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        once()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"workload": f"iFFT x2 -> coset FFT x2 -> pointwise product -> coset iFFT, n=2^{logn}, device resident (gsn_ntt768_device, gsn_coset_ntt768_device, gsn_fp768_binop_device)",
            "ms_per_pipeline": ms, "transforms_per_pipeline": 5, "butterflies_per_s": 5 * butterflies(logn) / (ms * 1e-3)}


def run_single(args, torch, g, F, ctx, dev, stream, logn, config, rates):
    n = 1 << logn
    omega = F.root_of_unity768(n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    data = torch.randint(-(1 << 31), (1 << 31) - 1, (n, 24), dtype=torch.int32, device=dev, generator=gen)
    data[:, 23] &= 0xFFFF
    ctx.prepare768(n, omega)
    info = ctx.plan_info768(n, omega)

    def step():
        ctx.ntt768_device(data.data_ptr(), n, omega, stream=stream.cuda_stream)
    passes = info["passes"]
    sampler = ClockSampler(dev.index)
    sampler.start()  # nvidia-smi takes ~0.1 s to start: begin before the warm-up so it is live for the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    launches1 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms_step = e0.elapsed_time(e1) / args.steps
    launches2 = ctx.launch_count()

    # ---- e2e: host-pointer public API, H2D + transform + D2H per step
    e2e_steps = max(3, min(args.steps, 10))
    host = torch.empty((n, 24), dtype=torch.int32, pin_memory=True)
    host.copy_(data)
    host_np = host.numpy().view(np.uint32)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) * 1e3 / reps
    e2e_ms = timed(lambda: ctx.best_fft768(host_np, omega), e2e_steps)
    # the same steps issued as ONE batch call: copies of consecutive transforms overlap (full-duplex PCIe)
    hosts = [torch.empty((n, 24), dtype=torch.int32, pin_memory=True) for _ in range(4)]
    for hbuf in hosts:
        hbuf.copy_(data)
    views = [hbuf.numpy().view(np.uint32) for hbuf in hosts]
    e2e_batched_ms = timed(lambda: ctx.best_fft768_batch(views, omega), 3) / len(views)
    # the real drop-in signature hands the library PAGEABLE memory (a std::vector): pinned bounce buffers inside
    pageable = np.array(host_np, copy=True)
    e2e_pageable_ms = timed(lambda: ctx.best_fft768(pageable, omega), e2e_steps)
    clocks = sampler.stop()

    bf = butterflies(logn)
    value = bf / (ms_step * 1e-3)
    h2d = d2h = n * 96
    line = {
        "metric": "768-bit NTT butterflies/s", "value": value, "unit": "butterflies/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": bf / (e2e_ms * 1e-3), "unit": "butterflies/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": e2e_steps, "api": "gsn_ntt768_host (pinned host buffer)"},
        "gpu_launches": launches2 - launches1,
        "roofline": roofline_block(value, 1, logn, ms_step, rates, True, passes, "gsn::ntt768_pass<256,2,false> (CTA-wide) / gsn::ntt768_pass2<1> (warp-owned tiles)",
                                   info["table_bytes"]),
        "plan": info,
        "e2e_batched": {"value": bf / (e2e_batched_ms * 1e-3), "unit": "butterflies/s", "ms_per_step": e2e_batched_ms,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "gsn_ntt768_host_batch, 4 pinned vectors per call: H2D of vector i+1 overlaps passes and D2H of vector i"},
        "e2e_pageable": {"value": bf / (e2e_pageable_ms * 1e-3), "unit": "butterflies/s", "ms_per_step": e2e_pageable_ms,
                         "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "api": "gsn_ntt768_host on pageable memory (what best_fft(std::vector&) passes): column blocks gathered into pinned bounce buffers by host threads, pipelined with the DMA and the passes"},
    }
    secondary = {}
    try:
        runs = []
        for _ in range(2):   # two fresh processes: the very first one on a box also pays for cold page cache / driver state
            out = subprocess.run([sys.executable, "-c", FIRST_CALL_SCRIPT % (ROOT, logn)], capture_output=True, text=True, timeout=180)
            runs.append(json.loads(out.stdout.strip().splitlines()[-1]))
        secondary["first_call"] = dict(min(runs, key=lambda r: r["first_call_ms"]))
        secondary["first_call"]["all_runs_first_call_ms"] = [r["first_call_ms"] for r in runs]
        secondary["first_call"]["what"] = (f"fresh process (best of 2): gsn_ctx_create, then the first and second best_fft (2^{logn}, pageable vector); "
                                           "the first call builds the plan, the pinned bounce buffers and the host copy threads")
    except Exception as e:
        secondary["first_call"] = {"error": str(e)}
    try:
        secondary["pipeline"] = bench_pipeline(ctx, torch, dev, stream, logn)
    except Exception as e:
        secondary["pipeline"] = {"error": str(e)}
    try:   # the N > 1 workload (configs[3], 2^24) on this one GPU: the same-size baseline of the strong-scaling runs
        n24 = 1 << 24
        w24 = F.root_of_unity768(n24)
        d24 = torch.randint(-(1 << 31), (1 << 31) - 1, (n24, 24), dtype=torch.int32, device=dev, generator=gen)
        d24[:, 23] &= 0xFFFF
        ms24 = ctx.time_ntt768(d24.data_ptr(), n24, w24, reps=7)
        m = float(np.median(ms24[2:]))
        secondary["cfg4_2pow24_on_one_gpu"] = {"workload": "MNT4-753 Fr forward NTT n=2^24, device resident, 1xB200 (three passes of 8 stages)", "ms_per_step": m,
                                               "value": butterflies(24) / (m * 1e-3), "unit": "butterflies/s", "plan": ctx.plan_info768(n24, w24)}
        del d24
        ctx.trim()
    except Exception as e:
        secondary["cfg4_2pow24_on_one_gpu"] = {"error": str(e)}
    try:
        secondary["ntt32_cfg2"] = bench_ntt32(ctx, peaks_file().get("hbm_gbs", 6650.0))
    except Exception as e:
        secondary["ntt32_cfg2"] = {"error": str(e)}
    try:
        secondary["g1_multiexp_2pow16"] = bench_g1(ctx)
    except Exception as e:
        secondary["g1_multiexp_2pow16"] = {"error": str(e)}
    line["secondary"] = secondary
    unbind_cpus()
    try:
        line["cpu_baseline"] = cpu_reference_run(min(args.cpu_sample_log_n, logn), 1, 0)
    except Exception as e:  # the bench line must still print
        line["cpu_baseline"] = {"error": str(e)}
    emit(line)


def gather_to_rank0(torch, dist, t, rank, world):
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t.contiguous(), parts, dst=0)
    return parts


def sharded_parity(torch, dist, ctx, F, plan, y, logn, omega, rank, world, dev, spot_budget_s=20.0, check_inverse=False):
    """outside the timed region: rank 0 gathers the shards, runs the SAME input through the single-GPU transform
    (itself oracle-checked at this size by tests/test_gpu_ntt768.py) and compares every word; then K output indices are
    evaluated directly by Horner on the CPU oracle (reference harness shape: device result == host result,
    test/main.cpp:80-84).  Returns (parity dict, single-GPU ms per transform)."""
    n = 1 << logn
    G, rb = world, plan.rank_bit
    n1, n2, C, R = plan.n1, plan.n2, plan.C, plan.R
    xs = gather_to_rank0(torch, dist, plan.x, rank, world)
    ys = gather_to_rank0(torch, dist, y, rank, world)
    res, single_ms = None, None
    if rank == 0:
        a = torch.empty((n1, C >> rb, G, 1 << rb, 24), dtype=torch.int32, device=dev)
        for gidx in range(G):
            a[:, :, gidx] = xs[gidx].view(n1, C >> rb, 1 << rb, 24)
        del xs
        a = a.view(n, 24)
        got = torch.stack(ys, dim=1).view(n, 24)   # A[(h R + r) + n1 k2] = y_h[k2][r]
        del ys
        ref = a.clone()
        ctx.ntt768_device(ref.data_ptr(), n, omega)
        ctx.synchronize()
        same = bool(torch.equal(ref, got))
        # same-size single-GPU time (device resident), for the scaling figure
        ms = ctx.time_ntt768(ref.data_ptr(), n, omega, reps=6)
        single_ms = float(np.median(ms[2:]))
        del ref
        # Horner spot checks on the CPU oracle; the number of indices is sized to a time budget
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        unbind_cpus()
        O.lib().oracle_set_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; the other ranks are idle here
        cal_n = 1 << 16
        cal = np.ascontiguousarray(a[:cal_n].cpu().numpy().view(np.uint32))
        t0 = time.perf_counter()
        O.dft_points768_mt(cal, F.root_of_unity768(cal_n), np.arange(8, dtype=np.uint64))
        per_spot = (time.perf_counter() - t0) / 8 * (n / cal_n)
        spots = int(max(2, min(8, spot_budget_s / max(per_spot, 1e-9))))
        rng = np.random.Generator(np.random.PCG64(2024 + logn))
        ks = np.unique(np.concatenate([[0, 1, n - 1], rng.integers(0, n, size=spots)]).astype(np.uint64))[:max(spots, 3)]
        a_host = np.ascontiguousarray(a.cpu().numpy().view(np.uint32))
        t0 = time.perf_counter()
        exp = O.dft_points768_mt(a_host, omega, ks)
        spot_s = time.perf_counter() - t0
        got_rows = got[torch.from_numpy(ks.astype(np.int64)).to(dev)].cpu().numpy().view(np.uint32)
        spots_ok = bool((exp == got_rows).all())
        res = {"vs_single_gpu": same, "elements_compared": n, "spot_checks": int(len(ks)), "spot_checks_ok": spots_ok,
               "spot_indices": [int(k) for k in ks], "spot_check_seconds": spot_s,
               "how": "rank 0: gathered shards == gsn_ntt768_device on the gathered input, every word; outputs at spot_indices == Horner evaluation by the CPU oracle (oracle_dft_points768_mt)"}
        del a, got
    dist.barrier()
    return res, single_ms


def run_multi(args, torch, dist, g, F, fourstep, ctx, dev, stream, logn, config, rates, rank, world):
    N = world
    n = 1 << logn
    omega = F.root_of_unity768(n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)

    def synth(shape):
        t = torch.randint(-(1 << 31), (1 << 31) - 1, shape, dtype=torch.int32, device=dev, generator=gen)
        t[..., 23] &= 0xFFFF
        return t

    def barrier():
        dist.barrier()
        torch.cuda.synchronize(dev)

    fused = args.exchange == "fused"
    if fused:
        plan = fourstep.FusedFourStepNTT768(ctx, dev, logn, omega, directions=("forward",))
        plan.x.copy_(synth(plan.column_block_shape()))
        data = plan.x
    else:
        plan = fourstep.FourStepNTT768(fourstep.CudaBackend(ctx, dev), logn, omega, directions=("forward",))
        data = synth(plan.column_block_shape())
    state = {}

    def step():
        state["y"] = plan.forward(data)   # fused: plan.x is only read, a row buffer of the plan receives the result
    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches1 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches2 = ctx.launch_count()

    # ---- phase split (fused plan): a few extra calls with per-phase events, outside the timed region
    phases = None
    if fused:
        plan.plan.set_timing(True)
        for _ in range(5):
            step()
        barrier()
        phases = plan.phase_ms()
        plan.plan.set_timing(False)

    # ---- e2e: pinned shard -> H2D -> forward -> D2H, per rank
    e2e_steps = max(3, min(args.steps, 10))
    shard = torch.empty(plan.column_block_shape(), dtype=torch.int32, pin_memory=True)
    shard.copy_(data)
    out_host = torch.empty(plan.row_block_shape(), dtype=torch.int32, pin_memory=True)
    dbuf = data if fused else torch.empty_like(data)

    def e2e_step():
        dbuf.copy_(shard, non_blocking=True)
        y = plan.forward(dbuf)
        out_host.copy_(y, non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    clocks = sampler.stop()

    t = torch.tensor([ms_total, e2e_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    bf = butterflies(logn)
    value = bf / (ms_step * 1e-3)

    # ---- parity of the sharded result (outside the timed region)
    parity, single_ms = None, None
    if fused and not args.no_parity:
        step()
        torch.cuda.synchronize(dev)
        parity, single_ms = sharded_parity(torch, dist, ctx, F, plan, state["y"], logn, omega, rank, world, dev)
    table_bytes = plan.plan.table_bytes if fused else None
    per_source = plan.plan.per_source if fused else None

    # ---- configs[4]: 2^26 forward + inverse on 8 GPUs (secondary), same checks plus the round trip
    cfg5 = None
    want5 = args.cfg5 == "on" or (args.cfg5 == "auto" and N == 8)
    if fused and want5:
        plan.close()
        ctx.trim()
        torch.cuda.empty_cache()
        cfg5 = run_cfg5(args, torch, dist, F, fourstep, ctx, dev, stream, rank, world, synth, barrier)

    if rank == 0:
        line = {
            "metric": "768-bit NTT butterflies/s", "value": value, "unit": "butterflies/s", "n_gpus": N, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": bf / (e2e_ms * 1e-3), "unit": "butterflies/s", "h2d_bytes_per_step": data.numel() * 4, "d2h_bytes_per_step": data.numel() * 4,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "api": "per rank: pinned shard -> H2D -> gsn_fourstep_forward -> D2H"},
            "gpu_launches": launches2 - launches1,
            "roofline": roofline_block(value, N, logn, ms_step, rates, False, None, "gsn::ntt768_pass<256,2,false> + gsn::ntt768_pass2<1> (row pass with arrival flags)", table_bytes),
            "parity": parity,
            "phases_ms": phases,
            "per_source_row_start": per_source,
            "single_gpu_same_size": ({"ms_per_step": single_ms, "value": bf / (single_ms * 1e-3), "speedup": single_ms / ms_step,
                                      "efficiency": single_ms / ms_step / N,
                                      "what": f"the same 2^{logn} transform device resident on rank 0's GPU alone (gsn_ntt768_device), measured in this run"}
                                     if single_ms else None),
        }
        if cfg5 is not None:
            line["secondary"] = {"cfg5_2pow26_forward_inverse": cfg5}
        emit(line)
    ok = True
    if rank == 0 and parity is not None:
        ok = parity["vs_single_gpu"] and parity["spot_checks_ok"]
    if rank == 0 and cfg5 is not None:
        ok = ok and cfg5["parity"]["vs_single_gpu"] and cfg5["parity"]["spot_checks_ok"] and cfg5["parity"]["round_trip"]
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    if fused and not want5:
        plan.close()
    dist.destroy_process_group()
    if int(flag[0]) == 0:
        print("bench.py: PARITY FAILURE (sharded result differs from the single-GPU transform / the oracle)", file=sys.stderr)
        sys.exit(1)


def run_cfg5(args, torch, dist, F, fourstep, ctx, dev, stream, rank, world, synth, barrier):
    logn = 26
    n = 1 << logn
    omega = F.root_of_unity768(n)
    plan = fourstep.FusedFourStepNTT768(ctx, dev, logn, omega, directions=("forward", "inverse"))
    x0 = synth(plan.column_block_shape())
    plan.x.copy_(x0)
    steps = max(2, min(args.steps, 8))

    def pair():
        plan.forward()
        plan.inverse()
    for _ in range(2):
        pair()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        pair()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_pair = float(t[0])
    # every forward + inverse pair must reproduce the input bit for bit, on every rank
    rt = torch.tensor([1 if bool(torch.equal(plan.x, x0)) else 0], device=dev)
    dist.all_reduce(rt, op=dist.ReduceOp.MIN)
    y = plan.forward()
    torch.cuda.synchronize(dev)
    parity, single_fwd_ms = sharded_parity(torch, dist, ctx, F, plan, y, logn, omega, rank, world, dev, spot_budget_s=30.0)
    single_pair_ms = None
    if rank == 0:
        # 1-GPU baseline of the same job: 2^26 forward + inverse, device resident (two-level boundary tables)
        buf = torch.zeros((n, 24), dtype=torch.int32, device=dev)
        buf[:, 0] = 7
        ctx.ntt768_device(buf.data_ptr(), n, omega)
        ctx.ntt768_device(buf.data_ptr(), n, omega, inverse=True)
        ctx.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(2):
            ctx.ntt768_device(buf.data_ptr(), n, omega, stream=stream.cuda_stream)
            ctx.ntt768_device(buf.data_ptr(), n, omega, inverse=True, stream=stream.cuda_stream)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        single_pair_ms = ev0.elapsed_time(ev1) / 2
        del buf
        parity["round_trip"] = bool(int(rt[0]) == 1)
    dist.barrier()
    info = {"table_bytes_per_rank": plan.plan.table_bytes, "per_source_row_start": plan.plan.per_source}
    plan.close()
    if rank != 0:
        return None
    bf2 = 2 * butterflies(logn)
    return {"workload": f"MNT4-753 Fr forward + inverse NTT (omega^-1, n^-1) n=2^26 on {world}xB200, fused four-step (BASELINE.json configs[4])",
            "ms_per_pair": ms_pair, "steps": steps, "value": bf2 / (ms_pair * 1e-3), "unit": "butterflies/s",
            "single_gpu_same_size": {"ms_per_pair": single_pair_ms, "speedup": single_pair_ms / ms_pair, "target": ">= 6x at 8 GPUs",
                                     "what": "2^26 forward + inverse device resident on rank 0's GPU alone, measured in this run"},
            "parity": parity, **info}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=0, help="override the transform size (default 20 at N=1, 24 at N>1)")
    ap.add_argument("--cpu-sample-log-n", type=int, default=20, help="size of the CPU reference sample (2^20 = the whole N=1 workload, ~4 s on 16 cores)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1: fused = gsn_fourstep (peer stores over NVLink + arrival flags); nccl = all_to_all_single + repack (comparison)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the bit-for-bit comparison with the single-GPU transform")
    ap.add_argument("--cfg5", default="auto", choices=["auto", "on", "off"], help="N>1: also run 2^26 forward+inverse (auto: at 8 GPUs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = max(args.gpus, 1)
    logn = args.log_n or (20 if N == 1 else 24)
    workload = (f"MNT4-753 Fr (768-bit) forward NTT n=2^{logn} on 1xB200" if N == 1 else
                f"MNT4-753 Fr (768-bit) forward NTT n=2^{logn}, four-step sharded across {N}xB200, exchange={args.exchange}"
                + (" (column-pass epilogue stores tiles into peer HBM over NVLink; row pass starts per source rank on arrival flags; no NCCL on the data path)"
                   if args.exchange == "fused" else " (NCCL all_to_all_single + repack copy)"))
    config = {"workload": workload, "log_n": logn, "field": "MNT4-753 Fr", "element_bytes": 96,
              "l2": "working set (data + workspace + twiddle table, 3 x n x 96 B) exceeds the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        sl = min(args.cpu_sample_log_n, logn)
        steps = max(1, min(args.steps, 3))  # bounded: each step is seconds of CPU work
        cb = cpu_reference_run(sl, steps, min(args.warmup, 1))
        if sl != logn:
            config = dict(config, cpu_sample=f"the reference's host FFT is timed on a bounded sample of n=2^{sl} and reported as a RATE (butterflies/s); "
                                             f"the 2^{logn} workload itself is not run on the CPU (it would take minutes per step)")
        line = {"impl": "reference", "metric": "768-bit NTT butterflies/s", "value": cb["value"], "unit": "butterflies/s", "n_gpus": N,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "strong" if N > 1 else "weak", "vs_baseline": None, "dtype": DTYPE,
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "butterflies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import gpusnarks_b200 as g
    from gpusnarks_b200 import field as F
    from gpusnarks_b200 import fourstep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gpusnarks_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == N or world == 1, f"--gpus {N} but WORLD_SIZE={world}"

    numa_cores = bind_to_gpu_numa_node(local_rank)
    config["host_affinity"] = f"process bound to the {numa_cores} cores local to its GPU (NVML)" if numa_cores else "not bound"
    ctx = g.Context(local_rank)
    stream = torch.cuda.Stream(dev)  # a real (non-default) stream: the library launches on it, torch events time it
    torch.cuda.set_stream(stream)
    rates = ctx.int32_issue_rates() if rank == 0 else None
    if world == 1:
        run_single(args, torch, g, F, ctx, dev, stream, logn, config, rates)
    else:
        run_multi(args, torch, dist, g, F, fourstep, ctx, dev, stream, logn, config, rates, rank, world)


if __name__ == "__main__":
    main()

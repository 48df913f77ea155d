"""Round-2 features on the GPU, through the C ABI, bit for bit against the oracle:
large-tile kernel variants (warp-owned blocks, wide lazy ranges, twiddle prefetch), two-level (compact) pre-twiddle
tables, per-context field constants, per-stream workspaces, coset transforms in the library, the plan cache bound,
the four-step plan object and the single-process multi-GPU entry point."""
import ctypes as C

import numpy as np
import pytest

import fieldgen
import oracle_lib as O
import pyref

pytestmark = pytest.mark.gpu


def _oracle(a, w, inverse=False):
    n = a.shape[0]
    return O.fft768(a, w, 3 if n >= 64 else (0 if inverse else -1), inverse=inverse)


@pytest.fixture()
def fresh():
    import gpusnarks_b200 as g
    c = g.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("variant", [-1, 1, 5, 4])
def test_kernel_variants_full_parity(fresh, variant):
    """every variant of the 768-bit pass kernel (CTA-wide kernel; warp-owned tiles with lazy ranges, without / with
    the CTA-wide stages 3-4) gives the oracle's bits: 2^18 (two passes of 9 stages: phase B of the warp-owned scheme) and 2^20"""
    fresh.set_option("kernel_variant", variant)
    for logn in (18, 20):
        n = 1 << logn
        a = fieldgen.random_elements(n, 4200 + logn)
        w = fieldgen.omega768(n)
        fwd = fresh.ntt768(a, w)
        assert (fwd == _oracle(a, w)).all(), (variant, logn)
        assert (fresh.ntt768(fwd, w, inverse=True) == a).all(), (variant, logn)


@pytest.mark.parametrize("variant", [1, 5, 4])
@pytest.mark.parametrize("logn,batch", [(1, 1 << 18), (3, 1 << 15), (4, 1 << 14), (7, 1 << 12), (8, 1 << 11), (9, 1 << 10), (10, 1 << 9), (12, 256), (15, 32), (21, 1)])
def test_large_tile_kernel_geometries(fresh, logn, batch, variant):
    """digit widths 1..10 in 1024-element tiles (several sub-transforms per tile, phase A only / phase A + B, with and
    without the CTA-wide stages 3-4), batched, for each kernel that handles 1024-element tiles"""
    fresh.set_option("kernel_variant", variant)
    n = 1 << logn
    a = fieldgen.random_elements(n * batch, 4300 + logn)
    w = fieldgen.omega768(n)
    d = fresh.device_alloc(a.nbytes)
    try:
        fresh.h2d(d, a)
        fresh.ntt768_device(d, n, w, batch=batch)
        out = np.empty_like(a)
        fresh.d2h(out, d)
        for b in sorted({0, 1, batch // 2, batch - 1}):
            assert (out[b * n:(b + 1) * n] == _oracle(a[b * n:(b + 1) * n], w)).all(), (logn, b)
        fresh.ntt768_device(d, n, w, inverse=True, batch=batch)
        fresh.d2h(out, d)
        assert (out == a).all()
    finally:
        fresh.device_free(d)


def test_edge_values_through_lazy_ranges(fresh):
    """all p-1 inputs drive every lazy intermediate to its bound (values up to 36p inside a 10-stage pass)"""
    p = pyref.FR
    for logn in (10, 20):
        n = 1 << logn
        w = fieldgen.omega768(n)
        a = np.ascontiguousarray(np.tile(pyref.ints_to_array([p - 1]), (n, 1)))
        got = fresh.ntt768(a, w)
        assert pyref.from_limbs(got[0]) == (p - 1) * n % p and not got[1:].any()
        e = np.ascontiguousarray(np.tile(fieldgen.edge_elements(), (n // 8, 1))[:n])
        assert (fresh.ntt768(e, w) == _oracle(e, w)).all()


@pytest.mark.parametrize("logn", [12, 16, 20, 22])
def test_two_level_tables_equal_flat_tables(fresh, logn):
    """a plan whose pass-boundary tables exceed the flat-table limit uses the two-level tables (two products per
    element): same bits, a fraction of the memory"""
    n = 1 << logn
    a = fieldgen.random_elements(n, 4400 + logn)
    w = fieldgen.omega768(n)
    flat = fresh.ntt768(a, w)
    info_flat = fresh.plan_info768(n, w)
    assert info_flat["two_level_boundaries"] == 0
    if logn <= 20:
        assert (flat == _oracle(a, w)).all()
    fresh.trim()
    fresh.set_option("flat_table_limit", 1 << 16)
    two = fresh.ntt768(a, w)
    info = fresh.plan_info768(n, w)
    assert info["two_level_boundaries"] >= 1 and info["table_bytes"] < info_flat["table_bytes"] // 8, (info, info_flat)
    assert (two == flat).all()
    assert (fresh.ntt768(two, w, inverse=True) == a).all()   # inverse plan: n^-1 rides on the scaled low table


def test_2pow26_plan_tables_stay_small(fresh):
    """a 2^26 plan held 13 GiB of tables; with the default limit its big boundary is two-level"""
    n = 1 << 26
    w = fieldgen.omega768(n)
    info = fresh.plan_info768(n, w)
    assert info["passes"] == 3 and info["two_level_boundaries"] == 1 and info["table_bytes"] < (1 << 30), info


def test_fields_are_per_context(fresh):
    """two contexts on one device with different fields must not disturb each other"""
    import gpusnarks_b200 as g
    other = g.Context(0)
    try:
        other.set_field768(g.FIELD_FQ)
        n = 1 << 10
        a_fr = fieldgen.random_elements(n, 1, pyref.FR)
        a_fq = fieldgen.random_elements(n, 2, pyref.FQ)
        w_fr = fieldgen.omega768(n, pyref.FR, 17)
        w_fq = fieldgen.omega768(n, pyref.FQ, 13)
        got_fr = fresh.ntt768(a_fr, w_fr)
        got_fq = other.ntt768(a_fq, w_fq)      # would run with the Fr modulus under a device-wide constant
        third = g.Context(0)                   # creating another context must not reset anything either
        got_fq2 = other.ntt768(a_fq, w_fq)
        got_fr2 = fresh.ntt768(a_fr, w_fr)
        third.close()
        assert (got_fr == O.fft768(a_fr, w_fr, 3)).all() and (got_fr2 == got_fr).all()
        O.set_field768("fq")
        try:
            assert (got_fq == O.fft768(a_fq, w_fq, 3)).all() and (got_fq2 == got_fq).all()
        finally:
            O.set_field768("fr")
    finally:
        other.close()


def test_two_streams_share_a_context(fresh):
    """device-pointer calls on different caller streams must not share the scratch buffer"""
    import torch
    dev = torch.device("cuda", 0)
    n = 1 << 20   # two passes: pass 1 writes the scratch buffer
    w = fieldgen.omega768(n)
    a = fieldgen.random_elements(n, 4500)
    b = fieldgen.random_elements(n, 4501)
    ta = torch.from_numpy(a.view(np.int32)).to(dev)
    tb = torch.from_numpy(b.view(np.int32)).to(dev)
    fresh.prepare768(n, w)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    for _ in range(3):   # forward on both streams at once, then inverse: any scratch sharing corrupts one of them
        fresh.ntt768_device(ta.data_ptr(), n, w, stream=s1.cuda_stream)
        fresh.ntt768_device(tb.data_ptr(), n, w, stream=s2.cuda_stream)
        fresh.ntt768_device(ta.data_ptr(), n, w, inverse=True, stream=s1.cuda_stream)
        fresh.ntt768_device(tb.data_ptr(), n, w, inverse=True, stream=s2.cuda_stream)
    torch.cuda.synchronize()
    assert (ta.cpu().numpy().view(np.uint32) == a).all() and (tb.cpu().numpy().view(np.uint32) == b).all()


@pytest.mark.parametrize("logn,limit", [(3, None), (10, None), (13, None), (16, None), (16, 1 << 12), (20, None), (21, 1 << 12)])
def test_coset_transforms_in_the_library(fresh, logn, limit):
    """gsn_coset_ntt768_*: forward = transform of a[i] * g^i (pre-twiddle of pass 1), inverse = its inverse with
    n^-1 g^-i as the post-twiddle of the last pass; flat and two-level shift tables"""
    p = pyref.FR
    if limit:
        fresh.set_option("flat_table_limit", limit)
    n = 1 << logn
    a = fieldgen.random_elements(n, 4600 + logn)
    w = fieldgen.omega768(n)
    g_int = 17
    shift = pyref.ints_to_array([pyref.mont(g_int)])[0]
    ev = fresh.coset_ntt768(a, w, shift)
    if logn <= 16:
        pw, cur = [], 1
        for _ in range(n):
            pw.append(pyref.mont(cur))
            cur = cur * g_int % p
        assert (ev == O.fft768(O.fp768_binop("mul", a, pyref.ints_to_array(pw)), w, 3 if n >= 64 else -1)).all()
    else:   # spot values: ev[k] = sum a[i] (g w^k)^i = plain DFT points of the twisted input, via linearity in g
        ks = np.array([0, 1, n // 2 + 3, n - 1], dtype=np.uint64)
        d = fresh.device_alloc(a.nbytes)
        t = fresh.device_alloc(a.nbytes)
        try:
            fresh.h2d(d, a)
            fresh.fp768_powers_device(t, n, shift)
            fresh.fp768_binop_device("mul", d, d, t, n)
            tw = np.empty_like(a)
            fresh.d2h(tw, d)
        finally:
            fresh.device_free(d)
            fresh.device_free(t)
        assert (ev[ks.astype(np.int64)] == O.dft_points768(tw, w, ks)).all()
    assert (fresh.coset_ntt768(ev, w, shift, inverse=True) == a).all()


def test_prover_pipeline_device_resident(fresh):
    """iFFT -> coset FFT -> pointwise product -> coset iFFT, all device resident: the quotient
    pipeline of a prover on polynomials a, b of degree < n/2: c = a * b computed through coset evaluations equals the
    schoolbook product"""
    p = pyref.FR
    logn = 8
    n = 1 << logn
    rinv = pow(pyref.RMONT, -1, p)
    ca = [pyref.from_limbs(x) * rinv % p for x in fieldgen.random_elements(n // 2, 4701)]
    cb = [pyref.from_limbs(x) * rinv % p for x in fieldgen.random_elements(n // 2, 4702)]
    w = fieldgen.omega768(n)
    shift = pyref.ints_to_array([pyref.mont(17)])[0]
    wi = pyref.root_of_unity(p, 17, n)
    ea = pyref.ints_to_array([pyref.mont(sum(c * pow(wi, i * k, p) for i, c in enumerate(ca)) % p) for k in range(n)])
    eb = pyref.ints_to_array([pyref.mont(sum(c * pow(wi, i * k, p) for i, c in enumerate(cb)) % p) for k in range(n)])
    da, db = fresh.device_alloc(n * 96), fresh.device_alloc(n * 96)
    try:
        fresh.h2d(da, ea)
        fresh.h2d(db, eb)
        for d in (da, db):
            fresh.ntt768_device(d, n, w, inverse=True)          # evaluations on <w> -> coefficients
            fresh.coset_ntt768_device(d, n, w, shift)           # coefficients -> evaluations on g<w>
        fresh.fp768_binop_device("mul", da, da, db, n)          # pointwise
        fresh.coset_ntt768_device(da, n, w, shift, inverse=True)
        out = np.empty((n, 24), dtype=np.uint32)
        fresh.d2h(out, da)
    finally:
        fresh.device_free(da)
        fresh.device_free(db)
    prod = [0] * n
    for i, x in enumerate(ca):
        for j, y in enumerate(cb):
            prod[i + j] = (prod[i + j] + x * y) % p
    assert pyref.array_to_ints(out) == [pyref.mont(c) for c in prod]


def test_plan_cache_is_bounded(fresh):
    n = 1 << 12
    fresh.set_option("plan_cache_bytes", 3 * (1 << 12) * 192)   # room for about two 2^12 plans
    p = pyref.FR
    base = pyref.root_of_unity(p, 17, n)
    infos = []
    for k in (1, 3, 5, 7, 9):    # five different primitive roots -> five plans
        w = np.array(pyref.to_limbs(pow(base, k, p) * pyref.RMONT % p), dtype=np.uint32)
        a = fieldgen.random_elements(n, 4800 + k)
        assert (fresh.ntt768(a, w) == _oracle(a, w)).all()
        infos.append(fresh.plan_info768(n, w))
    assert infos[-1]["cached_plans"] <= 3 and infos[-1]["cached_bytes"] <= 3 * (1 << 12) * 192 + 4096, infos


@pytest.mark.parametrize("logn,threads", [(16, "1"), (18, "3"), (20, None), (21, "8")])
def test_pageable_host_path_equals_device_path(logn, threads, monkeypatch):
    """best_fft(std::vector&) hands the library pageable memory: gsn_ntt768_host then gathers column blocks into pinned
    bounce buffers with a few host threads, pipelined with the DMA and the passes (2-pass plans at 2^16..2^20, a 3-pass
    plan at 2^21).  Same bits as the device-resident call (itself oracle-checked), forward and inverse."""
    import gpusnarks_b200 as g
    if threads is None:
        monkeypatch.delenv("GSN_HOST_THREADS", raising=False)
    else:
        monkeypatch.setenv("GSN_HOST_THREADS", threads)
    ctx = g.Context(0)
    try:
        n = 1 << logn
        w = fieldgen.omega768(n)
        a = fieldgen.random_elements(n, 7100 + logn)       # a numpy array: pageable
        d = ctx.device_alloc(a.nbytes)
        for inverse in (False, True):
            ctx.h2d(d, a)
            ctx.ntt768_device(d, n, w, inverse=inverse)
            want = np.empty_like(a)
            ctx.d2h(want, d)
            got = ctx.ntt768(a, w, inverse=inverse)
            assert (got == want).all(), (logn, inverse)
        if logn <= 16:
            assert (want == _oracle(a, w, inverse=True)).all()
        ctx.device_free(d)
    finally:
        ctx.close()


@pytest.mark.parametrize("logn", [6, 12, 16, 21])
def test_fourstep_plan_single_rank(fresh, logn):
    """gsn_fourstep with one rank: column transforms + (self) scatter + row transforms == the transform, both directions"""
    import torch
    from gpusnarks_b200 import fourstep
    dev = torch.device("cuda", 0)
    n = 1 << logn
    a = fieldgen.random_elements(n, 4900 + logn)
    w = fieldgen.omega768(n)
    plan = fourstep.FusedFourStepNTT768(fresh, dev, logn, w)
    try:
        x0 = torch.from_numpy(fourstep.to_column_layout(a, logn, 1, 0).view(np.int32)).to(dev)
        plan.x.copy_(x0)
        y = plan.forward()
        torch.cuda.synchronize()
        got = fourstep.from_row_layouts([y.cpu().numpy().view(np.uint32)], logn)
        ref = fresh.ntt768(a, w)
        assert (got == ref).all()
        if logn <= 16:
            assert (got == _oracle(a, w)).all()
        back = plan.inverse()
        torch.cuda.synchronize()
        assert bool((back == x0).all())
        # a second round trip exercises the alternate row buffer and the epochs
        y = plan.forward()
        back = plan.inverse()
        torch.cuda.synchronize()
        assert bool((back == x0).all())
    finally:
        plan.close()


@pytest.mark.parametrize("logn,pinned", [(14, False), (21, False), (21, True)])
def test_multi_gpu_entry_point_one_process(logn, pinned):
    """gsn_multi_*: every visible device (1, 2, 4 or 8) driven from this one process; host vector in natural order.
    A pageable vector (numpy / std::vector) of 4 MiB and more is staged through pinned bounce buffers by host threads,
    a pinned one is copied by strided 2-D DMAs directly: both forms are checked."""
    import torch
    import gpusnarks_b200 as g
    from gpusnarks_b200.ntt import MultiGpu
    cnt = g.device_count()
    G = 1
    while G * 2 <= min(cnt, 8):
        G *= 2
    n = 1 << logn
    a = fieldgen.random_elements(n, 5000 + logn)
    w = fieldgen.omega768(n)
    m = MultiGpu(list(range(G)), n, w)
    try:
        if pinned:
            keep = torch.from_numpy(a.view(np.int32)).clone().pin_memory()
            v = keep.numpy().view(np.uint32)
        else:
            v = a.copy()
        m.ntt_host(v)
        c = g.Context(0)
        try:
            assert (v == c.ntt768(a, w)).all(), f"{G} devices"
        finally:
            c.close()
        if logn <= 16:
            assert (v == _oracle(a, w)).all()
        m.ntt_host(v, inverse=True)
        assert (v == a).all()
    finally:
        m.close()

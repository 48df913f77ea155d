"""32-bit prime-field NTT on the GPU (through the C ABI) against the CPU oracle, bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest

import fieldgen
import oracle_lib as O
import pyref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P = pyref.P32


@pytest.mark.parametrize("logn", list(range(0, 21)))
def test_forward_and_inverse_vs_oracle(ctx, logn):
    n = 1 << logn
    a = fieldgen.random_u32(n, 500 + logn, P)
    w = fieldgen.omega32(n)
    lc = 3 if n >= 64 else -1
    fwd = ctx.ntt32(a, w, P)
    assert (fwd == O.fft32(a, w, P, lc)).all(), f"forward 2^{logn}"
    inv = ctx.ntt32(a, w, P, inverse=True)
    assert (inv == O.fft32(a, w, P, max(lc, 0), inverse=True)).all(), f"inverse 2^{logn}"
    assert (ctx.ntt32(fwd, w, P, inverse=True) == a).all()


def test_cfg1_and_cfg2_sizes(ctx):
    """BASELINE.json configs[0] (2^16, the reference's CPU-runnable case) and configs[1] (2^22)"""
    for logn in (16, 22):
        n = 1 << logn
        a = fieldgen.random_u32(n, 1, P)
        w = fieldgen.omega32(n)
        got = ctx.ntt32(a, w, P)
        assert (got == O.fft32(a, w, P, 3)).all(), logn
        assert (ctx.ntt32(got, w, P, inverse=True) == a).all()


@pytest.mark.parametrize("mod,gen", [(998244353, 3), (469762049, 3), (2013265921, 31), (7340033, 3), (97, 5)])
def test_other_primes(ctx, mod, gen):
    s = pyref.two_adicity(mod)
    for logn in sorted({1, min(s, 5), min(s, 13), min(s, 18)}):
        n = 1 << logn
        a = fieldgen.random_u32(n, logn, mod)
        w = pyref.root_of_unity(mod, gen, n)
        assert (ctx.ntt32(a, w, mod) == O.fft32(a, w, mod, -1)).all(), (mod, logn)
        assert (ctx.ntt32(ctx.ntt32(a, w, mod), w, mod, inverse=True) == a).all()


def test_edge_values(ctx):
    n = 1 << 12
    w = fieldgen.omega32(n)
    for name, a in {"zeros": np.zeros(n, np.uint32), "pm1": np.full(n, P - 1, np.uint32),
                    "impulse": np.eye(1, n, 0, dtype=np.uint32)[0], "ramp": (np.arange(n, dtype=np.uint64) * 77773 % P).astype(np.uint32)}.items():
        a = np.ascontiguousarray(a)
        assert (ctx.ntt32(a, w, P) == O.fft32(a, w, P, 2)).all(), name


def test_batched_device(ctx):
    for logn, batch in [(4, 7), (11, 3), (12, 2), (16, 3), (17, 5), (22, 2)]:
        n = 1 << logn
        a = fieldgen.random_u32(batch * n, 9 + logn, P)
        w = fieldgen.omega32(n)
        d = ctx.device_alloc(a.nbytes)
        try:
            ctx.h2d(d, a)
            ctx.ntt32_device(d, n, w, P, batch=batch)
            got = np.empty_like(a)
            ctx.d2h(got, d)
        finally:
            ctx.device_free(d)
        for b in range(batch):
            assert (got[b * n:(b + 1) * n] == O.fft32(a[b * n:(b + 1) * n], w, P, 3 if n >= 64 else -1)).all(), (logn, batch, b)


def test_error_codes(ctx):
    import gpusnarks_b200 as g
    with pytest.raises(g.GsnError) as e:
        ctx.ntt32(np.zeros(12, np.uint32), 1, P)
    assert e.value.code == 2
    with pytest.raises(g.GsnError) as e:
        ctx.ntt32(np.zeros(16, np.uint32), fieldgen.omega32(8), P)
    assert e.value.code == 4
    with pytest.raises(g.GsnError) as e:
        ctx.ntt32(np.zeros(16, np.uint32), 3, 2013265923)  # not prime
    assert e.value.code == 7
    with pytest.raises(g.GsnError) as e:
        ctx.ntt32(np.zeros(1 << 20, np.uint32), 3, 7340033 * 0 + 97)  # n does not divide mod-1
    assert e.value.code == 3


def test_golden_fixtures(ctx):
    with open(os.path.join(GOLDEN, "ntt32.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        n = 1 << case["logn"]
        a = fieldgen.random_u32(n, case["seed"], case["mod"])
        out = ctx.ntt32(a, fieldgen.omega32(n, case["mod"]), case["mod"], inverse=case["inverse"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"], case


@pytest.mark.parametrize("logn", [23, 24, 25, 26, 27])
def test_large_properties(ctx, logn):
    """two-pass (<= 2^24) and three-pass (2^25..2^27, digits of 8 and 9 stages) fast-path sizes"""
    n = 1 << logn
    a = fieldgen.random_u32(n, 70 + logn, P)
    w = fieldgen.omega32(n)
    fwd = ctx.ntt32(a, w, P)
    ks = np.array([0, 1, n - 1, n // 2, 54321, (3 * n) // 4 + 3], dtype=np.uint64)
    assert (fwd[ks.astype(np.int64)] == O.dft_points32(a, w, P, ks)).all()
    assert (ctx.ntt32(fwd, w, P, inverse=True) == a).all()

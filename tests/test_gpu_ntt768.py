"""768-bit NTT on the GPU (through the C ABI) against the CPU oracle, bit for bit."""
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


def _oracle(a, w, inverse=False):
    n = a.shape[0]
    log_cpus = 3 if n >= 64 else -1
    return O.fft768(a, w, log_cpus if not inverse else max(log_cpus, 0), inverse=inverse)


@pytest.mark.parametrize("logn", list(range(0, 15)))
def test_forward_and_inverse_vs_oracle(ctx, logn):
    n = 1 << logn
    a = fieldgen.random_elements(n, 100 + logn)
    w = fieldgen.omega768(n)
    fwd = ctx.ntt768(a, w)
    assert (fwd == _oracle(a, w)).all(), f"forward 2^{logn}"
    inv = ctx.ntt768(a, w, inverse=True)
    assert (inv == _oracle(a, w, inverse=True)).all(), f"inverse 2^{logn}"
    assert (ctx.ntt768(fwd, w, inverse=True) == a).all(), f"roundtrip 2^{logn}"


def test_reference_test_shape_constant_input(ctx):
    """reference test/main.cpp:38-84: 2^16 elements all equal to Scalar(1234), GPU vs host FFT."""
    n = 1 << 16
    a = np.zeros((n, 24), dtype=np.uint32)
    a[:, 0] = 1234
    w = fieldgen.omega768(n)
    got = ctx.ntt768(a, w)
    assert (got == _oracle(a, w)).all()
    # DFT of a constant: n*c at index 0, zero elsewhere
    assert pyref.from_limbs(got[0]) == (1234 * n) % pyref.FR
    assert not got[1:].any()


def test_edge_values(ctx):
    """all-zero, all p-1, unit impulse, and values that sit just below p"""
    n = 1 << 11
    w = fieldgen.omega768(n)
    p = pyref.FR
    cases = {
        "zeros": np.zeros((n, 24), dtype=np.uint32),
        "pm1": np.tile(pyref.ints_to_array([p - 1]), (n, 1)),
        "impulse": np.concatenate([pyref.ints_to_array([pyref.RMONT % p]), np.zeros((n - 1, 24), dtype=np.uint32)]),
        "edges": np.tile(fieldgen.edge_elements(), (n // 8, 1))[:n],
    }
    for name, a in cases.items():
        a = np.ascontiguousarray(a)
        assert (ctx.ntt768(a, w) == _oracle(a, w)).all(), name


def test_batched_and_strided_device(ctx):
    """batch of transforms and the partial (inner stride 2^log_r) form used by the four-step driver"""
    for logn, batch, log_r in [(6, 5, 0), (10, 3, 0), (11, 2, 0), (5, 2, 3), (9, 1, 2), (12, 1, 1), (3, 8, 4)]:
        n, R = 1 << logn, 1 << log_r
        a = fieldgen.random_elements(batch * n * R, 7 + logn)
        w = fieldgen.omega768(n)
        d = ctx.device_alloc(a.nbytes)
        try:
            ctx.h2d(d, a)
            ctx.ntt768_device(d, n, w, batch=batch, log_r=log_r)
            got = np.empty_like(a)
            ctx.d2h(got, d)
        finally:
            ctx.device_free(d)
        v = a.reshape(batch, n, R, 24)
        exp = np.empty_like(v)
        for b in range(batch):
            for r in range(R):
                exp[b, :, r] = O.fft768(np.ascontiguousarray(v[b, :, r]), w, -1)
        assert (got.reshape(exp.shape) == exp).all(), (logn, batch, log_r)


@pytest.mark.parametrize("logn,count", [(6, 5), (12, 3), (18, 4), (20, 3)])
def test_host_batch_matches_single_calls(ctx, logn, count):
    """gsn_ntt768_host_batch (overlapped copies, two staging buffers) == one gsn_ntt768_host call per vector"""
    n = 1 << logn
    w = fieldgen.omega768(n)
    vecs = [fieldgen.random_elements(n, 1000 + 10 * logn + i) for i in range(count)]
    exp = [ctx.ntt768(v, w) for v in vecs]
    got = [v.copy() for v in vecs]
    ctx.best_fft768_batch(got, w)
    for g_, e in zip(got, exp):
        assert (g_ == e).all()
    ctx.best_fft768_batch(got, w, inverse=True)
    for g_, v in zip(got, vecs):
        assert (g_ == v).all()


def test_fq_field(ctx):
    """the reference's literal modulus (MNT4-753 Fq, 2-adicity 15)"""
    import gpusnarks_b200 as g
    ctx.set_field768(g.FIELD_FQ)
    O.set_field768("fq")
    try:
        for logn in (4, 10, 13, 15):
            n = 1 << logn
            a = fieldgen.random_elements(n, 55 + logn, pyref.FQ)
            w = fieldgen.omega768(n, pyref.FQ, 13)
            assert (ctx.ntt768(a, w) == _oracle(a, w)).all(), logn
        with pytest.raises(g.GsnError) as e:
            ctx.ntt768(np.zeros((1 << 16, 24), dtype=np.uint32), fieldgen.omega768(1 << 15, pyref.FQ, 13))
        assert e.value.code == 3  # TOO_LARGE: no 2^16-th root of unity in Fq
    finally:
        ctx.set_field768(g.FIELD_FR)
        O.set_field768("fr")


def test_error_codes(ctx):
    import gpusnarks_b200 as g
    w8 = fieldgen.omega768(8)
    with pytest.raises(g.GsnError) as e:
        ctx.ntt768(np.zeros((12, 24), dtype=np.uint32), w8)
    assert e.value.code == 2  # NOT_POW2
    with pytest.raises(g.GsnError) as e:
        ctx.ntt768(np.zeros((16, 24), dtype=np.uint32), w8)  # omega of the wrong order
    assert e.value.code == 4  # BAD_OMEGA
    with pytest.raises(g.GsnError) as e:
        ctx.ntt768(np.zeros((8, 24), dtype=np.uint32), np.full(24, 0xFFFFFFFF, dtype=np.uint32))
    assert e.value.code == 4


def test_golden_fixtures(ctx):
    """seeded inputs -> SHA-256 of the output limbs, pinned in tests/golden/ntt768.json
    (generated by tests/golden/make_golden.py from the CPU oracle)"""
    with open(os.path.join(GOLDEN, "ntt768.json")) as f:
        gold = json.load(f)
    for case in gold["cases"]:
        n = 1 << case["logn"]
        a = fieldgen.random_elements(n, case["seed"])
        w = fieldgen.omega768(n)
        out = ctx.ntt768(a, w, inverse=case["inverse"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == case["sha256"], case


@pytest.mark.parametrize("logn", [16, 18])
def test_mid_sizes_vs_oracle(ctx, logn):
    n = 1 << logn
    a = fieldgen.random_elements(n, 300 + logn)
    w = fieldgen.omega768(n)
    assert (ctx.ntt768(a, w) == O.fft768(a, w, 3)).all()


def test_cfg3_2pow20_full_oracle(ctx):
    """BASELINE.json configs[2]: MNT4-753 Fr forward NTT 2^20, full comparison with the oracle"""
    n = 1 << 20
    a = fieldgen.random_elements(n, 20)
    w = fieldgen.omega768(n)
    got = ctx.ntt768(a, w)
    log_cpus = max(0, min(6, int(np.log2(max(1, O.lib().oracle_max_threads())))))
    assert (got == O.fft768(a, w, log_cpus)).all()
    assert (ctx.ntt768(got, w, inverse=True) == a).all()


@pytest.mark.parametrize("logn", [22, 24])
def test_large_properties(ctx, logn):
    """sizes the oracle cannot finish in seconds: spot outputs by Horner on the CPU,
    inverse(forward(x)) == x, and linearity F(x + y) == F(x) + F(y)."""
    n = 1 << logn
    a = fieldgen.random_elements(n, 40 + logn)
    w = fieldgen.omega768(n)
    fwd = ctx.ntt768(a, w)
    ks = np.array([0, 1, n - 1, n // 2, n // 2 + 1, 12345 % n, (3 * n) // 4 + 7, 1 << (logn // 2)], dtype=np.uint64)
    assert (fwd[ks.astype(np.int64)] == O.dft_points768(a, w, ks)).all()
    assert (ctx.ntt768(fwd, w, inverse=True) == a).all()
    if logn <= 22:
        b = fieldgen.random_elements(n, 41 + logn)
        s = ctx.fp768_binop("add", a, b)
        assert (ctx.ntt768(s, w) == ctx.fp768_binop("add", fwd, ctx.ntt768(b, w))).all()


def test_2pow27_device_round_trip(ctx):
    """12.9 GB of elements, three passes (9+9+9), more than 2^32 limbs: generated, transformed, inverted and compared on the
    device (torch only allocates and compares).  inverse(forward(x)) == x bit for bit, and forward(x) != x."""
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    if free < 48 * (1 << 30):
        pytest.skip("needs ~40 GB of free device memory (12.9 GB of data x3; the plans' big boundary uses the two-level tables, a few MB)")
    logn = 27
    n = 1 << logn
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(27)
    x = torch.randint(-(1 << 31), (1 << 31) - 1, (n, 24), dtype=torch.int32, device=dev, generator=gen)
    x[:, 23] &= 0xFFFF   # < 2^752 < r: canonical
    x0 = x.clone()
    w = fieldgen.omega768(n)
    stream = torch.cuda.current_stream(dev).cuda_stream or 1
    try:
        ctx.ntt768_device(x.data_ptr(), n, w, stream=stream)
        torch.cuda.synchronize()
        assert not bool((x[:4096] == x0[:4096]).all())
        # two spot outputs by the definition on a subsample would cost n products each on the CPU; instead check the
        # DC coefficient on the device: A[0] = sum_j a[j] is not available without field adds, so rely on the round trip
        ctx.ntt768_device(x.data_ptr(), n, w, inverse=True, stream=stream)
        torch.cuda.synchronize()
        assert bool((x == x0).all())
        # absolute check at this size: the transform of c * delta_{j0} is A[k] = c * omega^(j0 * k); j0 and the probed k
        # sit beyond limb offset 2^32.  omega^(j0*k) comes from the oracle's square-and-multiply.
        j0 = n - 12345
        c_el = fieldgen.random_elements(1, 99)[0]
        x.zero_()
        x[j0] = torch.from_numpy(c_el.view(np.int32)).to(dev)
        ctx.ntt768_device(x.data_ptr(), n, w, stream=stream)
        torch.cuda.synchronize()
        ks = [0, 1, n - 1, n // 2 + 3, (3 * n) // 4 + 77, 190000001 % n, n - 4096]
        got = x[torch.tensor(ks, device=dev)].cpu().numpy().view(np.uint32)
        for row, k in zip(got, ks):
            tw = O.fp768_pow(w, (j0 * k) % n)
            assert (row == O.fp768_binop("mul", c_el[None, :], tw[None, :])[0]).all(), k
    finally:
        del x, x0
        ctx.trim()
        torch.cuda.empty_cache()

"""SURVEY.md section 8f rows built on the same device arithmetic: device-resident element-wise
ops, powers table / coset transforms, and the reference's multiexp<Scalar, Scalar> (a field inner
product, reference test/multiexp.h:3-13 is the CPU form) -- all against the oracle, bit for bit."""
import numpy as np
import pytest

import fieldgen
import oracle_lib as O
import pyref

pytestmark = pytest.mark.gpu


def _sum_products_oracle(a, b):
    prod = O.fp768_binop("mul", a, b)
    acc = np.zeros((1, 24), dtype=np.uint32)
    for i in range(prod.shape[0]):   # reference test/multiexp.h: result = result + a[i] * b[i]
        acc = O.fp768_binop("add", acc, prod[i:i + 1])
    return acc[0]


@pytest.mark.parametrize("count", [0, 1, 2, 127, 128, 129, 5000, 1 << 16])
def test_multiexp_scalar_inner_product(ctx, count):
    a = fieldgen.random_elements(count, 800 + count) if count else np.zeros((0, 24), np.uint32)
    b = fieldgen.random_elements(count, 900 + count) if count else np.zeros((0, 24), np.uint32)
    got = ctx.multiexp768(a, b)
    if count <= 5000:
        exp = _sum_products_oracle(a, b) if count else np.zeros(24, np.uint32)
    else:  # big-int check of the sum (the oracle add loop is slow in Python)
        p = pyref.FR
        rinv = pow(pyref.RMONT, -1, p)
        exp = np.array(pyref.to_limbs(sum(x * y for x, y in zip(pyref.array_to_ints(a), pyref.array_to_ints(b))) * rinv % p), dtype=np.uint32)
    assert (got == exp).all()


def test_reference_multiexp_shape(ctx):
    """reference test/main.cpp:89-133: 2^18 elements, all Scalar(1234) x Scalar(1234) -> n * 1234^2 * R^-1"""
    n = 1 << 18
    a = np.zeros((n, 24), dtype=np.uint32)
    a[:, 0] = 1234
    got = pyref.from_limbs(ctx.multiexp768(a, a))
    p = pyref.FR
    assert got == n * 1234 * 1234 * pow(pyref.RMONT, -1, p) % p


def test_powers_and_device_binop(ctx):
    p = pyref.FR
    n = 3000
    g = pyref.ints_to_array([pyref.mont(7)])[0]
    sc = pyref.ints_to_array([pyref.mont(12345)])[0]
    d = ctx.device_alloc(n * 96)
    try:
        ctx.fp768_powers_device(d, n, g, scale=sc)
        out = np.empty((n, 24), dtype=np.uint32)
        ctx.d2h(out, d)
        assert pyref.array_to_ints(out) == [pyref.mont(12345 * pow(7, i, p) % p) for i in range(n)]
        a = fieldgen.random_elements(n, 5)
        da = ctx.device_alloc(n * 96)
        ctx.h2d(da, a)
        for op in ("mul", "add", "sub"):
            ctx.fp768_binop_device(op, d, da, da, n)
            ctx.d2h(out, d)
            assert (out == O.fp768_binop(op, a, a)).all(), op
        ctx.device_free(da)
    finally:
        ctx.device_free(d)


@pytest.mark.parametrize("logn", [3, 10, 13])
def test_coset_transform_roundtrip_and_definition(ctx, logn):
    p = pyref.FR
    n = 1 << logn
    a = fieldgen.random_elements(n, 600 + logn)
    w = fieldgen.omega768(n)
    shift_int = 17  # libff's coset shift = the multiplicative generator
    shift = pyref.ints_to_array([pyref.mont(shift_int)])[0]
    ev = ctx.coset_ntt768(a, w, shift)
    # definition: plain transform of a[i] * shift^i
    pw = pyref.ints_to_array([pyref.mont(pow(shift_int, i, p)) for i in range(n)])
    assert (ev == O.fft768(O.fp768_binop("mul", a, pw), w, -1)).all()
    assert (ctx.coset_ntt768(ev, w, shift, inverse=True) == a).all()


def test_c_abi_from_plain_c_on_device(tmp_path):
    """tests/c/capi_smoke.c (C99) against the library on the GPU: an 8-point 32-bit transform equals the DFT definition"""
    import test_host_cpp
    import subprocess
    exe = test_host_cpp._build_c_smoke(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "capi ok (device)" in out.stdout, out.stdout + out.stderr

"""The oracle is only as good as what pins it.  These tests (CPU only) check oracle/ against
  * Python big-int arithmetic and the DFT definition (tests/pyref.py);
  * the reference's own unit-test KATs (cuda/device_field_operator_test.cpp:222-320);
  * the REFERENCE ITSELF compiled from /root/reference into oracle/_ref/ (oracle/Makefile):
      libref_verbatim.so  unmodified headers -- add/sub must agree with the oracle; its multiply
                          and its FFT must NOT, which is recorded here so the
                          divergence is documented rather than hidden (Oracle-V fingerprint);
      libref_patched.so   the same headers after three one-line sed corrections (Montgomery
                          constant, normalise size, butterfly statement) -- multiply must agree
                          with the oracle on the reference's modulus, and the reference's FFT
                          templates instantiated over the oracle's field must agree bit for bit.
The reference has no golden vectors for the FFT (its only check is GPU == host on a constant
input, test/main.cpp:80-84): FFT parity is pinned by the last bullet plus the DFT definition."""
import ctypes as C
import os
import random

import numpy as np
import pytest

import fieldgen
import oracle_lib as O
import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def _ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", f"libref_{name}.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/cuda"):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    L = C.CDLL(path)
    L.ref_scalar_binop.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t]
    L.ref_fft_scalar.argtypes = [_u32p, C.c_size_t, _u32p, C.c_int, C.c_int]
    L.ref_fft_scalar.restype = C.c_double
    L.ref_fft_over_oracle768.argtypes = [_u32p, C.c_size_t, _u32p, C.c_int, C.c_int]
    L.ref_fft_over_oracle32.argtypes = [_u32p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int]
    L.ref_mod.argtypes = [_u32p]
    return L


@pytest.fixture(autouse=True)
def _fr():
    O.set_field768("fr")
    yield
    O.set_field768("fr")


# ---------------------------------------------------------------- oracle vs big-int
@pytest.mark.parametrize("field,p", [("fr", pyref.FR), ("fq", pyref.FQ)])
def test_field_ops_vs_bigint(field, p):
    O.set_field768(field)
    rnd = random.Random(1)
    xs = [rnd.randrange(p) for _ in range(300)] + [0, 1, p - 1, pyref.RMONT % p, p - 1, 0]
    ys = [rnd.randrange(p) for _ in range(300)] + [p - 1, p - 1, p - 1, 1, 1, 0]
    A, B = pyref.ints_to_array(xs), pyref.ints_to_array(ys)
    rinv = pow(pyref.RMONT, -1, p)
    assert pyref.array_to_ints(O.fp768_binop("mul", A, B)) == [x * y * rinv % p for x, y in zip(xs, ys)]
    assert pyref.array_to_ints(O.fp768_binop("add", A, B)) == [(x + y) % p for x, y in zip(xs, ys)]
    assert pyref.array_to_ints(O.fp768_binop("sub", A, B)) == [(x - y) % p for x, y in zip(xs, ys)]
    x = xs[0]
    assert pyref.from_limbs(O.fp768_pow(A[0], 12345)) == pow(x * rinv, 12345, p) * pyref.RMONT % p
    assert pyref.from_limbs(O.fp768_pow(A[0], 0)) == pyref.RMONT % p
    assert pyref.from_limbs(O.fp768_inverse(A[0])) == pow(x * rinv, -1, p) * pyref.RMONT % p


def test_reference_unit_test_kats():
    """cuda/device_field_operator_test.cpp: testAdd :222-231, test_subtract :233-244, testMultiply
    :246-266 (1234^2 = 1522756 through to/from Montgomery), testPow :278-299 (2^0, 2^2, 4^10, 2^20, 2^35)"""
    p = pyref.FR
    one = lambda v: pyref.ints_to_array([v])
    assert pyref.array_to_ints(O.fp768_binop("add", one(1234), one(1234))) == [2468]
    assert pyref.array_to_ints(O.fp768_binop("sub", one(1234), one(1234))) == [0]
    assert pyref.array_to_ints(O.fp768_binop("sub", one(1235), one(1234))) == [1]
    r2 = one(pyref.RMONT * pyref.RMONT % p)
    to_m = lambda v: O.fp768_binop("mul", one(v), r2)
    from_m = lambda a: pyref.array_to_ints(O.fp768_binop("mul", a.reshape(1, 24), one(1)))[0]
    assert from_m(O.fp768_binop("mul", to_m(1234), to_m(1234))) == 1522756
    for base, e in [(2, 0), (2, 2), (4, 10), (2, 20), (2, 35)]:
        assert from_m(O.fp768_pow(to_m(base)[0], e)) == base ** e


@pytest.mark.parametrize("n", [1, 2, 8, 64, 512])
def test_fft768_is_the_dft(n):
    p = pyref.FR
    rnd = random.Random(n)
    a = [rnd.randrange(p) for _ in range(n)]
    w = pyref.fr_omega(n) if n > 1 else 1
    exp = pyref.naive_dft(a, w, p) if n <= 64 else pyref.ntt(a, w, p)
    am = pyref.ints_to_array([pyref.mont(x) for x in a])
    wm = np.array(pyref.to_limbs(pyref.mont(w)), dtype=np.uint32)
    for lc in (-1, 0, 1, 3):
        assert [pyref.unmont(x) for x in pyref.array_to_ints(O.fft768(am, wm, lc))] == exp, (n, lc)
    assert [pyref.unmont(x) for x in pyref.array_to_ints(O.naive_dft768(am, wm))] == exp
    assert (O.fft768(O.fft768(am, wm, 2), wm, 2, inverse=True) == am).all()
    ks = np.array([0, n - 1, n // 2], dtype=np.uint64)
    assert (O.dft_points768(am, wm, ks) == O.fft768(am, wm, -1)[ks.astype(np.int64)]).all()


@pytest.mark.parametrize("n", [1, 2, 16, 1024, 1 << 14])
def test_fft32_is_the_dft(n):
    p = pyref.P32
    a = fieldgen.random_u32(n, n, p)
    w = fieldgen.omega32(n)
    exp = pyref.ntt([int(x) for x in a], w, p)
    for lc in (-1, 0, 2, 4):
        assert list(O.fft32(a, w, p, lc)) == exp
    assert (O.fft32(O.fft32(a, w, p, 1), w, p, 1, inverse=True) == a).all()
    if n <= 1024:
        assert list(O.naive_dft32(a, w, p)) == exp


# ---------------------------------------------------------------- oracle vs the reference build
def test_reference_modulus_is_fq():
    """the reference's `_mod` is MNT4-753 Fq (2-adicity 15), not the scalar field"""
    L = _ref("verbatim")
    m = np.empty(24, dtype=np.uint32)
    L.ref_mod(m)
    assert pyref.from_limbs(m) == pyref.FQ
    assert pyref.two_adicity(pyref.FQ) == 15 and pyref.two_adicity(pyref.FR) == 30


def test_verbatim_reference_add_sub_match_and_mul_does_not():
    L = _ref("verbatim")
    O.set_field768("fq")
    A, B = fieldgen.random_elements(3000, 5, pyref.FQ), fieldgen.random_elements(3000, 6, pyref.FQ)
    out = np.empty_like(A)
    for opi, op in ((1, "add"), (2, "sub")):
        L.ref_scalar_binop(opi, out, A, B, A.shape[0])
        assert (out == O.fp768_binop(op, A, B)).all(), op
    L.ref_scalar_binop(0, out, A, B, A.shape[0])
    agree = (out == O.fp768_binop("mul", A, B)).all(axis=1).sum()
    assert agree == 0, "the unpatched reference multiply was expected to be wrong everywhere (F2)"


def test_patched_reference_multiply_matches_oracle():
    L = _ref("patched")
    O.set_field768("fq")
    q = pyref.FQ
    A = np.concatenate([fieldgen.random_elements(5000, 7, q), fieldgen.edge_elements(q)])
    B = np.concatenate([fieldgen.random_elements(5000, 8, q), fieldgen.edge_elements(q)[::-1]])
    out = np.empty_like(A)
    for opi, op in ((0, "mul"), (1, "add"), (2, "sub")):
        L.ref_scalar_binop(opi, out, A, B, A.shape[0])
        assert (out == O.fp768_binop(op, A, B)).all(), op


@pytest.mark.parametrize("logn,log_cpus", [(1, -1), (5, -1), (9, -1), (6, 0), (8, 3), (11, 3), (12, 5), (3, 4)])
def test_patched_reference_fft_templates_match_oracle(logn, log_cpus):
    """reference test/fft_host.h (butterfly statement corrected by sed) over the oracle's field
    == oracle/fft_host_oracle.h, for both fields, serial and parallel entry points"""
    L = _ref("patched")
    n = 1 << logn
    a = fieldgen.random_elements(n, 60 + logn)
    w = fieldgen.omega768(n)
    v = a.copy()
    L.ref_fft_over_oracle768(v, n, w, log_cpus, 0)
    assert (v == O.fft768(a, w, log_cpus)).all()
    b = fieldgen.random_u32(n, 61 + logn, pyref.P32)
    w32 = fieldgen.omega32(n)
    v = b.copy()
    L.ref_fft_over_oracle32(v, n, w32, pyref.P32, log_cpus)
    assert (v == O.fft32(b, w32, pyref.P32, log_cpus)).all()


def test_verbatim_reference_fft_is_not_a_dft():
    """with a correct field, the unpatched templates disagree with the DFT"""
    L = _ref("verbatim")
    n = 64
    a = fieldgen.random_elements(n, 3)
    w = fieldgen.omega768(n)
    v = a.copy()
    L.ref_fft_over_oracle768(v, n, w, -1, 0)
    assert (v != O.naive_dft768(a, w)).any(axis=1).sum() > n // 2


def test_oracle_v_fingerprint():
    """the reference host FFT exactly as test/main.cpp:38-76 runs it (constant 1234 input,
    omega = Scalar(_mod), log_cpus = 10), at 2^12: v[1].im_rep[0] recorded in SURVEY.md section 8c"""
    L = _ref("verbatim")
    n = 1 << 12
    a = np.zeros((n, 24), dtype=np.uint32)
    a[:, 0] = 1234
    m = np.empty(24, dtype=np.uint32)
    L.ref_mod(m)
    L.ref_fft_scalar(a, n, m, 10, 8)
    assert int(a[1, 0]) == 1359985297

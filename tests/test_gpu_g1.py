"""MNT4-753 G1 multi-exponentiation on the GPU (reference multiexp<mnt4753_G1, Scalar>)
against the textbook affine group law in Python big-ints (tests/g1ref.py).  Results are compared as affine
points, so the projective representative does not matter."""
import random

import numpy as np
import pytest

import g1ref
import pyref

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fq_ctx(ctx):
    """the curve's base field Fq is built into the G1 kernels (a compile-time constant object), so the context's own
    768-bit field -- Fr here, the NTT default -- does not matter"""
    yield ctx


def _pack_points(pts):
    arr = np.zeros((len(pts), 3, 24), dtype=np.uint32)
    for i, P in enumerate(pts):
        for c, v in enumerate(g1ref.to_projective_mont(P)):
            arr[i, c] = pyref.to_limbs(v)
    return arr


def _affine(out):
    return g1ref.from_projective_mont(*[pyref.from_limbs(out[c]) for c in range(3)])


def test_g1_is_independent_of_the_context_field(ctx):
    """round 1 required gsn_set_field768(FQ) first (and that switch was device wide); now either setting gives the same point"""
    import gpusnarks_b200 as g
    rng = random.Random(5)
    P = g1ref.random_point(rng)
    a = ctx.g1_multiexp(_pack_points([P]), pyref.ints_to_array([987654321]))
    ctx.set_field768(g.FIELD_FQ)
    try:
        b = ctx.g1_multiexp(_pack_points([P]), pyref.ints_to_array([987654321]))
    finally:
        ctx.set_field768(g.FIELD_FR)
    assert _affine(a) == _affine(b) == g1ref.mul(987654321, P)


def test_small_scalars_and_group_law_cases(fq_ctx):
    rng = random.Random(11)
    P, R = g1ref.random_point(rng), g1ref.random_point(rng)
    negP = (P[0], (g1ref.Q - P[1]) % g1ref.Q)
    cases = [
        ([P], [0]), ([P], [1]), ([P], [2]), ([P], [3]), ([P], [0xFFFFFFFF]), ([P], [1 << 32]), ([P], [(1 << 64) + 5]),
        ([P, R], [1, 1]), ([P, P], [1, 1]),               # doubling inside the reduction
        ([P, negP], [1, 1]), ([P, negP], [7, 7]),         # inverse points: identity
        ([None, P], [5, 9]), ([P, None, R], [2, 3, 4]),   # identity inputs
        ([P, R, P, R], [1, 2, 3, 4]),
    ]
    for pts, ks in cases:
        out = fq_ctx.g1_multiexp(_pack_points(pts), pyref.ints_to_array(ks))
        assert _affine(out) == g1ref.multiexp(pts, ks), (pts, ks)
    # projective inputs with Z != 1: scale (X, Y, Z) by a field element
    lam = rng.randrange(1, g1ref.Q)
    arr = _pack_points([P])
    for c, v in enumerate(g1ref.to_projective_mont(P)):
        arr[0, c] = pyref.to_limbs(v * lam % g1ref.Q)
    assert _affine(fq_ctx.g1_multiexp(arr, pyref.ints_to_array([12345]))) == g1ref.mul(12345, P)


@pytest.mark.parametrize("n,seed", [(1, 1), (5, 2), (40, 3), (129, 4)])
def test_random_full_width_scalars(fq_ctx, n, seed):
    rng = random.Random(seed)
    pts = [g1ref.random_point(rng) for _ in range(n)]
    ks = [rng.randrange(pyref.FR) for _ in range(n)]
    out = fq_ctx.g1_multiexp(_pack_points(pts), pyref.ints_to_array(ks))
    assert _affine(out) == g1ref.multiexp(pts, ks)
    # canonical coordinates
    for c in range(3):
        assert pyref.from_limbs(out[c]) < g1ref.Q


def test_reference_test_shape_linearity(fq_ctx):
    """reference test/main.cpp:135-179 multiplies 2^16..2^20 copies of one point; here 2^12 points (several distinct),
    checked by linearity: multiexp(P, s) + multiexp(P, t) == multiexp(P, s + t)"""
    rng = random.Random(9)
    base = [g1ref.random_point(rng) for _ in range(8)]
    n = 1 << 12
    pts = [base[i % 8] for i in range(n)]
    P = _pack_points(base)[np.arange(n) % 8]
    s = [rng.randrange(1 << 64) for _ in range(n)]
    t = [rng.randrange(1 << 64) for _ in range(n)]
    a = _affine(fq_ctx.g1_multiexp(P, pyref.ints_to_array(s)))
    b = _affine(fq_ctx.g1_multiexp(P, pyref.ints_to_array(t)))
    c = _affine(fq_ctx.g1_multiexp(P, pyref.ints_to_array([x + y for x, y in zip(s, t)])))
    assert g1ref.add(a, b) == c
    # and against the model, folding equal points first: sum_i s_i P_(i mod 8) = sum_j (sum of s over class j) P_j
    folded = [sum(s[j::8]) for j in range(8)]
    assert a == g1ref.multiexp(base, folded)


@pytest.mark.parametrize("n,seed,c", [(8, 21, 2), (9, 22, 3), (33, 23, 5), (100, 24, 0), (300, 25, 7), (513, 26, 0)])
def test_bucket_method_equals_group_law(fq_ctx, n, seed, c):
    """Pippenger (signed windows, buckets, running sums, host Horner) against the affine group law, several window
    widths; equal points, inverse pairs and identities mixed in (they exercise the doubling / identity cases inside the
    bucket sums) and scalars with the top bits set (carry into the extra window)"""
    rng = random.Random(seed)
    pts = [g1ref.random_point(rng) for _ in range(n)]
    pts[1] = pts[0]                                            # same point twice
    pts[2] = (pts[0][0], (g1ref.Q - pts[0][1]) % g1ref.Q)      # its inverse
    if n > 5:
        pts[5] = None                                          # identity input
    ks = [rng.randrange(pyref.FR) for _ in range(n)]
    ks[0] = ks[1] = ks[2] = 3                                  # same bucket: P + P (doubling) then + (-P)
    ks[3] = (1 << 768) - 1                                     # every signed digit carries
    ks[4] = 0
    out = fq_ctx.g1_multiexp(_pack_points(pts), pyref.ints_to_array(ks), method="bucket", window_bits=c)
    assert _affine(out) == g1ref.multiexp(pts, ks)
    for cc in range(3):
        assert pyref.from_limbs(out[cc]) < g1ref.Q


def test_bucket_method_equals_naive_method_2pow12(fq_ctx):
    """both algorithms on 2^12 points with full-width scalars give the same affine point"""
    rng = random.Random(31)
    base = [g1ref.random_point(rng) for _ in range(16)]
    n = 1 << 12
    P = _pack_points(base)[np.arange(n) % 16]
    ks = pyref.ints_to_array([rng.randrange(pyref.FR) for _ in range(n)])
    a = _affine(fq_ctx.g1_multiexp(P, ks, method="naive"))
    b = _affine(fq_ctx.g1_multiexp(P, ks, method="bucket"))
    assert a == b
    folded = [sum(pyref.from_limbs(k) for k in ks[j::16]) for j in range(16)]
    assert a == g1ref.multiexp(base, folded)


def test_fp2_arithmetic(fq_ctx):
    """the reference's fp2 (Fq2 = Fq[u]/(u^2 - 13), cuda/device_field.h:220-294) against Python integers"""
    import fieldgen
    q = g1ref.Q
    n = 500
    a = np.stack([fieldgen.random_elements(n, 61, q), fieldgen.random_elements(n, 62, q)], axis=1)
    b = np.stack([fieldgen.random_elements(n, 63, q), fieldgen.random_elements(n, 64, q)], axis=1)
    a[0] = 0
    b[1] = 0
    a[2, 0] = pyref.to_limbs(q - 1)
    a[2, 1] = pyref.to_limbs(q - 1)
    rinv = pow(pyref.RMONT, -1, q)

    def ints(v):
        return [(pyref.from_limbs(r[0]) * rinv % q, pyref.from_limbs(r[1]) * rinv % q) for r in v]
    A, B = ints(a), ints(b)
    for op in ("mul", "add", "sub"):
        got = ints(fq_ctx.fp2_binop(op, a, b))
        for (x, y), (X, Y), g_ in zip(A, B, got):
            if op == "mul":
                exp = ((x * X + 13 * y * Y) % q, (x * Y + y * X) % q)
            elif op == "add":
                exp = ((x + X) % q, (y + Y) % q)
            else:
                exp = ((x - X) % q, (y - Y) % q)
            assert g_ == exp, op


def test_cpp_multiexp_driver_runs(tmp_path):
    """tests/cpp/test_multiexp_main.cpp = the reference's test_multiexp / test_multiexp_mnt4753_G1 drivers (reference
    test/main.cpp:89-179) through cuda/multi_exp.h: device multiexp == the reference's host loop, for Scalar x Scalar
    and for 24 curve points x full-width scalars (bucket method on the device, double-and-add on the host)"""
    import subprocess
    import test_host_cpp
    rng = random.Random(41)
    pts = [g1ref.random_point(rng) for _ in range(24)]
    path = tmp_path / "points.bin"
    path.write_bytes(_pack_points(pts).tobytes())
    exe = test_host_cpp.build_multiexp_driver(tmp_path)
    out = subprocess.run([exe, str(path), "24", "12"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("DONE") == 2 and "Missmatch" not in out.stdout


def test_heavy_buckets(fq_ctx):
    """300 points with the SAME small scalar all land in one bucket (more than the per-thread limit): the heavy-bucket
    kernel sums them with a block; and scalars below 2^20 leave every higher window empty"""
    rng = random.Random(51)
    base = [g1ref.random_point(rng) for _ in range(5)]
    n = 300
    pts = [base[i % 5] for i in range(n)]
    out = fq_ctx.g1_multiexp(_pack_points(pts), pyref.ints_to_array([5] * n), method="bucket", window_bits=4)
    assert _affine(out) == g1ref.multiexp(base, [5 * 60] * 5)
    ks = [rng.randrange(1 << 20) for _ in range(n)]
    out = fq_ctx.g1_multiexp(_pack_points(pts), pyref.ints_to_array(ks), method="bucket")
    assert _affine(out) == g1ref.multiexp(base, [sum(ks[j::5]) for j in range(5)])


def test_split_buckets(fq_ctx):
    """buckets of more than 4096 points are cut into parts summed by several blocks and combined afterwards (the carry of
    the signed recoding alone sends half of all points into one bucket of the last window): 10 000 and 9 000 points in
    two buckets, projective (Z != 1) and affine inputs mixed, negative digits included"""
    rng = random.Random(52)
    base = [g1ref.random_point(rng) for _ in range(5)]
    n = 19000
    pts = [base[i % 5] for i in range(n)]
    ks = [7 if i < 10000 else (9 << 8) for i in range(n)]        # window_bits 4: digit 7 of window 0; digit -7 of window 2 + carry
    packed = _pack_points(pts)
    out = fq_ctx.g1_multiexp(packed, pyref.ints_to_array(ks), method="bucket", window_bits=4)
    want = g1ref.multiexp(base, [sum(ks[j::5]) for j in range(5)])
    assert _affine(out) == want
    # the host entry point with the automatic window: 5.5 MB of pageable points go through the pinned bounce buffers
    assert _affine(fq_ctx.g1_multiexp(packed, pyref.ints_to_array(ks))) == want
    # the same points in a non-affine projective representation (x l, y l, z l): the general addition path
    p = pyref.FQ
    lam = 0x1234567
    scaled = packed.copy()
    for i in range(5):
        for c in range(3):
            v = pyref.from_limbs(packed[i, c]) * lam % p
            scaled[i::5, c] = np.array(pyref.to_limbs(v), dtype=np.uint32)
    out = fq_ctx.g1_multiexp(scaled, pyref.ints_to_array(ks), method="bucket", window_bits=4)
    assert _affine(out) == want


def test_narrow_windows_many_points(fq_ctx):
    """2-bit windows leave two buckets per window, each with about a quarter of ALL points: far above any per-thread
    limit, so every bucket must go to the block-per-item kernels (a limit that only followed the average bucket size
    would hand 7 500 points to one thread)"""
    rng = random.Random(53)
    base = [g1ref.random_point(rng) for _ in range(5)]
    n = 30000
    pts = [base[i % 5] for i in range(n)]
    ks = [rng.randrange(1 << 16) for _ in range(n)]
    out = fq_ctx.g1_multiexp(_pack_points(pts), pyref.ints_to_array(ks), method="bucket", window_bits=2)
    assert _affine(out) == g1ref.multiexp(base, [sum(ks[j::5]) for j in range(5)])


def test_multiexp_sharded_over_the_visible_devices(fq_ctx):
    """gsn_g1_multiexp_multi_host: one slice of the points per device (1, 2, ... 8 of them), partial sums added on the
    host; an uneven split and a device list with one entry included.  Same point as the single-context call."""
    import gpusnarks_b200 as g
    from gpusnarks_b200.ntt import g1_multiexp_multi
    rng = random.Random(61)
    base = [g1ref.random_point(rng) for _ in range(7)]
    n = 1003
    pts = [base[i % 7] for i in range(n)]
    ks = [rng.randrange(pyref.FR) for _ in range(n)]
    packed, scal = _pack_points(pts), pyref.ints_to_array(ks)
    want = g1ref.multiexp(base, [sum(ks[j::7]) for j in range(7)])
    assert _affine(fq_ctx.g1_multiexp(packed, scal)) == want
    assert _affine(g1_multiexp_multi(packed, scal)) == want                       # every visible device
    assert _affine(g1_multiexp_multi(packed, scal, devices=[0])) == want
    if g.device_count() >= 2:
        assert _affine(g1_multiexp_multi(packed, scal, devices=[1, 0])) == want
    assert _affine(g1_multiexp_multi(packed[:3], scal[:3], devices=[0] * 1)) == g1ref.multiexp(base[:3], ks[:3])

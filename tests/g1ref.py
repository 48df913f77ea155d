"""Python big-int model of the MNT4-753 G1 group law used by the reference's multiexp
(reference cuda/device_field.h:296-437: projective coordinates, coefficient a = 2, `operator+` = add-1998-cmo-2,
`dbl` = dbl-2007-bl, `operator*` = MSB-first double-and-add).  The reference's own routines cannot serve as the
oracle: `operator+` has no identity / doubling cases, so `zero() + P` is (0,0,0) and every scalar multiple comes
out as zero; unary minus on Scalar returns 0.  This model is the textbook affine group law
(independent of the projective formulas the CUDA code uses) plus helpers to build points.

The curve constant b never enters the addition formulas, so tests build points on y^2 = x^3 + 2x + B over Fq with a
synthetic B (the real MNT4-753 b is not in the reference tree); every point of one test lies on the same curve."""
import random

import pyref

Q = pyref.FQ
A = 2
B = 0x1337  # synthetic: the group law only needs all points on one curve (discriminant != 0 checked below)
assert (4 * A ** 3 + 27 * B * B) % Q != 0


def sqrt_mod(a, p=Q):
    """Tonelli-Shanks; returns None if a is a non-residue"""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    s, t = 0, p - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, x, b = s, pow(z, t, p), pow(a, (t + 1) // 2, p), pow(a, t, p)
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % p
            i += 1
        e = pow(c, 1 << (m - i - 1), p)
        m, c, x, b = i, e * e % p, x * e % p, b * e % p * e % p
    return x


def random_point(rng):
    while True:
        x = rng.randrange(Q)
        y = sqrt_mod((x * x * x + A * x + B) % Q)
        if y is not None:
            return (x, y if rng.random() < 0.5 else (Q - y) % Q)


def add(P, R):
    """affine group law; None is the identity"""
    if P is None:
        return R
    if R is None:
        return P
    (x1, y1), (x2, y2) = P, R
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1 + A) * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


def mul(k, P):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = add(acc, acc)
        if bit == "1":
            acc = add(acc, P)
    return acc


def multiexp(points, scalars):
    acc = None
    for P, k in zip(points, scalars):
        acc = add(acc, mul(k, P))
    return acc


def to_projective_mont(P):
    """affine point (or None) -> three Montgomery-form integers (X, Y, Z); identity = (0, R mod q, 0)"""
    if P is None:
        return (0, pyref.RMONT % Q, 0)
    return (P[0] * pyref.RMONT % Q, P[1] * pyref.RMONT % Q, pyref.RMONT % Q)


def from_projective_mont(X, Y, Z):
    """three Montgomery-form integers -> affine point or None"""
    rinv = pow(pyref.RMONT, -1, Q)
    x, y, z = X * rinv % Q, Y * rinv % Q, Z * rinv % Q
    if z == 0:
        return None
    zi = pow(z, -1, Q)
    return (x * zi % Q, y * zi % Q)


if __name__ == "__main__":
    rng = random.Random(1)
    P, R = random_point(rng), random_point(rng)
    assert add(add(P, R), P) == add(P, add(R, P))
    assert mul(5, P) == add(add(add(add(P, P), P), P), P)
    assert mul(7, add(P, R)) == add(mul(7, P), mul(7, R))
    print("g1ref ok")

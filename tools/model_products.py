#!/usr/bin/env python3
"""Executable models of the two 768-bit device products of gpusnarks_b200/csrc/fp768.cuh, limb for limb:

  cios_mul   the CIOS Montgomery product with two interleaved (even / odd aligned) accumulators whose roles swap at
             every one-limb Montgomery shift (`cios_step`, `mont_mul_lazy_w`)
  shoup_mul  the fixed-operand product t = x*w - q*p from three truncated half products (`row_mac`, `merge_evod`,
             `mul_lo768`, `shoup_mul_lazy`), q = limbs 24..47 of x*w'' summed from limb position 22 upwards

Both use 32-bit limbs with explicit carries exactly where the PTX has them (mad.lo.cc / madc.hi.cc chains, the addc
into the next limb, the dropped carries at the truncation boundary), so they prove the carry and truncation logic
rather than just the algebra.  `python tools/model_products.py` checks them against Python big-ints on random and
edge operands; tests/test_golden_and_model.py runs the same check."""
import random

M32 = (1 << 32) - 1
NL = 24
R = 1 << (32 * NL)


def limbs(x, n=NL):
    return [(x >> (32 * i)) & M32 for i in range(n)]


def val(l):
    return sum(v << (32 * i) for i, v in enumerate(l))


# ------------------------------------------------------------------ CIOS with even/odd accumulators
class _CC:
    c = 0


def _add_cc(a, b):
    s = a + b
    _CC.c = s >> 32
    return s & M32


def _addc_cc(a, b):
    s = a + b + _CC.c
    _CC.c = s >> 32
    return s & M32


def _addc(a, b):
    return (a + b + _CC.c) & M32


def _mad_wide(lo, hi, a, b, carry_in):
    t = a * b
    s = (t & M32) + lo + (_CC.c if carry_in else 0)
    c = s >> 32
    lo2 = s & M32
    s = (t >> 32) + hi + c
    _CC.c = s >> 32
    return lo2, s & M32


def _cios_step(first, ev, od, a, bi, P, np0):
    if first:
        for j in range(0, NL, 2):
            t = a[j + 1] * bi
            od[j], od[j + 1] = t & M32, t >> 32
        for j in range(0, NL, 2):
            t = a[j] * bi
            ev[j], ev[j + 1] = t & M32, t >> 32
    else:
        ev[0] = _add_cc(ev[0], od[1])
        for j in range(0, NL - 2, 2):
            od[j], od[j + 1] = _mad_wide(od[j + 2], od[j + 3], a[j + 1], bi, True)
        od[NL - 2], od[NL - 1] = _mad_wide(0, 0, a[NL - 1], bi, True)
        ev[0], ev[1] = _mad_wide(ev[0], ev[1], a[0], bi, False)
        for j in range(2, NL, 2):
            ev[j], ev[j + 1] = _mad_wide(ev[j], ev[j + 1], a[j], bi, True)
        od[NL - 1] = _addc(od[NL - 1], 0)
    m = (ev[0] * np0) & M32
    od[0], od[1] = _mad_wide(od[0], od[1], P[1], m, False)
    for j in range(2, NL, 2):
        od[j], od[j + 1] = _mad_wide(od[j], od[j + 1], P[j + 1], m, True)
    ev[0], ev[1] = _mad_wide(ev[0], ev[1], P[0], m, False)
    for j in range(2, NL, 2):
        ev[j], ev[j + 1] = _mad_wide(ev[j], ev[j + 1], P[j], m, True)
    od[NL - 1] = _addc(od[NL - 1], 0)


def cios_mul(x, y, p):
    """x * y * 2^-768 mod p, lazy (< 2p), via the device's even/odd CIOS"""
    P, a, b = limbs(p), limbs(x), limbs(y)
    np0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    ev, od = [0] * NL, [0] * NL
    _cios_step(True, ev, od, a, b[0], P, np0)
    _cios_step(False, od, ev, a, b[1], P, np0)
    for i in range(2, NL, 2):
        _cios_step(False, ev, od, a, b[i], P, np0)
        _cios_step(False, od, ev, a, b[i + 1], P, np0)
    r = [0] * NL
    r[0] = _add_cc(od[1], ev[0])
    for k in range(1, NL - 1):
        r[k] = _addc_cc(od[k + 1], ev[k])
    r[NL - 1] = _addc(ev[NL - 1], 0)
    return val(r)


# ------------------------------------------------------------------ fixed-operand product
class _Acc:
    """EV/OD accumulators over limb positions [base, ...): EV[k] holds limb base+k, OD[k] limb base+k+1"""

    def __init__(self, base, size):
        self.base = base
        self.ev = [0] * (size + 2)
        self.od = [0] * (size + 2)

    def row(self, a, b, i, jlo, jhi, lo_pos=None, top=None, fresh_top=False):
        """fresh_top: the device code drops the carry propagation after a chain whose last link is the product a[NL-1]*b
        (limb position i + NL - 1): that (even/odd) register pair is touched by no earlier product and holds at most a
        propagated carry bit in its low limb, so lo + product + carry-in < 2^64 and the link cannot carry out."""
        for parity in (0, 1):
            arr = self.ev if parity == 0 else self.od
            carry, last, last_j = 0, None, None
            for j in range(jlo, jhi):
                pos = i + j
                if pos % 2 != parity:
                    continue
                k = pos - self.base - parity
                prod = a[j] * b
                s = arr[k] + (prod & M32) + carry
                arr[k], carry = s & M32, s >> 32
                if lo_pos is not None and pos == lo_pos:
                    carry, last = 0, None
                    continue
                s = arr[k + 1] + (prod >> 32) + carry
                arr[k + 1], carry = s & M32, s >> 32
                last, last_j = k + 1, j
            if fresh_top and last is not None and last_j == NL - 1:
                assert carry == 0, "the top link of a row carried out"
                continue
            if last is not None and carry and (top is None or self.base + last + 1 + parity < top):
                s = arr[last + 1] + carry
                assert s >> 32 == 0
                arr[last + 1] = s & M32

    def merge(self, n):
        out, carry = [], 0
        for k in range(n):
            s = self.ev[k] + (self.od[k - 1] if k else 0) + carry
            out.append(s & M32)
            carry = s >> 32
        return out


def mul_lo768(a, b):
    acc = _Acc(0, NL + 2)
    for i in range(NL):
        acc.row(a, b[i], i, 0, NL - i, lo_pos=NL - 1, top=NL)
    return acc.merge(NL)


def shoup_mul(x, w, p):
    """x * w mod p, lazy (< 2p), via the device's truncated half products; returns (t, q)"""
    X, W, W2, P = limbs(x), limbs(w), limbs((w * R) // p), limbs(p)
    acc = _Acc(22, NL + 4)
    for i in range(NL):
        acc.row(W2, X[i], i, max(0, 22 - i), NL, fresh_top=True)
    q = acc.merge(NL + 2)[2:]
    t = (val(mul_lo768(W, X)) - val(mul_lo768(q, P))) % R
    if t >= 2 * p:
        t -= 2 * p
    return t, val(q)


def shoup_mul_3p(x, w, p):
    """the same product without the final conditional subtraction (`shoup_mul_3p`): t = x*w - q*p for ANY x < 2^768"""
    X, W, W2, P = limbs(x), limbs(w), limbs((w * R) // p), limbs(p)
    acc = _Acc(22, NL + 4)
    for i in range(NL):
        acc.row(W2, X[i], i, max(0, 22 - i), NL, fresh_top=True)
    q = acc.merge(NL + 2)[2:]
    return (val(mul_lo768(W, X)) - val(mul_lo768(q, P))) % R, val(q)


def reduce_small(v, p):
    """`reduce_small`: v < 1024p -> [0, p) with a quotient estimate from the top limb (q*p by a 24-step mad chain,
    then two conditional subtractions)"""
    qmagic = (1 << 32) // ((p >> 736) + 1)
    q = ((v >> 736) * qmagic) >> 32
    P, carry, qp = limbs(p), 0, []
    for i in range(NL):
        t = P[i] * q + carry
        qp.append(t & M32)
        carry = t >> 32
    assert carry == 0
    r = (v - val(qp)) % R
    assert r < 3 * p, "quotient estimate too small"
    if r >= 2 * p:
        r -= 2 * p
    if r >= p:
        r -= p
    return r


def lazy_pass_bounds(p, stages=10, unit_stages=2):
    """value bounds inside a pass with wide lazy ranges (fp768.cuh): returns the bound after the last stage in units of
    p.  unit_stages = 2: warp-owned kernel (stage 1 and half of stage 2 skip the product); 4: CTA-wide kernel
    (stages 1..4 have unit-twiddle butterflies, which subtract from 3p * 2^(s-1))."""
    b = 3                      # canonical input (< p) or pre-twiddle product (< 3p)
    for s in range(1, stages + 1):
        if s <= unit_stages:   # unit butterflies take t as it is: t < b, subtracted from k p with k = 3 * 2^(s-1)
            k = 3 << (s - 1)
            assert b <= k, (s, b, k)
            b = b + k
        else:                  # products give t < 3p
            b = b + 3
    return b


def shoup_constant_from_montgomery(w_mont, p):
    """w'' = floor(w * 2^768 / p) computed the device's way: lo768(w_mont * (-p^-1 mod 2^768))"""
    nprime = (-pow(p, -1, R)) % R
    return val(mul_lo768(limbs(w_mont), limbs(nprime)))


def self_check(p, cases=400, seed=1):
    rnd = random.Random(seed)
    rinv = pow(R, -1, p)
    edge_w = [0, 1, 2, p - 1, p // 2, (p + 1) // 2]
    edge_x = [0, 1, p - 1, p, p + 1, 2 * p - 1, 2 * p - 2]
    worst = 0
    for it in range(cases):
        w = edge_w[it % len(edge_w)] if it % 5 == 0 else rnd.randrange(p)
        x = edge_x[it % len(edge_x)] if it % 7 == 0 else rnd.randrange(2 * p)
        t, q = shoup_mul(x, w, p)
        assert t < 2 * p and t % p == x * w % p
        assert 0 <= (x * w) // p - q <= 2
        worst = max(worst, (x * w) // p - q)
        c = cios_mul(w, x, p)
        assert c < 2 * p and c % p == x * w * rinv % p
        wm = w * R % p
        assert shoup_constant_from_montgomery(wm, p) == (w * R) // p
        # wide lazy ranges: any x below 2^768 (not only below 2p) gives t in [0, 3p) and t = x*w mod p
        xs = [R - 1, R - p, 36 * p, 64 * p - 1, rnd.randrange(R), rnd.randrange(36 * p), R - (1 << 32), (R - 1) ^ ((1 << 384) - 1)]
        xl = xs[it % len(xs)]
        t3, q3 = shoup_mul_3p(xl, w, p)
        assert t3 < 3 * p and t3 % p == xl * w % p and 0 <= (xl * w) // p - q3 <= 2
        worst = max(worst, (xl * w) // p - q3)
        # reduce_small over its whole domain [0, 1024p)
        vs = [0, p - 1, p, 2 * p, 3 * p - 1, 36 * p, 66 * p - 1, 1024 * p - 1, rnd.randrange(1024 * p), rnd.randrange(1024) * p,
              rnd.randrange(1, 1024) * p - 1]
        vv = vs[it % len(vs)]
        assert reduce_small(vv, p) == vv % p
    assert lazy_pass_bounds(p, 10, 2) == 36 and lazy_pass_bounds(p, 10, 4) == 66 and 1024 * p < R
    return worst


if __name__ == "__main__":
    FR = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888458477323173057491593855069696241854796396165721416325350064441470418137846398469611935719059908164220784476160001
    print("model_products: ok, worst quotient error", self_check(FR))

#!/usr/bin/env python3
"""Executable model of the multi-pass NTT index math used by csrc/ntt768.cu / ntt32.cu.

The CUDA pass kernel is a transcription of `run_pass` below (same names, same formulas);
this model runs it over a small prime field and checks the result against the DFT
definition, for every geometry the host planner can produce (batch, digit splits, inner
stride log_r, partial transforms for the multi-GPU column step).  Run: python tools/model_passes.py
"""
import itertools
import random

P = 2013265921
GEN = 31


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def plan_digits(logn, max_log=10):
    """even split into the fewest passes of at most max_log stages"""
    if logn == 0:
        return [0]
    npass = -(-logn // max_log)
    base, extra = divmod(logn, npass)
    return [base + (1 if i < extra else 0) for i in range(npass)]


def make_tables(w_n, logn, digits, n_inv=None):
    """local table + one pre-twiddle table per pass boundary.
    wloc[k] = w_T^k (k < T/2) with T = 2^max(digits);  TW_q[idx] = w_{N_q}^(k * rest), idx = (k, rest)"""
    n = 1 << logn
    lmax = max(digits)
    w_t = pow(w_n, n >> lmax, P) if logn >= lmax else None
    wloc = [pow(w_t, k, P) for k in range(max(1, (1 << lmax) // 2))]
    pre = [None] * len(digits)
    for q in range(1, len(digits)):
        logN = sum(digits[q - 1:])          # sub-problem size at level q-1
        rest_bits = logN - digits[q - 1]
        w_N = pow(w_n, n >> logN, P)
        tab = []
        for idx in range(1 << logN):
            k, rest = idx >> rest_bits, idx & ((1 << rest_bits) - 1)
            tab.append(pow(w_N, k * rest, P))
        pre[q] = tab
    if n_inv is not None:
        if len(digits) > 1:
            pre[1] = [x * n_inv % P for x in pre[1]]
        else:
            pre[0] = [n_inv]               # scalar pre-multiply (mask 0)
    return wloc, lmax, pre


def run_pass(src, dst, q, digits, log_r, log_tile, wloc, lmax, pre_tab, final_natural):
    """one pass = every tile of 2^log_tile elements; transforms digit q (0-based)."""
    total = len(src)
    P_ = len(digits)
    lq = digits[q]
    L = 1 << lq
    log_s = log_r + sum(digits[q + 1:])     # stride of digit q
    T = 1 << log_tile
    subs_per_tile = T >> lq
    n_tiles = total >> log_tile
    logn_t = sum(digits)
    for tile in range(n_tiles):
        smem = [None] * T
        # ---- load (bit-reversed placement) + optional pre-twiddle
        for e in range(T):
            slot, j = e >> lq, e & (L - 1)
            t = tile * subs_per_tile + slot
            o, rlow = t >> log_s, t & ((1 << log_s) - 1)
            g = (((o << lq) | j) << log_s) | rlow
            x = src[g]
            if pre_tab is not None:
                if len(pre_tab) == 1:
                    x = x * pre_tab[0] % P
                else:
                    x = x * pre_tab[(g >> log_r) & (len(pre_tab) - 1)] % P
            smem[(slot << lq) | bitrev(j, lq)] = x
        # ---- stages
        for s in range(1, lq + 1):
            m = 1 << (s - 1)
            for b in range(T // 2):
                jj = b & (m - 1)
                lo = ((b >> (s - 1)) << s) | jj
                hi = lo + m
                w = wloc[(jj << (lq - s)) << (lmax - lq)]
                tval = smem[hi] * w % P
                u = smem[lo]
                smem[lo] = (u + tval) % P
                smem[hi] = (u - tval) % P
        # ---- store
        for e in range(T):
            slot, k = e >> lq, e & (L - 1)
            t = tile * subs_per_tile + slot
            o, rlow = t >> log_s, t & ((1 << log_s) - 1)
            if not final_natural:
                g = (((o << lq) | k) << log_s) | rlow
            else:
                # o = (batch, k_0 .. k_{P-2}) most significant first; rlow = r (log_s == log_r)
                assert q == P_ - 1 and log_s == log_r
                inner_bits = logn_t - lq
                batch, rest = o >> inner_bits, o & ((1 << inner_bits) - 1)
                out = 0
                shift = 0
                pos = inner_bits
                for qq in range(P_ - 1):          # k_0 is least significant in the output
                    pos -= digits[qq]
                    kq = (rest >> pos) & ((1 << digits[qq]) - 1)
                    out |= kq << shift
                    shift += digits[qq]
                out |= k << shift
                g = (((batch << logn_t) | out) << log_r) | rlow
            dst[g] = smem[e]


def transform(data, w_n, logn, log_r, batch, max_log, log_tile_max, n_inv=None):
    digits = plan_digits(logn, max_log)
    wloc, lmax, pre = make_tables(w_n, logn, digits, n_inv)
    total = len(data)
    log_tile = min(log_tile_max, (total - 1).bit_length())
    cur = list(data)
    for q in range(len(digits)):
        nxt = [None] * total
        run_pass(cur, nxt, q, digits, log_r, max(log_tile, digits[q]), wloc, lmax, pre[q], q == len(digits) - 1)
        cur = nxt
    return cur


def check(logn, log_r, batch, max_log, log_tile):
    n = 1 << logn
    R = 1 << log_r
    w = pow(GEN, (P - 1) // n, P) if n > 1 else 1
    rnd = random.Random(logn * 1000 + log_r * 100 + batch)
    data = [rnd.randrange(P) for _ in range(batch * n * R)]
    got = transform(data, w, logn, log_r, batch, max_log, log_tile)
    n_inv = pow(n, -1, P)
    winv = pow(w, -1, P)
    back = transform(got, winv, logn, log_r, batch, max_log, log_tile, n_inv=n_inv)
    assert back == data, ("roundtrip", logn, log_r, batch, max_log)
    for b in range(batch):
        for r in range(R):
            a = [data[((b << logn) | i) << log_r | r] for i in range(n)]
            exp = [sum(a[j] * pow(w, i * j, P) for j in range(n)) % P for i in range(n)]
            out = [got[((b << logn) | i) << log_r | r] for i in range(n)]
            assert out == exp, (logn, log_r, batch, max_log)


if __name__ == "__main__":
    cases = 0
    for logn, log_r, batch, max_log in itertools.product(range(1, 9), (0, 1, 2), (1, 3), (2, 3, 4)):
        log_tile = max_log + 1 if logn + log_r > max_log else max_log
        if (batch << (logn + log_r)) < (1 << log_tile):
            continue
        if batch == 3 and (3 << (logn + log_r)) % (1 << log_tile):
            continue
        check(logn, log_r, batch, max_log, log_tile)
        cases += 1
    print("model_passes: ok,", cases, "geometries")

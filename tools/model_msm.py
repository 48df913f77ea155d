#!/usr/bin/env python3
"""Executable model of the bucket-method bookkeeping of gpusnarks_b200/csrc/g1.cuh over the additive group of integers
(a "point" is an integer, k * P is a product), so every identity can be checked exactly on the CPU:

  digits        g1_digits_kernel: signed c-bit windows with carry, |digit| <= 2^(c-1), keys = w * bs + |digit|
  heavy items   g1_bucket_kernel / g1_heavy_bucket_kernel / g1_heavy_combine_kernel: a bucket above `limit` points
                becomes work items, above G1_SPLIT_POINTS it is cut into `parts` slices [lo + s k / parts, lo + s (k+1) / parts)
  window sums   g1_window_reduce_kernel: P blocks x T threads per window, thread t of block j takes m buckets above
                b0 = j nb/P + t m and contributes sum (b - b0) B_b + b0 * (chunk total)
  Horner        acc = 2^c acc + S_w over the windows, most significant first (host side of g1_multiexp_pippenger)

`python tools/model_msm.py` runs the checks; tests/test_golden_and_model.py calls check()."""
import random

SPLIT_POINTS, MAX_PARTS = 4096, 256


def digits(k, c, windows):
    """(window, |digit|, negative) per window, as the device recodes a raw 768-bit scalar"""
    out, carry = [], 0
    for w in range(windows):
        v = ((k >> (w * c)) & ((1 << c) - 1)) + carry
        neg, carry = 0, 0
        if v > (1 << (c - 1)):
            v, neg, carry = (1 << c) - v, 1, 1
        out.append((w, v, neg))
    assert carry == 0, "the extra window absorbs the last carry"
    return out


def heavy_limit(n, c):
    return min(1024, max(128, 4 * (n >> (c - 1))))


def bucket_sums(points, scalars, c):
    n = len(points)
    windows, nb = (768 + 1 + c - 1) // c, 1 << (c - 1)
    bs = nb + 1
    pairs = []
    for i, k in enumerate(scalars):
        for w, v, neg in digits(k, c, windows):
            pairs.append((w * bs + v, 2 * i + neg))
    pairs.sort(key=lambda p: p[0])           # the radix sort (stable; only the key order matters)
    keys = [p[0] for p in pairs]
    buckets = [0] * (windows * bs)
    limit = heavy_limit(n, c)
    import bisect
    stats = {"heavy": 0, "split": 0}
    for key in range(windows * bs):
        if key % bs == 0:
            continue
        lo, hi = bisect.bisect_left(keys, key), bisect.bisect_left(keys, key + 1)
        size = hi - lo

        def signed(v):
            return -points[v >> 1] if v & 1 else points[v >> 1]
        if size > limit:
            stats["heavy"] += 1
            parts = 1
            if size > SPLIT_POINTS:
                parts = min(MAX_PARTS, (size + SPLIT_POINTS - 1) // SPLIT_POINTS)
                stats["split"] += 1
            covered, total = 0, 0
            for k in range(parts):
                a, b = lo + size * k // parts, lo + size * (k + 1) // parts
                covered += b - a
                total += sum(signed(pairs[i][1]) for i in range(a, b))
            assert covered == size, "the parts tile the bucket exactly"
            buckets[key] = total
        else:
            buckets[key] = sum(signed(pairs[i][1]) for i in range(lo, hi))
    return buckets, windows, bs, stats


def window_sum(B, c, P, T):
    """S = sum_b b * B[b] the way the P x T threads of one window compute it"""
    nb = (1 << (c - 1)) // P
    total = 0
    for j in range(P):
        base = j * nb
        t_cnt = min(nb, T)
        m = nb // t_cnt
        for t in range(t_cnt):
            b0 = base + t * m
            run = s = 0
            for b in range(b0 + m, b0, -1):
                run += B[b]
                s += run
            total += s + b0 * run
    return total


def multiexp(points, scalars, c, P=1, T=256):
    buckets, windows, bs, stats = bucket_sums(points, scalars, c)
    acc = 0
    for w in range(windows - 1, -1, -1):
        acc = (acc << c) + window_sum(buckets[w * bs:(w + 1) * bs], c, P, T)
    return acc, stats


def check():
    rng = random.Random(5)
    for c, n, P, T in [(2, 300, 1, 256), (4, 500, 1, 4), (7, 800, 2, 8), (9, 1500, 4, 16), (13, 200, 2, 256)]:
        pts = [rng.randrange(1, 1 << 40) for _ in range(n)]
        ks = [rng.randrange(1 << 768) if i % 3 else rng.randrange(1 << 20) for i in range(n)]
        ks[0], ks[1] = (1 << 768) - 1, 0
        for k in ks[:20]:
            assert sum((-v if neg else v) << (w * c) for w, v, neg in digits(k, c, (768 + 1 + c - 1) // c)) == k
        got, _ = multiexp(pts, ks, c, P, T)
        assert got == sum(p * k for p, k in zip(pts, ks)), (c, n)
    # heavy and split buckets: many points share a scalar
    pts = [rng.randrange(1, 1 << 40) for _ in range(12000)]
    ks = [7] * 9000 + [(9 << 8)] * 3000
    got, stats = multiexp(pts, ks, 4)
    assert got == sum(p * k for p, k in zip(pts, ks)) and stats["split"] >= 1 and stats["heavy"] >= 2, stats
    assert heavy_limit(1 << 20, 16) == 128 and heavy_limit(1 << 22, 16) == 512 and heavy_limit(30000, 2) == 1024
    return True


if __name__ == "__main__":
    check()
    print("model_msm: ok")

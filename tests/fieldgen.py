"""Seeded synthetic field elements: uniform in [0, p), Montgomery form
is implied (any canonical limb pattern is the Montgomery form of some element)."""
import numpy as np

import pyref

NL = 24


def _less_than(a, p_limbs):
    """a: (m, 24) uint32; returns bool mask a < p (lexicographic from the top limb)"""
    lt = np.zeros(a.shape[0], dtype=bool)
    decided = np.zeros(a.shape[0], dtype=bool)
    for i in range(NL - 1, -1, -1):
        pi = np.uint32(p_limbs[i])
        lt |= (~decided) & (a[:, i] < pi)
        decided |= a[:, i] != pi
    return lt


def random_elements(count, seed, p=pyref.FR):
    """count canonical elements of Z/p, (count, 24) uint32, rejection sampled from bits(p)-bit integers"""
    rng = np.random.Generator(np.random.PCG64(seed))
    p_limbs = pyref.to_limbs(p)
    top_bits = p.bit_length() - 32 * (NL - 1)
    out = np.empty((count, NL), dtype=np.uint32)
    filled = 0
    while filled < count:
        m = max(16, int((count - filled) * 1.3))
        cand = rng.integers(0, 1 << 32, size=(m, NL), dtype=np.uint64).astype(np.uint32)
        cand[:, NL - 1] &= np.uint32((1 << top_bits) - 1)
        cand = cand[_less_than(cand, p_limbs)]
        take = min(count - filled, cand.shape[0])
        out[filled:filled + take] = cand[:take]
        filled += take
    return out


def edge_elements(p=pyref.FR):
    vals = [0, 1, 2, p - 1, p - 2, pyref.RMONT % p, (pyref.RMONT * pyref.RMONT) % p, (p - 1) // 2, (p + 1) // 2,
            (1 << 752), (1 << 32) - 1, (1 << 64), p - (1 << 32)]
    return pyref.ints_to_array(vals)


def omega768(n, p=pyref.FR, gen=17):
    """primitive n-th root of unity in Montgomery form, (24,) uint32"""
    w = pyref.root_of_unity(p, gen, n) if n > 1 else 1
    return np.array(pyref.to_limbs(w * pyref.RMONT % p), dtype=np.uint32)


def random_u32(count, seed, mod):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, mod, size=count, dtype=np.uint64).astype(np.uint32)


def omega32(n, mod=pyref.P32, gen=pyref.P32_GEN):
    return pyref.root_of_unity(mod, gen, n) if n > 1 else 1

"""Python big-int model of the two fields and of the DFT definition.

Third, independent judge (after oracle/ in C++ and the CUDA path): everything here is
plain Python integer arithmetic derived from the modulus alone.
"""
import numpy as np

R_DEC = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888458477323173057491593855069696241854796396165721416325350064441470418137846398469611935719059908164220784476160001
Q_LIMBS = [610172929, 1586521054, 752685471, 3818738770, 2596546032, 1669861489, 1987204260, 1750781161,
           3411246648, 3087994277, 4061660573, 2971133814, 2707093405, 2580620505, 3902860685, 134068517,
           1821890675, 1589111033, 1536143341, 3086587728, 4007841197, 270700578, 764593169, 115910]
NL = 24
RMONT = 1 << (32 * NL)
FR = R_DEC
FQ = sum(v << (32 * i) for i, v in enumerate(Q_LIMBS))
P32 = 2013265921
P32_GEN = 31


def to_limbs(x, n=NL):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def from_limbs(l):
    return sum(int(v) << (32 * i) for i, v in enumerate(l))


def ints_to_array(xs):
    """list of ints -> (len, 24) uint32 array, little-endian limbs (AoS, the reference layout)."""
    out = np.zeros((len(xs), NL), dtype=np.uint32)
    for i, x in enumerate(xs):
        out[i] = to_limbs(x)
    return out


def array_to_ints(a):
    a = np.asarray(a, dtype=np.uint32).reshape(-1, NL)
    return [from_limbs(row) for row in a]


def two_adicity(p):
    s, t = 0, p - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    return s


def root_of_unity(p, gen, n):
    """primitive n-th root of unity (n a power of two dividing p-1)"""
    assert (p - 1) % n == 0
    w = pow(gen, (p - 1) // n, p)
    assert n == 1 or pow(w, n // 2, p) == p - 1
    return w


def fr_omega(n):
    return root_of_unity(FR, 17, n)


def fq_omega(n):
    return root_of_unity(FQ, 13, n)


def mont(x, p=FR):
    return x * RMONT % p


def unmont(x, p=FR):
    return x * pow(RMONT, -1, p) % p


def naive_dft(a, w, p):
    n = len(a)
    return [sum(a[j] * pow(w, i * j, p) for j in range(n)) % p for i in range(n)]


def ntt(a, w, p):
    """recursive radix-2, natural order in/out; O(n log n) big-int reference"""
    n = len(a)
    if n == 1:
        return list(a)
    even = ntt(a[0::2], w * w % p, p)
    odd = ntt(a[1::2], w * w % p, p)
    out = [0] * n
    x = 1
    for i in range(n // 2):
        t = x * odd[i] % p
        out[i] = (even[i] + t) % p
        out[i + n // 2] = (even[i] - t) % p
        x = x * w % p
    return out

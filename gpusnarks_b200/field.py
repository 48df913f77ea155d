"""Host-side constants and big-int helpers for the two fields (product code, no oracle).

Used by the four-step driver and bench.py to derive roots of unity (omega^(n2), omega^(n1))
from the caller's omega.  Elements of the 768-bit field travel as 24 little-endian uint32
limbs in Montgomery form (R = 2^768), exactly the reference's `fields::Scalar::im_rep`
(reference cuda/device_field.h:75)."""
import numpy as np

NL = 24
RMONT = 1 << (32 * NL)
# MNT4-753 scalar field (the reference quotes it in a comment, cuda/device_field.h:445-446)
FR = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888458477323173057491593855069696241854796396165721416325350064441470418137846398469611935719059908164220784476160001
FR_GENERATOR = 17
FR_TWO_ADICITY = 30
# MNT4-753 base field: the reference's literal `_mod` limbs (cuda/device_field.h:62-65)
FQ = 41898490967918953402344214791240637128170709919953949071783502921025352812571106773058893763790338921418070971888253786114353726529584385201591605722013126468931404347949840543007986327743462853720628051692141265303114721689601
FQ_GENERATOR = 13
FQ_TWO_ADICITY = 15
P32 = 2013265921
P32_GENERATOR = 31


def to_limbs(x):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(NL)], dtype=np.uint32)


def from_limbs(l):
    return sum(int(v) << (32 * i) for i, v in enumerate(np.asarray(l).reshape(-1)))


def to_mont768(x, p=FR):
    """integer x -> Montgomery-form limbs of x mod p"""
    return to_limbs(x % p * RMONT % p)


def mont_pow(w_limbs, e, p=FR):
    """(w^e) for w given and returned in Montgomery form"""
    w = from_limbs(w_limbs) * pow(RMONT, -1, p) % p
    return to_limbs(pow(w, e, p) * RMONT % p)


def mont_inverse(w_limbs, p=FR):
    """multiplicative inverse, Montgomery form in and out"""
    w = from_limbs(w_limbs) * pow(RMONT, -1, p) % p
    return to_limbs(pow(w, -1, p) * RMONT % p)


def root_of_unity768(n, p=FR, gen=FR_GENERATOR):
    """primitive n-th root of unity (n a power of two), Montgomery form limbs"""
    assert n >= 1 and n & (n - 1) == 0 and (p - 1) % n == 0
    return to_limbs(pow(gen, (p - 1) // n, p) * RMONT % p)


def root_of_unity32(n, p=P32, gen=P32_GENERATOR):
    assert n >= 1 and n & (n - 1) == 0 and (p - 1) % n == 0
    return pow(gen, (p - 1) // n, p)

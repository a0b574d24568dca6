"""
Self-check helpers for results that come back over the C ABI: a few dozen lines of big-int BN254 arithmetic (one scalar
multiplication of the generator, Jacobian -> affine) so that `bench.py` and the tools can verify a timed MSM result against
the O(n) checksum  sum_i s_i [z_i]G = [sum_i s_i z_i mod r] G  (h2b_msm_checksum_dev) without touching `oracle/`.
Not on the compute path: the MSM / NTT themselves only ever run in libh2b200.so.

Constants: halo2curves 0.3.x src/bn256/{fq,fr,curve}.rs [UP] ("Verified constants").
"""
from __future__ import annotations

FQ_MODULUS = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
FR_MODULUS = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
_R_INV_Q = pow(1 << 256, -1, FQ_MODULUS)


def words_to_int(w) -> int:
    """4 little-endian u64 limbs -> int"""
    return sum(int(w[i]) << (64 * i) for i in range(4))


def fq_from_mont_words(w) -> int:
    return words_to_int(w) * _R_INV_Q % FQ_MODULUS


def jacobian_words_to_affine(jac):
    """12 u64 words x | y | z (Montgomery, z = 0: identity) -> (x, y) canonical ints, or None for the identity"""
    x, y, z = (fq_from_mont_words(jac[4 * i: 4 * i + 4]) for i in range(3))
    if z == 0:
        return None
    zi = pow(z, -1, FQ_MODULUS)
    zi2 = zi * zi % FQ_MODULUS
    return (x * zi2 % FQ_MODULUS, y * zi2 * zi % FQ_MODULUS)


def affine_add(p, q):
    """y^2 = x^3 + 3 over Fq, affine, None = identity"""
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % FQ_MODULUS == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, FQ_MODULUS) % FQ_MODULUS
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, FQ_MODULUS) % FQ_MODULUS
    x3 = (lam * lam - x1 - x2) % FQ_MODULUS
    return (x3, (lam * (x1 - x3) - y1) % FQ_MODULUS)


def scalar_mul_generator(c: int):
    """[c] G for G = (1, 2), double-and-add on canonical ints"""
    c %= FR_MODULUS
    acc, base = None, (1, 2)
    while c:
        if c & 1:
            acc = affine_add(acc, base)
        base = affine_add(base, base)
        c >>= 1
    return acc

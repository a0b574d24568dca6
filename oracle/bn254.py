"""
ORACLE (test infrastructure only -- never imported by the product path).

Pure-Python big-integer restatement of the BN254 arithmetic and of the two
hot-path functions `best_multiexp` / `best_fft` that DCMMC/halo2-scaffold
reaches through `halo2_proofs`.

PARITY UNPINNED: the reference tree holds no golden vector, known-answer test or
serialized proof for this path, and the arithmetic
lives in un-vendored git dependencies that are absent from /root/reference:
  * halo2_proofs  @ PSE tag v2023_02_02             (Cargo.toml:13)
  * halo2_proofs  @ axiom-crypto/halo2 `axiom/dev`  (Cargo.toml:16, via halo2-base)
  * halo2curves 0.3.x (transitive)
What pins this file instead: (1) both outputs are mathematically unique (affine
value of sum s_i*P_i; natural-order DFT), (2) external anchors that were not
produced by this code (EIP-196 generator (1,2) and 2*G, r*G = infinity, the
halo2curves constants ROOT_OF_UNITY / R / R2 / INV listed in SURVEY.md section 8),
(3) cross-checks between three independent algorithms (naive double-and-add MSM
vs. the Pippenger restatement; O(n^2) DFT vs. the recursive FFT restatement).

Reference call sites that reach this path: src/scaffold.rs:132,135,191-199,
207-214,223-230,284,287,322-346,354-361; examples/standard_plonk.rs:33,34,41-49,57-64.
"""
from __future__ import annotations

import math

# --------------------------------------------------------------------------
# constants ("Verified constants")
# --------------------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr modulus r
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq modulus p
MONT_R = 1 << 256
FR_S = 28                               # two-adicity of r-1
FR_GENERATOR = 7                        # multiplicative generator of Fr
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23  # cube root of unity (halo2curves Fr::ZETA)
CURVE_B = 3
G1_GEN = (1, 2)
MASK64 = (1 << 64) - 1


def mont_constants(mod: int):
    """R mod m, R^2 mod m, -m^{-1} mod 2^64 (the halo2curves `R`, `R2`, `INV`)."""
    return MONT_R % mod, (MONT_R * MONT_R) % mod, (-pow(mod, -1, 1 << 64)) % (1 << 64)


# --------------------------------------------------------------------------
# Montgomery <-> canonical, limb packing (4 x u64 little-endian, SURVEY 8 sizes)
# --------------------------------------------------------------------------
def to_mont(x: int, mod: int) -> int:
    return (x * MONT_R) % mod


def from_mont(x: int, mod: int) -> int:
    return (x * pow(MONT_R, -1, mod)) % mod


def int_to_limbs(x: int):
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def limbs_to_int(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


# --------------------------------------------------------------------------
# G1: y^2 = x^3 + 3 over Fq. Affine points are (x, y) tuples, infinity = None.
# (halo2curves encodes the affine identity as (0,0); see `affine_to_words`.)
# --------------------------------------------------------------------------
def is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - CURVE_B) % P_MOD == 0


def g1_neg(P):
    if P is None:
        return None
    return (P[0], (-P[1]) % P_MOD)


def g1_add(P, Q):
    """Affine addition, all special cases (P+P, P+(-P), identity)."""
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    y3 = (lam * (x1 - x3) - y1) % P_MOD
    return (x3, y3)


# Jacobian arithmetic (used for anything bigger than a handful of points)
def jac_from_affine(P):
    if P is None:
        return (0, 1, 0)
    return (P[0], P[1], 1)


def jac_to_affine(J):
    X, Y, Z = J
    if Z % P_MOD == 0:
        return None
    zi = pow(Z, -1, P_MOD)
    zi2 = zi * zi % P_MOD
    return (X * zi2 % P_MOD, Y * zi2 * zi % P_MOD)


def jac_double(J):
    X, Y, Z = J
    if Z == 0:
        return J
    A = X * X % P_MOD
    B = Y * Y % P_MOD
    C = B * B % P_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % P_MOD
    E = 3 * A % P_MOD
    F = E * E % P_MOD
    X3 = (F - 2 * D) % P_MOD
    Y3 = (E * (D - X3) - 8 * C) % P_MOD
    Z3 = 2 * Y * Z % P_MOD
    return (X3, Y3, Z3)


def jac_add(J1, J2):
    X1, Y1, Z1 = J1
    X2, Y2, Z2 = J2
    if Z1 == 0:
        return J2
    if Z2 == 0:
        return J1
    Z1Z1 = Z1 * Z1 % P_MOD
    Z2Z2 = Z2 * Z2 % P_MOD
    U1 = X1 * Z2Z2 % P_MOD
    U2 = X2 * Z1Z1 % P_MOD
    S1 = Y1 * Z2 * Z2Z2 % P_MOD
    S2 = Y2 * Z1 * Z1Z1 % P_MOD
    if U1 == U2:
        if S1 == S2:
            return jac_double(J1)
        return (0, 1, 0)
    H = (U2 - U1) % P_MOD
    Rr = (S2 - S1) % P_MOD
    HH = H * H % P_MOD
    HHH = H * HH % P_MOD
    V = U1 * HH % P_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % P_MOD
    Y3 = (Rr * (V - X3) - S1 * HHH) % P_MOD
    Z3 = Z1 * Z2 * H % P_MOD
    return (X3, Y3, Z3)


def jac_add_affine(J, P):
    return jac_add(J, jac_from_affine(P))


def g1_mul(P, k: int):
    """[k]P by double-and-add on Jacobian coordinates; k is reduced mod r."""
    k %= R_MOD
    acc = (0, 1, 0)
    base = jac_from_affine(P)
    while k:
        if k & 1:
            acc = jac_add(acc, base)
        base = jac_double(base)
        k >>= 1
    return jac_to_affine(acc)


def msm_naive(scalars, points):
    """sum_i [s_i] P_i, computed the slow obvious way (ground truth)."""
    acc = (0, 1, 0)
    for s, P in zip(scalars, points):
        if P is None or s % R_MOD == 0:
            continue
        acc = jac_add(acc, jac_from_affine(g1_mul(P, s)))
    return jac_to_affine(acc)


# --------------------------------------------------------------------------
# best_multiexp / multiexp_serial  -- [UP] halo2_proofs/src/arithmetic.rs @ v2023_02_02
# (restated from SURVEY.md Appendix B; source absent from /root/reference)
# --------------------------------------------------------------------------
def _get_at(segment: int, c: int, le_bytes: bytes) -> int:
    """[UP] multiexp_serial::get_at: c bits starting at bit segment*c of the 32-byte LE repr."""
    skip_bits = segment * c
    skip_bytes = skip_bits // 8
    if skip_bytes >= 32:
        return 0
    v = bytearray(8)
    chunk = le_bytes[skip_bytes:skip_bytes + 8]
    v[:len(chunk)] = chunk
    tmp = int.from_bytes(v, "little")
    tmp >>= skip_bits - skip_bytes * 8
    return tmp % (1 << c)


def multiexp_serial(coeffs, bases, acc=(0, 1, 0)):
    """[UP] multiexp_serial: unsigned-window Pippenger over one chunk. coeffs canonical ints."""
    m = len(coeffs)
    reprs = [int(s % R_MOD).to_bytes(32, "little") for s in coeffs]
    if m < 4:
        c = 1
    elif m < 32:
        c = 3
    else:
        c = int(math.ceil(math.log(m)))
    segments = 256 // c + 1
    for seg in range(segments - 1, -1, -1):
        for _ in range(c):
            acc = jac_double(acc)
        buckets = [None] * ((1 << c) - 1)       # None / Jacobian
        for rep, base in zip(reprs, bases):
            d = _get_at(seg, c, rep)
            if d != 0 and base is not None:
                b = buckets[d - 1]
                buckets[d - 1] = jac_from_affine(base) if b is None else jac_add_affine(b, base)
        running = (0, 1, 0)
        for b in reversed(buckets):
            if b is not None:
                running = jac_add(running, b)
            acc = jac_add(acc, running)
    return acc


def best_multiexp(coeffs, bases, num_threads: int = 1):
    """[UP] best_multiexp: chunk-per-thread Pippenger, partial sums folded. Returns affine."""
    assert len(coeffs) == len(bases)
    n = len(coeffs)
    if n > num_threads:
        chunk = n // num_threads
        acc = (0, 1, 0)
        for i in range(0, n, chunk):
            acc = jac_add(acc, multiexp_serial(coeffs[i:i + chunk], bases[i:i + chunk]))
        return jac_to_affine(acc)
    return jac_to_affine(multiexp_serial(coeffs, bases))


# --------------------------------------------------------------------------
# best_fft  -- [UP] halo2_proofs/src/arithmetic.rs @ v2023_02_02
# --------------------------------------------------------------------------
def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def dft_naive(a, omega):
    """O(n^2) natural-order DFT: out[i] = sum_j a[j] * omega^(i*j). Ground truth."""
    n = len(a)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * omega % R_MOD
    return [sum(a[j] * pw[(i * j) % n] for j in range(n)) % R_MOD for i in range(n)]


def _recursive_butterfly(a, lo, n, twiddle_chunk, twiddles):
    """[UP] recursive_butterfly_arithmetic (rayon::join replaced by plain recursion)."""
    if n == 2:
        t = a[lo + 1]
        a[lo + 1] = (a[lo] - t) % R_MOD
        a[lo] = (a[lo] + t) % R_MOD
        return
    half = n // 2
    _recursive_butterfly(a, lo, half, twiddle_chunk * 2, twiddles)
    _recursive_butterfly(a, lo + half, half, twiddle_chunk * 2, twiddles)
    for i in range(half):
        t = a[lo + half + i]
        if i:
            t = t * twiddles[i * twiddle_chunk] % R_MOD
        a[lo + half + i] = (a[lo + i] - t) % R_MOD
        a[lo + i] = (a[lo + i] + t) % R_MOD


def best_fft(a, omega: int, log_n: int):
    """[UP] best_fft for G = Fr: in place, natural order in/out, no scaling. a: canonical ints."""
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    if n == 1:
        return a
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % R_MOD
    _recursive_butterfly(a, 0, n, 1, tw)
    return a


# --------------------------------------------------------------------------
# EvaluationDomain  -- [UP] halo2_proofs/src/poly/domain.rs (callers of best_fft)
# --------------------------------------------------------------------------
class EvaluationDomain:
    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        self.extended_k = ek
        w = FR_ROOT_OF_UNITY
        for _ in range(ek, FR_S):
            w = w * w % R_MOD
        self.extended_omega = w
        for _ in range(k, ek):
            w = w * w % R_MOD
        self.omega = w
        self.omega_inv = pow(self.omega, -1, R_MOD)
        self.extended_omega_inv = pow(self.extended_omega, -1, R_MOD)
        self.ifft_divisor = pow(1 << k, -1, R_MOD)
        self.extended_ifft_divisor = pow(1 << ek, -1, R_MOD)
        self.g_coset = FR_ZETA
        self.g_coset_inv = FR_ZETA * FR_ZETA % R_MOD
        # [UP] domain.rs: evaluations of t(X) = X^n - 1 over the zeta coset, one period, inverted
        orig, step = pow(FR_ZETA, self.n, R_MOD), pow(self.extended_omega, self.n, R_MOD)
        cur, t = orig, []
        while True:
            t.append(cur)
            cur = cur * step % R_MOD
            if cur == orig:
                break
        self.t_evaluations = [pow(v - 1, -1, R_MOD) for v in t]

    def divide_by_vanishing_poly(self, a):
        return [x * self.t_evaluations[i % len(self.t_evaluations)] % R_MOD for i, x in enumerate(a)]

    def lagrange_to_coeff(self, a):
        a = list(a)
        best_fft(a, self.omega_inv, self.k)
        return [x * self.ifft_divisor % R_MOD for x in a]

    def coeff_to_extended(self, a):
        z = [1, self.g_coset, self.g_coset_inv]
        a = [x * z[i % 3] % R_MOD for i, x in enumerate(a)]
        a += [0] * ((1 << self.extended_k) - len(a))
        return best_fft(a, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a):
        a = list(a)
        best_fft(a, self.extended_omega_inv, self.extended_k)
        z = [1, self.g_coset_inv, self.g_coset]
        a = [x * self.extended_ifft_divisor % R_MOD * z[i % 3] % R_MOD for i, x in enumerate(a)]
        return a[: self.n * self.quotient_poly_degree]


def omega_for(log_n: int) -> int:
    """omega of order 2^log_n as EvaluationDomain::new derives it (repeated squaring of ROOT_OF_UNITY)."""
    w = FR_ROOT_OF_UNITY
    for _ in range(log_n, FR_S):
        w = w * w % R_MOD
    return w


# --------------------------------------------------------------------------
# deterministic synthetic inputs: SplitMix64
# --------------------------------------------------------------------------
def splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & MASK64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return state, z ^ (z >> 31)


def random_fr(seed: int, n: int):
    """n canonical Fr values: 512-bit SplitMix64 draw reduced mod r."""
    out = []
    st = seed & MASK64
    for _ in range(n):
        v = 0
        for _ in range(8):
            st, w = splitmix64(st)
            v = (v << 64) | w
        out.append(v % R_MOD)
    return out


# ---------------------------------------------------------------------------------------------------------
# Quotient evaluation ([UP] halo2_proofs/src/plonk/evaluation.rs): GraphEvaluator::evaluate on canonical
# integers, one row at a time -- the independent (big-int) check of oracle/h2_oracle.cpp's restatement.
# Encoding as in h2_oracle.cpp: value source (kind, index, rotation); calculation
# (op, target, a[3], b[3], parts_offset, parts_len).
# ---------------------------------------------------------------------------------------------------------
def graph_evaluate_row(constants, rotations, calcs, parts, n_intermediates, fixed, advice, instance, challenges, beta, gamma, theta, y,
                       previous_value, idx, rot_scale, isize):
    rot = [(idx + r * rot_scale) % isize for r in rotations]          # get_rotation_idx (Python % is rem_euclid)
    inter = [0] * n_intermediates

    def get(vs):
        kind, index, rotation = vs
        if kind == 0:
            return constants[index]
        if kind == 1:
            return inter[index]
        if kind == 2:
            return fixed[index][rot[rotation]]
        if kind == 3:
            return advice[index][rot[rotation]]
        if kind == 4:
            return instance[index][rot[rotation]]
        if kind == 5:
            return challenges[index]
        return {6: beta, 7: gamma, 8: theta, 9: y, 10: previous_value}[kind]

    last = None
    for c in calcs:
        op, target, a, b, po, pl = c[0], c[1], tuple(c[2:5]), tuple(c[5:8]), c[8], c[9]
        if op == 0:
            v = get(a) + get(b)
        elif op == 1:
            v = get(a) - get(b)
        elif op == 2:
            v = get(a) * get(b)
        elif op == 3:
            v = get(a) ** 2
        elif op == 4:
            v = 2 * get(a)
        elif op == 5:
            v = -get(a)
        elif op == 6:
            v, f = get(a), get(b)
            for j in range(pl):
                v = (v * f + get(tuple(parts[po + j]))) % R_MOD
        else:
            v = get(a)
        inter[target] = v % R_MOD
        last = target
    return inter[last] if last is not None else 0


# ---------------------------------------------------------------------------------------------------------
# SRS on-disk format ([UP] halo2_proofs/src/poly/kzg/commitment.rs ParamsKZG::{read_custom, write_custom};
# [UP] halo2curves 0.3.x GroupEncoding / SerdeObject for G1Affine).  Points are canonical (x, y) or None.
# File: k (u32 LE) | g[2^k] | g_lagrange[2^k] | g2 | s_g2 (the G2 section is carried as opaque bytes).
# ---------------------------------------------------------------------------------------------------------
SERDE_PROCESSED, SERDE_RAW_BYTES, SERDE_RAW_BYTES_UNCHECKED = 0, 1, 2


def g1_to_bytes(P) -> bytes:
    """G1Affine::to_bytes: canonical little-endian x, (y & 1) << 7 in byte 31; the identity is all-zero"""
    if P is None:
        return bytes(32)
    b = bytearray(P[0].to_bytes(32, "little"))
    b[31] |= (P[1] & 1) << 7
    return bytes(b)


def g1_from_bytes(b: bytes):
    """G1Affine::from_bytes -> point, None for the identity; raises ValueError on an invalid encoding"""
    assert len(b) == 32
    ysign = b[31] >> 7
    x = int.from_bytes(b, "little") & ((1 << 255) - 1)
    if x >= P_MOD:
        raise ValueError("x is not canonical")
    if x == 0 and not ysign:
        return None
    rhs = (x * x * x + CURVE_B) % P_MOD
    y = pow(rhs, (P_MOD + 1) // 4, P_MOD)          # p = 3 mod 4
    if y * y % P_MOD != rhs:
        raise ValueError("x^3 + 3 is not a square")
    if (y & 1) != ysign:
        y = (-y) % P_MOD
    return (x, y)


def g1_write_raw(P) -> bytes:
    """SerdeObject::write_raw: the Montgomery limbs of x and y as they sit in memory; identity = (0, 0)"""
    if P is None:
        return bytes(64)
    return to_mont(P[0], P_MOD).to_bytes(32, "little") + to_mont(P[1], P_MOD).to_bytes(32, "little")


def g1_read_raw(b: bytes, checked: bool = True):
    assert len(b) == 64
    xm, ym = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:], "little")
    if checked and (xm >= P_MOD or ym >= P_MOD):
        raise ValueError("limbs are not below the modulus")
    if xm == 0 and ym == 0:
        return None
    P = (from_mont(xm, P_MOD), from_mont(ym, P_MOD))
    if checked and not is_on_curve(P):
        raise ValueError("not on the curve")
    return P


def srs_write(path, k, g, g_lagrange, g2_bytes: bytes, fmt: int = SERDE_RAW_BYTES):
    enc = g1_to_bytes if fmt == SERDE_PROCESSED else g1_write_raw
    with open(path, "wb") as f:
        f.write(int(k).to_bytes(4, "little"))
        for v in (g, g_lagrange):
            assert len(v) == 1 << k
            f.write(b"".join(enc(P) for P in v))
        f.write(g2_bytes)


def srs_read(path, fmt: int = SERDE_RAW_BYTES):
    with open(path, "rb") as f:
        k = int.from_bytes(f.read(4), "little")
        n, ps = 1 << k, 32 if fmt == SERDE_PROCESSED else 64
        out = []
        for _ in range(2):
            raw = f.read(n * ps)
            assert len(raw) == n * ps
            if fmt == SERDE_PROCESSED:
                out.append([g1_from_bytes(raw[i * ps:(i + 1) * ps]) for i in range(n)])
            else:
                out.append([g1_read_raw(raw[i * ps:(i + 1) * ps], fmt == SERDE_RAW_BYTES) for i in range(n)])
        return k, out[0], out[1], f.read()


# ---------------------------------------------------------------------------------------------------------
# Grand products and the lookup permutation on canonical integers -- the independent check of h2_oracle.cpp's
# restatements ([UP] plonk/permutation/prover.rs Argument::commit, plonk/lookup/prover.rs commit_product and
# permute_expression_pair).  Written from the formulas of the arguments, not from the C++ code.
# ---------------------------------------------------------------------------------------------------------
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R_MOD)


def permutation_product(values, sigma, beta, gamma, deltaomega, omega, last_z):
    """z[0] = last_z; z[i+1] = z[i] * prod_j (v_j[i] + deltaomega delta^j omega^i beta + gamma) / prod_j (v_j[i] + beta s_j[i] + gamma)"""
    n = len(values[0])
    z, run = [], last_z % R_MOD
    for i in range(n):
        z.append(run)
        num = den = 1
        for j, (v, s) in enumerate(zip(values, sigma)):
            num = num * (v[i] + deltaomega * pow(FR_DELTA, j, R_MOD) * pow(omega, i, R_MOD) * beta + gamma) % R_MOD
            den = den * (v[i] + beta * s[i] + gamma) % R_MOD
        run = run * num % R_MOD * (pow(den, -1, R_MOD) if den else 0) % R_MOD
    return z


def lookup_product(a, s, a_perm, s_perm, beta, gamma):
    """z[0] = 1; z[i+1] = z[i] (a[i] + beta)(s[i] + gamma) / ((a'[i] + beta)(s'[i] + gamma))"""
    z, run = [], 1
    for i in range(len(a)):
        z.append(run)
        den = (a_perm[i] + beta) * (s_perm[i] + gamma) % R_MOD
        run = run * (a[i] + beta) % R_MOD * (s[i] + gamma) % R_MOD * (pow(den, -1, R_MOD) if den else 0) % R_MOD
    return z


def permute_expression_pair(a, s, usable_rows):
    """-> (A', S') over the first usable_rows rows; raises ValueError if an input value is missing from the table"""
    from collections import Counter
    a_sorted = sorted(a[:usable_rows])
    leftover = Counter(s[:usable_rows])
    s_perm = [None] * usable_rows
    repeated = []
    for row, v in enumerate(a_sorted):
        if row == 0 or v != a_sorted[row - 1]:
            if leftover[v] <= 0:
                raise ValueError("ConstraintSystemFailure")
            leftover[v] -= 1
            s_perm[row] = v
        else:
            repeated.append(row)
    for v in sorted(leftover):                      # BTreeMap order
        for _ in range(leftover[v]):
            s_perm[repeated.pop()] = v
    assert not repeated
    return a_sorted, s_perm

#!/usr/bin/env python3
"""
Generates tests/golden/*.npz with the pure-Python big-int oracle (oracle/bn254.py): naive
double-and-add MSM and the O(n^2) DFT -- i.e. algorithms that share nothing with the Pippenger / FFT
code they pin.  The reference tree has no golden vector for this path ("parity
unpinned"), so these fixtures anchor the C++ oracle, the emulated kernels and the GPU path to the
same independently computed values.  Deterministic: re-running reproduces the committed files.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bn254 as o  # noqa: E402


def words(vals, mod):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = o.to_mont(v, mod)
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & o.MASK64
    return out


def affine_words(points):
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, P in enumerate(points):
        if P is None:
            continue
        out[i, :4] = words([P[0]], o.P_MOD)[0]
        out[i, 4:] = words([P[1]], o.P_MOD)[0]
    return out


def make_ntt():
    data = {}
    for k in (1, 2, 3, 5, 8):
        n = 1 << k
        a = o.random_fr(0x900D0000 + k, n)
        w = o.omega_for(k)
        wi = pow(w, -1, o.R_MOD)
        data["k%d_in" % k] = words(a, o.R_MOD)
        data["k%d_omega" % k] = words([w], o.R_MOD)[0]
        data["k%d_omega_inv" % k] = words([wi], o.R_MOD)[0]
        data["k%d_fwd" % k] = words(o.dft_naive(a, w), o.R_MOD)
        data["k%d_inv" % k] = words(o.dft_naive(a, wi), o.R_MOD)
    np.savez_compressed(os.path.join(HERE, "ntt_golden.npz"), **data)


def make_msm():
    data = {}
    st = 0x5EED
    for n in (1, 2, 8, 64, 200):
        scal = o.random_fr(0x900D1000 + n, n)
        pts = []
        for _ in range(n):
            st, z = o.splitmix64(st)
            pts.append(o.g1_mul(o.G1_GEN, z))
        if n >= 8:
            scal[0] = 0                       # zero scalar
            scal[1] = 1                       # one
            scal[2] = o.R_MOD - 1             # -1
            scal[3] = (o.R_MOD - 1) // 2      # boundary of the sign trick
            scal[4] = (o.R_MOD + 1) // 2
            pts[5] = None                     # identity base
            pts[7] = pts[6]                   # duplicate base
            scal[7] = scal[6]                 # ... with equal scalar (P + P inside a bucket)
        if n >= 64:
            pts[9] = o.g1_neg(pts[8])         # P and -P with equal scalars (cancels to the identity)
            scal[9] = scal[8]
            for i in range(20, 40):
                scal[i] = i - 19              # small scalars
            for i in range(40, 50):
                scal[i] = o.R_MOD - (i - 39)  # small negatives
        res = o.msm_naive(scal, pts)
        data["n%d_scalars" % n] = words(scal, o.R_MOD)
        data["n%d_bases" % n] = affine_words(pts)
        data["n%d_result" % n] = affine_words([res])[0]
    # all-cancelling input: result is the identity
    P = o.g1_mul(o.G1_GEN, 12345)
    data["cancel_scalars"] = words([5, 5], o.R_MOD)
    data["cancel_bases"] = affine_words([P, o.g1_neg(P)])
    data["cancel_result"] = affine_words([None])[0]
    # external anchor (EIP-196): 2 * (1, 2)
    data["anchor_scalars"] = words([2], o.R_MOD)
    data["anchor_bases"] = affine_words([o.G1_GEN])
    data["anchor_result"] = affine_words([(1368015179489954701390400359078579693043519447331113978918064868415326638035,
                                           9918110051302171585080402603319702774565515993150576347155970296011118125764)])[0]
    np.savez_compressed(os.path.join(HERE, "msm_golden.npz"), **data)


def make_domain():
    d = o.EvaluationDomain(4, 4)
    a = o.random_fr(0x900D2000, 16)
    coeff = d.lagrange_to_coeff(a)
    ext = d.coeff_to_extended(coeff)
    back = d.extended_to_coeff(ext)
    np.savez_compressed(os.path.join(HERE, "domain_golden.npz"), j=4, k=4, lagrange=words(a, o.R_MOD), coeff=words(coeff, o.R_MOD),
                        extended=words(ext, o.R_MOD), back=words(back, o.R_MOD))


def make_prover():
    """Fixtures of the widened rows, all from the big-int code of oracle/bn254.py: GraphEvaluator walk, permutation and
    lookup grand products, permute_expression_pair, divide_by_vanishing_poly, the G1 point encodings."""
    data = {}
    n = 16
    # a small GraphEvaluator: q * (a + b[next] * c - d[prev]) combined with y into the previous value, via every op kind
    consts = [0, 1, 2, 0x1234567]
    rotations = [0, 1, -1]
    calcs = [[7, 0, 2, 0, 0, 0, 0, 0, 0, 0],          # t0 = fixed0
             [2, 1, 3, 1, 1, 3, 2, 0, 0, 0],          # t1 = advice1[next] * advice2
             [0, 2, 3, 0, 0, 1, 1, 0, 0, 0],          # t2 = advice0 + t1
             [1, 3, 1, 2, 0, 3, 0, 2, 0, 0],          # t3 = t2 - advice0[prev]
             [2, 4, 1, 0, 0, 1, 3, 0, 0, 0],          # t4 = t0 * t3
             [3, 5, 1, 4, 0, 0, 0, 0, 0, 0],          # t5 = t4^2
             [4, 6, 4, 0, 0, 0, 0, 0, 0, 0],          # t6 = 2 * instance0
             [5, 7, 5, 0, 0, 0, 0, 0, 0, 0],          # t7 = -challenge0
             [0, 8, 0, 3, 0, 6, 0, 0, 0, 0],          # t8 = const3 + beta      (dead)
             [6, 9, 10, 0, 0, 9, 0, 0, 0, 4]]         # t9 = Horner(previous, [t4, t5, t6, t7], y)
    parts = [[1, 4, 0], [1, 5, 0], [1, 6, 0], [1, 7, 0]]
    cols = [o.random_fr(0x900D2000 + j, n) for j in range(5)]          # fixed0, advice0..2, instance0
    sc = o.random_fr(0x900D2100, 6)                                     # challenge0, beta, gamma, theta, y + spare
    prev = o.random_fr(0x900D2200, n)
    out = [o.graph_evaluate_row(consts, rotations, calcs, parts, 10, cols[:1], cols[1:4], cols[4:], sc[:1], sc[1], sc[2], sc[3], sc[4], prev[i], i, 2, n)
           for i in range(n)]
    data["graph_constants"] = words(consts, o.R_MOD)
    data["graph_rotations"] = np.array(rotations, dtype=np.int32)
    data["graph_calcs"] = np.array(calcs, dtype=np.uint32)
    data["graph_parts"] = np.array(parts, dtype=np.uint32)
    data["graph_cols"] = np.stack([words(c, o.R_MOD) for c in cols])
    data["graph_scalars"] = words(sc, o.R_MOD)
    data["graph_prev"] = words(prev, o.R_MOD)
    data["graph_out"] = words(out, o.R_MOD)
    # grand products
    k = 4
    w = o.omega_for(k)
    vals = [o.random_fr(0x900D3000 + j, n) for j in range(3)]
    sig = [o.random_fr(0x900D3100 + j, n) for j in range(3)]
    beta, gamma, last_z = o.random_fr(0x900D3200, 3)
    dw = pow(o.FR_DELTA, 2, o.R_MOD)
    data["perm_values"] = np.stack([words(v, o.R_MOD) for v in vals])
    data["perm_sigma"] = np.stack([words(v, o.R_MOD) for v in sig])
    data["perm_scalars"] = words([beta, gamma, o.FR_DELTA, dw, w, last_z], o.R_MOD)
    data["perm_z"] = words(o.permutation_product(vals, sig, beta, gamma, dw, w, last_z), o.R_MOD)
    lk = [o.random_fr(0x900D3300 + j, n) for j in range(4)]
    data["lookup_cols"] = np.stack([words(v, o.R_MOD) for v in lk])
    data["lookup_z"] = words(o.lookup_product(*lk, beta, gamma), o.R_MOD)
    # permute_expression_pair: 13 usable rows, table with duplicates, input with repeats, a negative value
    table = [5, 0, 3, 3, o.R_MOD - 1, 9, 2**200 + 7, 1, 4, 5, 5, 8, 6, 77, 78, 79]
    inp = [3, 5, 5, o.R_MOD - 1, 0, 0, 0, 2**200 + 7, 9, 3, 1, 5, 8, 111, 112, 113]
    pa, pt = o.permute_expression_pair(inp, table, 13)
    data["permute_input"] = words(inp, o.R_MOD)
    data["permute_table"] = words(table, o.R_MOD)
    data["permute_out_input"] = words(pa, o.R_MOD)
    data["permute_out_table"] = words(pt, o.R_MOD)
    # divide_by_vanishing_poly for j = 4, k = 3 (period 4)
    dom = o.EvaluationDomain(4, 3)
    a = o.random_fr(0x900D3400, 1 << dom.extended_k)
    data["vanishing_in"] = words(a, o.R_MOD)
    data["vanishing_out"] = words(dom.divide_by_vanishing_poly(a), o.R_MOD)
    # point encodings: generator, its double, a negation, the identity, larger multiples
    pts = [o.G1_GEN, o.g1_add(o.G1_GEN, o.G1_GEN), o.g1_neg(o.G1_GEN), None] + [o.g1_mul(o.G1_GEN, 0xABCDEF + 977 * i) for i in range(12)]
    data["codec_points"] = affine_words(pts)
    data["codec_bytes"] = np.frombuffer(b"".join(o.g1_to_bytes(P) for P in pts), dtype=np.uint8).reshape(-1, 32).copy()
    np.savez_compressed(os.path.join(HERE, "prover_golden.npz"), **data)


if __name__ == "__main__":
    make_ntt()
    make_msm()
    make_domain()
    make_prover()
    print("golden fixtures written to", HERE)

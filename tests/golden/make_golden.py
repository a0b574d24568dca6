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


if __name__ == "__main__":
    make_ntt()
    make_msm()
    make_domain()
    print("golden fixtures written to", HERE)

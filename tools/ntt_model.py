#!/usr/bin/env python3
"""
Index-math model of the multi-pass NTT in csrc/ntt.cu, checked against the O(n^2) DFT of
oracle/bn254.py.  (Development aid: documents and proves the pass decomposition.)

n = N_1 * ... * N_P.  Pass p works on segments of length L_p = n / (N_1..N_{p-1}); inside a segment
it runs an N_p-point DIF over the top digit (stride M_p = L_p / N_p), leaves the digit in
bit-reversed order *in place*, and multiplies element (q, jr) by w^(S_p * jr * bitrev(q)),
S_p = n / L_p.  After the last pass position `pos` holds out[bitrev_k(pos)].
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import bn254 as o

R = o.R_MOD

def brev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1); x >>= 1
    return r

def dif_inplace(v, omega_small):
    """N-point radix-2 DIF, output left in bit-reversed order. v: list, omega_small of order N."""
    N = len(v); h = N // 2
    tw = [pow(omega_small, t, R) for t in range(max(N // 2, 1))]
    while h >= 1:
        for bq in range(N // 2):
            q = (bq // h) * 2 * h + (bq % h)
            x, y = v[q], v[q + h]
            v[q] = (x + y) % R
            v[q + h] = (x - y) * tw[(bq % h) * (N // (2 * h))] % R
        h //= 2

def ntt_multipass(a, omega, bits):
    k = sum(bits); n = 1 << k
    assert len(a) == n
    a = list(a)
    L = n
    for p, b in enumerate(bits):
        N = 1 << b; M = L // N; S = n // L
        om_small = pow(omega, n // N, R)
        for seg in range(n // L):
            base = seg * L
            for jr in range(M):
                v = [a[base + q * M + jr] for q in range(N)]
                dif_inplace(v, om_small)
                for q in range(N):
                    i_p = brev(q, b)
                    if M > 1:
                        v[q] = v[q] * pow(omega, S * jr * i_p, R) % R
                    a[base + q * M + jr] = v[q]
        L = M
    out = [0] * n
    for pos in range(n):
        out[brev(pos, k)] = a[pos]
    return out

if __name__ == "__main__":
    for bits in ([1], [3], [2, 2], [3, 2], [2, 3, 1], [3, 3, 2], [1, 1, 1, 1], [4, 3]):
        k = sum(bits); n = 1 << k
        a = o.random_fr(k * 31 + len(bits), n)
        w = o.omega_for(k)
        assert ntt_multipass(a, w, bits) == o.dft_naive(a, w), bits
        wi = pow(w, -1, R)
        assert ntt_multipass(a, wi, bits) == o.dft_naive(a, wi), bits
    print("ntt pass model OK")

"""CPU: the oracle itself -- big-int Python vs external anchors, C++ restatement vs Python, both vs golden fixtures."""
import numpy as np
import pytest

import bn254 as o


def test_constants_match_halo2curves_anchors():
    # constants of halo2curves bn256::{Fr,Fq} (external, not produced by this code)
    assert o.FR_ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert pow(o.FR_ROOT_OF_UNITY, 1 << 28, o.R_MOD) == 1 and pow(o.FR_ROOT_OF_UNITY, 1 << 27, o.R_MOD) != 1
    r1, r2, inv = o.mont_constants(o.R_MOD)
    assert r1 == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert r2 == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert inv == 0xC2E1F593EFFFFFFF
    q1, q2, qinv = o.mont_constants(o.P_MOD)
    assert q1 == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert q2 == 0x06D89F71CAB8351F47AB1EFF0A417FF6B5E71911D44501FBF32CFC5B538AFA89
    assert qinv == 0x87D20782E4866389
    assert pow(o.FR_ZETA, 3, o.R_MOD) == 1 and o.FR_ZETA != 1


def test_group_anchors():
    # EIP-196 generator and its double; group order
    assert o.is_on_curve(o.G1_GEN)
    assert o.g1_add(o.G1_GEN, o.G1_GEN) == (
        1368015179489954701390400359078579693043519447331113978918064868415326638035,
        9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert o.g1_mul(o.G1_GEN, o.R_MOD - 1) == o.g1_neg(o.G1_GEN)
    assert o.g1_add(o.g1_mul(o.G1_GEN, o.R_MOD - 1), o.G1_GEN) is None


def test_python_pippenger_equals_naive():
    import random
    rnd = random.Random(5)
    for n in (1, 3, 4, 31, 32, 70):
        pts = [o.g1_mul(o.G1_GEN, rnd.randrange(1, 1 << 64)) for _ in range(n)]
        sc = o.random_fr(n, n)
        if n > 4:
            sc[0] = 0
            pts[2] = None
            pts[4] = pts[3]
        want = o.msm_naive(sc, pts)
        for threads in (1, 4):
            assert o.best_multiexp(sc, pts, threads) == want


def test_python_fft_equals_naive_dft():
    for k in range(0, 8):
        a = o.random_fr(31 + k, 1 << k)
        for w in (o.omega_for(k), pow(o.omega_for(k), -1, o.R_MOD)):
            assert o.best_fft(list(a), w, k) == o.dft_naive(a, w)


def test_c_oracle_streams_match_python(oc):
    a = oc.random_fr(123, 64)
    assert [o.from_mont(x, o.R_MOD) for x in oc.words_to_ints(a)] == o.random_fr(123, 64)
    pts = oc.words_to_ints(oc.gen_points(7, 12).reshape(-1, 4))
    st = 7
    for i in range(12):
        st, z = o.splitmix64(st)
        assert (o.from_mont(pts[2 * i], o.P_MOD), o.from_mont(pts[2 * i + 1], o.P_MOD)) == o.g1_mul(o.G1_GEN, z)


def test_c_oracle_against_golden_ntt(oc, golden):
    g = golden["ntt"]
    for k in (1, 2, 3, 5, 8):
        for threads in (1, 3, 8):
            assert (oc.best_fft(g["k%d_in" % k], g["k%d_omega" % k], k, threads) == g["k%d_fwd" % k]).all()
            assert (oc.best_fft(g["k%d_in" % k], g["k%d_omega_inv" % k], k, threads) == g["k%d_inv" % k]).all()


def test_c_oracle_against_golden_msm(oc, golden):
    g = golden["msm"]
    for tag in ("n1", "n2", "n8", "n64", "n200", "cancel", "anchor"):
        for threads in (1, 3, 8):
            got = oc.g1_to_affine(oc.best_multiexp(g[tag + "_scalars"], g[tag + "_bases"], threads))
            assert (got == g[tag + "_result"]).all(), (tag, threads)


def test_c_oracle_larger_cross_check(oc):
    # thread-count independence (chunking) and iNTT(NTT(a)) = n * a
    n = 1 << 12
    s, P = oc.random_fr(3, n), oc.gen_points(4, n)
    r1 = oc.g1_to_affine(oc.best_multiexp(s, P, 1))
    r8 = oc.g1_to_affine(oc.best_multiexp(s, P, 8))
    assert (r1 == r8).all()
    k = 12
    w = oc.ints_to_words([o.to_mont(o.omega_for(k), o.R_MOD)])[0]
    wi = oc.ints_to_words([o.to_mont(pow(o.omega_for(k), -1, o.R_MOD), o.R_MOD)])[0]
    ninv = oc.ints_to_words([o.to_mont(pow(n, -1, o.R_MOD), o.R_MOD)])[0]
    back = oc.fr_scale(oc.best_fft(oc.best_fft(s, w, k), wi, k), ninv)
    assert (back == s).all()


def test_domain_golden_matches_python_oracle(golden):
    g = golden["domain"]
    d = o.EvaluationDomain(int(g["j"]), int(g["k"]))
    assert d.extended_k == 6
    from oracle_c import words_to_ints
    lag = [o.from_mont(x, o.R_MOD) for x in words_to_ints(g["lagrange"])]
    coeff = d.lagrange_to_coeff(lag)
    assert [o.to_mont(x, o.R_MOD) for x in coeff] == words_to_ints(g["coeff"])
    # extended_to_coeff(coeff_to_extended(p)) = p (zero padded)
    back = [o.from_mont(x, o.R_MOD) for x in words_to_ints(g["back"])]
    assert back[:16] == coeff and all(v == 0 for v in back[16:])


def test_graph_evaluator_restatement_matches_big_int_walk(oc):
    """oracle/h2_oracle.cpp's GraphEvaluator::evaluate against the independent big-int interpreter of oracle/bn254.py"""
    import bn254 as o
    import parity_cases as pc
    for seed, size, n_calcs in ((1, 9, 6), (2, 16, 40), (5, 12, 90), (8, 7, 25)):
        rng = np.random.default_rng(seed)
        graph, n_const = pc.random_graph(rng, 2, 3, 1, 2, n_calcs, reuse_targets=bool(seed % 3 == 2))
        g, _ = pc._graph_pair(oc, graph, n_const, seed)
        cols = [oc.random_fr(seed * 31 + j, size) for j in range(6)]
        sc = oc.random_fr(seed * 7 + 1, 6)
        values = oc.random_fr(seed * 7 + 2, size)
        got = oc.evaluate_graph(g, cols[:2], cols[2:5], cols[5:], sc[:2], *sc[2:], values, 3)

        def ints(a):
            return [o.from_mont(v, o.R_MOD) for v in oc.words_to_ints(np.ascontiguousarray(a).reshape(-1, 4))]
        ci = [ints(c) for c in cols]
        si = ints(sc)
        vi = ints(values)
        for idx in range(size):
            w = o.graph_evaluate_row(ints(g[0]), [int(r) for r in g[1]], g[2].tolist(), g[3].tolist(), g[4], ci[:2], ci[2:5], ci[5:], si[:2],
                                     si[2], si[3], si[4], si[5], vi[idx], idx, 3, size)
            assert o.from_mont(oc.words_to_ints(got[idx:idx + 1])[0], o.R_MOD) == w, (seed, idx)


def test_g1_encoding_anchors(oc):
    """G1Affine::to_bytes / from_bytes of the restatements on the EIP-196 anchors: G = (1, 2) and 2G"""
    import bn254 as o
    G = (1, 2)
    G2 = o.g1_add(G, G)
    assert G2 == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                  9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert o.g1_to_bytes(G) == bytes([1] + [0] * 31) and o.g1_to_bytes(None) == bytes(32)
    b2 = bytearray(G2[0].to_bytes(32, "little"))
    b2[31] |= (G2[1] & 1) << 7
    assert o.g1_to_bytes(G2) == bytes(b2)
    for P in (G, G2, o.g1_neg(G), o.g1_neg(G2), None, o.g1_mul(G, 0xDEADBEEF)):
        assert o.g1_from_bytes(o.g1_to_bytes(P)) == P
        assert o.g1_read_raw(o.g1_write_raw(P)) == P
    # the C++ restatement agrees with the big-int one, point by point
    pts = [o.g1_mul(G, 3 + 7 * i) for i in range(20)] + [None]
    raw = np.frombuffer(b"".join(o.g1_write_raw(P) for P in pts), dtype=np.uint64).reshape(-1, 8)
    enc = oc.g1_to_bytes(raw)
    assert enc.tobytes() == b"".join(o.g1_to_bytes(P) for P in pts)
    dec, first = oc.g1_from_bytes(enc)
    assert first == len(pts) and (dec == raw).all()
    with pytest.raises(ValueError):
        o.g1_from_bytes((o.P_MOD + 5).to_bytes(32, "little"))


def test_prover_rows_golden(oc, golden):
    import parity_cases as pc
    pc.check_golden_prover_oracle(oc, golden["prover"])

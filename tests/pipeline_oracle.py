"""
CPU recomputation of tools/proof_pipeline_core.py with the oracle (test infrastructure): the same sequence of prover steps
([UP] halo2_proofs plonk/prover.rs order) on the recorded inputs, every commitment and every evaluation.
Each step calls the oracle's restatement of the upstream function (oracle/h2_oracle.cpp via oracle_c) -- none of the product code.
"""
import math

import numpy as np

import bn254 as o

BLINDING = 6
DELTA = pow(7, 1 << 28, o.R_MOD)


def W(v):
    """canonical int -> Montgomery 4 x u64"""
    m = o.to_mont(v % o.R_MOD, o.R_MOD)
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def scale_pattern(oc, a, factors):
    """a[i] *= factors[i % len(factors)]"""
    f = np.stack(factors)
    reps = (a.shape[0] + f.shape[0] - 1) // f.shape[0]
    return oc.field_op("fr", "mul", np.ascontiguousarray(a), np.ascontiguousarray(np.tile(f, (reps, 1))[:a.shape[0]]))


def lincomb(oc, cols, coeffs):
    acc = None
    for c, w in zip(cols, coeffs):
        t = oc.fr_scale(c, w)
        acc = t if acc is None else oc.field_op("fr", "add", acc, t)
    return acc


def oracle_pipeline(oc, shape, inputs, evaluator):
    """-> (commitments as affine (m, 8) words in pipeline order, evaluations (e, 4) words)"""
    k, A, LK, d = shape["k"], shape["A"], shape["LK"], shape["d"]
    n = 1 << k
    dom = o.EvaluationDomain(d, k)
    ek, en = dom.extended_k, 1 << dom.extended_k
    rot_scale = 1 << (ek - k)
    chunk, n_adv = d - 2, A + LK
    sets = math.ceil(n_adv / chunk)
    usable = n - BLINDING
    g, gl = inputs["g"], inputs["g_lagrange"]
    theta, beta, gamma, y, x, v = inputs["challenges"][:6]
    weights = inputs["weights"]
    commitments, evals = [], []

    def commit(col, bases):
        col = np.ascontiguousarray(col)
        commitments.append(oc.g1_to_affine(oc.best_multiexp(col, np.ascontiguousarray(bases[:col.shape[0]]))))

    def to_coeff(col):
        return oc.fr_scale(oc.best_fft(col, W(dom.omega_inv), k), W(dom.ifft_divisor))

    def to_ext(c):
        e = np.zeros((en, 4), dtype=np.uint64)
        e[:n] = scale_pattern(oc, c, [W(1), W(dom.g_coset), W(dom.g_coset_inv)])
        return oc.best_fft(e, W(dom.extended_omega), ek)

    adv = [inputs["advice%d" % j] for j in range(n_adv)]
    table = inputs["table_lagrange"]
    for a in adv:
        commit(a, gl)
    perm_l = []
    for j in range(LK):
        pa, pt = oc.lookup_permute(adv[A + j], table, usable)
        a2, s2 = adv[A + j].copy(), table.copy()
        a2[:usable], s2[:usable] = pa, pt
        perm_l.append((a2, s2))
    for a2, s2 in perm_l:
        commit(a2, gl)
        commit(s2, gl)
    z_l, last_z = [], W(1)
    for si in range(sets):
        cols = list(range(si * chunk, min((si + 1) * chunk, n_adv)))
        z = oc.permutation_product([adv[c] for c in cols], [inputs["sigma_lagrange%d" % c] for c in cols], beta, gamma, W(DELTA),
                                   W(pow(DELTA, cols[0], o.R_MOD)), W(dom.omega), last_z)
        z_l.append(z)
        last_z = z[n - BLINDING].copy()
    zl_l = [oc.lookup_product(adv[A + j], table, perm_l[j][0], perm_l[j][1], beta, gamma) for j in range(LK)]
    for z in z_l + zl_l:
        commit(z, gl)
    rnd = inputs["random_poly"]
    commit(rnd, g)
    inst_c = to_coeff(inputs["instance"])
    adv_c = [to_coeff(a) for a in adv]
    z_c = [to_coeff(z) for z in z_l]
    zl_c = [to_coeff(z) for z in zl_l]
    perm_c = [(to_coeff(a2), to_coeff(s2)) for a2, s2 in perm_l]
    adv_e = [to_ext(c) for c in adv_c]
    inst_e = to_ext(inst_c)
    z_e = [to_ext(c) for c in z_c]
    lk_e = [(to_ext(zc), to_ext(pc[0]), to_ext(pc[1])) for zc, pc in zip(zl_c, perm_c)]
    fixed_e = [inputs["fixed_ext%d" % j] for j in range(A + 1)]
    sigma_e = [inputs["sigma_ext%d" % j] for j in range(n_adv)]
    l0, l_last, l_active = inputs["l0"], inputs["l_last"], inputs["l_active"]

    def tup(gr):
        a = gr.arrays()
        return (a.constants, a.rotations, a.calculations, a.parts, a.n_intermediates)
    ch = np.zeros((0, 4), dtype=np.uint64)
    values = oc.evaluate_graph(tup(evaluator.custom_gates), fixed_e, adv_e, [inst_e], ch, beta, gamma, theta, y, np.zeros((en, 4), dtype=np.uint64), rot_scale)
    values = oc.evaluate_h_permutation(values, rot_scale, z_e, adv_e, sigma_e, chunk, -BLINDING, l0, l_last, l_active, beta, gamma, y, W(DELTA),
                                       W(dom.g_coset), W(dom.extended_omega))
    for gr, (ze, ae, se) in zip(evaluator.lookups, lk_e):
        values = oc.evaluate_h_lookup(tup(gr), fixed_e, adv_e, [inst_e], ch, beta, gamma, theta, y, values, rot_scale, ze, ae, se, l0, l_last, l_active)
    values = scale_pattern(oc, values, [W(t) for t in dom.t_evaluations])
    values = oc.best_fft(values, W(dom.extended_omega_inv), ek)
    values = scale_pattern(oc, values, [W(dom.extended_ifft_divisor), W(dom.extended_ifft_divisor * dom.g_coset_inv), W(dom.extended_ifft_divisor * dom.g_coset)])
    pieces = [np.ascontiguousarray(values[p * n:(p + 1) * n]) for p in range(d - 1)]
    for p in pieces:
        commit(p, g)
    queried = [(c, 4) for c in adv_c] + [(c, 3) for c in z_c] + [(c, 2) for c in zl_c] + [(p, 1) for pc in perm_c for p in pc] + \
              [(p, 1) for p in pieces] + [(rnd, 1)]
    for c, rotations in queried:
        for _ in range(rotations):
            evals.append(oc.fr_eval_polynomial(c, x))
    quot = []
    for npts in (4, 3, 2, 1):
        members = [c for c, r in queried if r == npts]
        if not members:
            continue
        q = lincomb(oc, members, weights[:len(members)])
        for _ in range(npts):
            nxt = np.zeros((n, 4), dtype=np.uint64)
            nxt[:n - 1] = oc.fr_kate_division(q, x)
            q = nxt
        quot.append(q)
    hq = lincomb(oc, quot, weights[:len(quot)])
    commit(hq, g)
    fin = np.zeros((n, 4), dtype=np.uint64)
    fin[:n - 1] = oc.fr_kate_division(hq, v)
    commit(fin, g)
    return np.stack(commitments), np.stack(evals)


def check_pipeline(L, oc, shape, batched=True, seed=0):
    """device pipeline == oracle pipeline, every commitment (affine) and every evaluation; returns the counts"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import proof_pipeline_core as core
    rec = {}
    P = core.Pipeline(L, shape, record=rec, batched=batched, seed=seed)
    try:
        P.run()                      # a second run over the same arena must give the same answers (scratch reuse)
        rec.pop("commitments")
        _, blocks, evals = P.run()
        want_c, want_e = oracle_pipeline(oc, shape, rec["inputs"], P.E)
        got_c = np.stack([oc.g1_to_affine(np.ascontiguousarray(b[:12])) for b in blocks])
        assert got_c.shape == want_c.shape, (got_c.shape, want_c.shape)
        bad = [i for i in range(got_c.shape[0]) if not (got_c[i] == want_c[i]).all()]
        assert not bad, ("commitments differ from the oracle pipeline", bad)
        assert evals.shape == want_e.shape and (evals == want_e).all(), "evaluations differ from the oracle pipeline"
        return dict(P.counts)
    finally:
        P.close()

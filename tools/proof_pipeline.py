#!/usr/bin/env python3
"""
The data-parallel part of ONE create_proof with every column resident on the device, in the order
of halo2_proofs' plonk/prover.rs: witness columns are uploaded once (pageable host arrays), everything
else -- commitments, (i)NTTs, grand products, the quotient evaluation, the evaluations at x and the opening quotients -- runs
through the device-pointer entry points and only 96-byte commitments / 32-byte evaluations come back.

Not included (host work of the prover that stays on the host): witness generation, transcript hashing, blinding-row randomness.  Proving-key columns (fixed,
sigma, l_0 / l_last / l_active cosets) and the SRS tables are resident before the timed region, as after keygen.
The reference's own create_proof cannot be run here; column counts are the estimates of SURVEY.md section 3.3 / appendix.
usage: python tools/proof_pipeline.py [cfg ...]      one JSON line per configuration
"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import halo2_scaffold_b200 as h2
from halo2_scaffold_b200 import _lib, evaluation as ev
from halo2_scaffold_b200.domain import EvaluationDomain, fr_to_words, FR_MODULUS

# A = gate advice columns, LK = lookup-advice columns (one lookup argument each), d = cs.degree()
SHAPES = {
    "halo2_lib_k16": dict(k=16, A=1, LK=0, d=3),
    "linear_regression_k20": dict(k=20, A=2, LK=1, d=4),
    "logistic_regression_k22": dict(k=22, A=6, LK=2, d=4),
}
DELTA = pow(7, 1 << 28, FR_MODULUS)


def main():
    L = h2.load(); L.init_device(0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    for name in (sys.argv[1:] or list(SHAPES)):
        s = SHAPES[name]
        k, A, LK, d = s["k"], s["A"], s["LK"], s["d"]
        n = 1 << k
        dom = EvaluationDomain(d, k, lib=L)
        ek, en, rot_scale = dom.extended_k, 1 << dom.extended_k, 1 << (dom.extended_k - k)
        chunk = d - 2
        n_adv = A + LK
        sets = math.ceil(n_adv / chunk)
        W = lambda v: fr_to_words(v)
        seed = [7000 * k]

        def dcol(rows, kind=0):
            t = torch.empty(rows * 4, dtype=torch.int64, device=dev)
            seed[0] += 1
            L.gen_scalars_dev(0, seed[0], rows, kind, t.data_ptr(), st)
            return t
        # ---- resident before the proof: SRS (two vectors with tables) and proving-key columns --------------------------------
        pts = torch.empty(n * 8, dtype=torch.int64, device=dev)
        L.gen_points_dev(0, 99 + k, n, pts.data_ptr(), st)
        torch.cuda.synchronize()
        hp = pts.cpu().numpy().view(np.uint64).reshape(n, 8)
        del pts
        h_g, h_gl = L.register_bases(hp), L.register_bases(hp[::-1].copy())
        del hp
        fixed_ext = [dcol(en) for _ in range(A + 1)]           # one selector per gate column + the lookup table column
        host_table = L.gen_scalars(450 + k, n, 0)
        table_lagrange = torch.from_numpy(host_table.view(np.int64).reshape(-1)).to(dev)
        usable = n - 6                                          # blinding_factors + 1 = 6 rows at the end
        sigma_lagrange = [dcol(n) for _ in range(n_adv)]
        sigma_ext = [dcol(en) for _ in range(n_adv)]
        l0, l_last, l_active = dcol(en), dcol(en), dcol(en)
        polys = [ev.Product(ev.Fixed(c), ev.Sum(ev.Sum(ev.Advice(c, 0), ev.Product(ev.Advice(c, 1), ev.Advice(c, 2))), ev.Negated(ev.Advice(c, 3))))
                 for c in range(A)]
        E = ev.Evaluator(polys, [([ev.Advice(A + j)], [ev.Fixed(A)]) for j in range(LK)])
        g_gates, g_lk = E.custom_gates.arrays(), [g.arrays() for g in E.lookups]
        sc = L.gen_scalars(5, 8)
        theta, beta, gamma, y, x, v = sc[:6]
        one = W(1)
        zs = np.stack([W(1), W(dom.g_coset), W(dom.g_coset_inv)])
        e2c = np.stack([W(dom.extended_ifft_divisor), W(dom.extended_ifft_divisor * dom.g_coset_inv), W(dom.extended_ifft_divisor * dom.g_coset)])
        tev = np.stack([W(t) for t in dom.t_evaluations])
        # the witness as the host holds it: pageable arrays
        host_adv = [L.gen_scalars(300 + j, n, 1) for j in range(n_adv)]
        for j in range(LK):                                     # lookup-advice columns hold table values (a permutation of the usable rows)
            host_adv[A + j] = host_table.copy()
            host_adv[A + j][:usable] = host_table[:usable][::-1]
        host_inst = L.gen_scalars(400, n, 1)
        weights = L.gen_scalars(900, 64)                         # powers of y / v of the multi-open argument (host scalars)
        blocks = torch.empty(64 * 28, dtype=torch.int64, device=dev)
        evals = torch.empty(256 * 4, dtype=torch.int64, device=dev)
        counts = {"msm": 0, "intt": 0, "coset_ntt": 0, "eval": 0, "kate": 0}

        def commit(d_scalars, handle, rows=n):
            L.msm_dev_registered(0, d_scalars.data_ptr(), handle, 0, rows, blocks.data_ptr() + 224 * (counts["msm"] % 64), st)
            counts["msm"] += 1

        def to_coeff(t):
            c = t.clone()
            L.lagrange_to_coeff_dev(0, c.data_ptr(), k, W(dom.omega_inv), W(dom.ifft_divisor), st)
            counts["intt"] += 1
            return c

        def to_ext(c):
            e = torch.empty(en * 4, dtype=torch.int64, device=dev)
            e[: n * 4] = c
            L.coeff_to_extended_dev(0, e.data_ptr(), k, ek, W(dom.extended_omega), zs, st)
            counts["coset_ntt"] += 1
            return e

        def run():
            for key in counts:
                counts[key] = 0
            phase, t_prev = {}, [time.perf_counter()]

            def mark(label):
                torch.cuda.synchronize()
                t = time.perf_counter()
                phase[label] = round((t - t_prev[0]) * 1e3, 3)
                t_prev[0] = t
            # instance -> coefficient form
            d_inst = torch.empty(n * 4, dtype=torch.int64, device=dev)
            L.h2d_async(0, d_inst.data_ptr(), host_inst, st)
            inst_c = to_coeff(d_inst)
            # advice: upload, commit_lagrange
            adv = []
            for j in range(n_adv):
                t = torch.empty(n * 4, dtype=torch.int64, device=dev)
                L.h2d_async(0, t.data_ptr(), host_adv[j], st)
                commit(t, h_gl)
                adv.append(t)
            mark("advice_upload_commit")
            # lookups: permute_expression_pair on the device (sort + table matching), commit both permuted columns
            perm_l = []
            for j in range(LK):
                a, s_ = adv[A + j].clone(), table_lagrange.clone()                 # rows >= usable: the blinding rows (host randomness)
                L.lookup_permute_dev(0, adv[A + j].data_ptr(), table_lagrange.data_ptr(), usable, a.data_ptr(), s_.data_ptr(), st)
                commit(a, h_gl); commit(s_, h_gl)
                perm_l.append((a, s_))
            mark("lookup_permuted_commit")
            # permutation grand products: z per set, commit, coefficient + extended form
            z_l, last_z = [], one
            for si in range(sets):
                cols = list(range(si * chunk, min((si + 1) * chunk, n_adv)))
                z = torch.empty(n * 4, dtype=torch.int64, device=dev)
                L.permutation_product_dev(0, [adv[c].data_ptr() for c in cols], [sigma_lagrange[c].data_ptr() for c in cols], n, beta, gamma, W(DELTA),
                                          W(pow(DELTA, cols[0], FR_MODULUS)), W(dom.omega), last_z, z.data_ptr(), st)
                commit(z, h_gl)
                z_l.append(z)
                last_z = gamma                                  # stands in for z[n - (blinding_factors + 1)] (a 32-byte read-back)
            # lookup grand products
            zl_l = []
            for j in range(LK):
                z = torch.empty(n * 4, dtype=torch.int64, device=dev)
                L.lookup_product_dev(0, adv[A + j].data_ptr(), table_lagrange.data_ptr(), perm_l[j][0].data_ptr(), perm_l[j][1].data_ptr(), n, beta, gamma,
                                     z.data_ptr(), st)
                commit(z, h_gl)
                zl_l.append(z)
            mark("grand_products_commit")
            # vanishing argument's random polynomial
            rnd = dcol(n)
            commit(rnd, h_g)
            # everything to coefficient form (kept for the evaluations and the opening) and to the extended coset
            adv_c = [to_coeff(t) for t in adv]
            z_c = [to_coeff(t) for t in z_l]
            zl_c = [to_coeff(t) for t in zl_l]
            perm_c = [(to_coeff(a), to_coeff(s_)) for a, s_ in perm_l]
            mark("lagrange_to_coeff")
            adv_e = [to_ext(c) for c in adv_c]
            inst_e = to_ext(inst_c)
            z_e = [to_ext(c) for c in z_c]
            lk_e = [(to_ext(zc), to_ext(pc[0]), to_ext(pc[1])) for zc, pc in zip(zl_c, perm_c)]
            mark("coeff_to_extended")
            # evaluate_h
            values = torch.zeros(en * 4, dtype=torch.int64, device=dev)
            cols = _lib.EvalColumns([t.data_ptr() for t in fixed_ext], [t.data_ptr() for t in adv_e], [inst_e.data_ptr()], np.zeros((0, 4), dtype=np.uint64),
                                    beta, gamma, theta, y)
            L.evaluate_graph_dev(0, g_gates, cols, values.data_ptr(), en, rot_scale, st)
            L.evaluate_h_permutation_dev(0, values.data_ptr(), en, rot_scale, [t.data_ptr() for t in z_e], [t.data_ptr() for t in adv_e],
                                         [t.data_ptr() for t in sigma_ext], chunk, -6, l0.data_ptr(), l_last.data_ptr(), l_active.data_ptr(), beta, gamma, y,
                                         W(DELTA), W(dom.g_coset), W(dom.extended_omega), st)
            for g, (ze, ae, se) in zip(g_lk, lk_e):
                L.evaluate_h_lookup_dev(0, g, cols, values.data_ptr(), en, rot_scale, ze.data_ptr(), ae.data_ptr(), se.data_ptr(), l0.data_ptr(),
                                        l_last.data_ptr(), l_active.data_ptr(), st)
            mark("evaluate_h")
            # h = values / (X^n - 1) -> coefficients -> d - 1 pieces of n -> commit each
            L.fr_scale_dev(0, values.data_ptr(), en, tev, st)
            L.extended_to_coeff_dev(0, values.data_ptr(), ek, W(dom.extended_omega_inv), e2c, st)
            counts["coset_ntt"] += 1
            for piece in range(d - 1):
                commit(values[piece * n * 4:(piece + 1) * n * 4], h_g)
            mark("quotient_commit")
            # evaluations at x (and rotations of x): Horner over every queried polynomial
            queried = [(c, 4) for c in adv_c] + [(c, 3) for c in z_c] + [(c, 2) for c in zl_c] + [(p, 1) for pc in perm_c for p in pc] + \
                      [(values[piece * n * 4:(piece + 1) * n * 4], 1) for piece in range(d - 1)] + [(rnd, 1)]
            for c, rotations in queried:
                for r in range(rotations):
                    L.check(L.L.h2b_fr_eval_polynomial_dev(0, c.data_ptr(), n, x.ctypes.data, evals.data_ptr() + 32 * (counts["eval"] % 256), st))
                    counts["eval"] += 1
            mark("evaluations")
            # SHPLONK: per rotation set a y-weighted sum of its polynomials, divided by (X - point) for every point of the set; the
            # v-weighted sum of the quotients is committed; then the final quotient at u is committed
            rot_sets = {4: [c for c, r in queried if r == 4], 3: [c for c, r in queried if r == 3], 2: [c for c, r in queried if r == 2],
                        1: [c for c, r in queried if r == 1]}
            quot = []
            for npts, members in rot_sets.items():
                if not members:
                    continue
                comb = torch.empty(n * 4, dtype=torch.int64, device=dev)
                L.fr_lincomb_dev(0, [m_.data_ptr() for m_ in members], weights[:len(members)], n, comb.data_ptr(), st)
                q = comb
                for _ in range(npts):
                    nxt = torch.empty(n * 4, dtype=torch.int64, device=dev)
                    L.check(L.L.h2b_fr_kate_division_dev(0, q.data_ptr(), n, x.ctypes.data, nxt.data_ptr(), st))
                    counts["kate"] += 1
                    q = nxt
                quot.append(q)
            hq = torch.empty(n * 4, dtype=torch.int64, device=dev)
            L.fr_lincomb_dev(0, [q.data_ptr() for q in quot], weights[:len(quot)], n, hq.data_ptr(), st)
            commit(hq, h_g)
            fin = torch.empty(n * 4, dtype=torch.int64, device=dev)
            L.check(L.L.h2b_fr_kate_division_dev(0, hq.data_ptr(), n, v.ctypes.data, fin.data_ptr(), st))
            counts["kate"] += 1
            commit(fin, h_g)
            out = blocks.cpu()                                   # the commitments (224 B blocks) come back
            ev_out = evals.cpu()
            mark("multiopen")
            return phase

        run()                                                    # first use: twiddle tables, scratch buffers
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        phase = run()
        total_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"config": name, "k": k, "extended_k": ek, "gate_advice": A, "lookup_advice": LK, "degree": d, "permutation_sets": sets,
                          "device_resident_hot_path_ms": round(total_ms, 2), "phases_ms": phase, "calls": dict(counts),
                          "h2d_bytes": (n_adv + 1) * n * 32, "note": "1 x B200; witness columns uploaded from pageable host arrays; host-side prover work "
                          "(witness generation, transcript, blinding) not included; column counts are estimates"}), flush=True)
        L.unregister_bases(h_g); L.unregister_bases(h_gl)
        del fixed_ext, sigma_ext, sigma_lagrange
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""
tools/proof_pipeline.py across the D GPUs of one process (BASELINE.json configs[4]: "many advice columns; commits sharded across 8
GPUs"): the SRS (with window tables) is resident on every device; column j lives on device j mod D, where it is
uploaded (pinned host arrays, one PCIe link per GPU), committed, brought to coefficient form and to the extended coset; every
permutation set and every lookup argument runs on one device (the set's columns are peer-copied there); evaluate_h is sharded by rows: every device receives its
row slice (+ halo) of each extended column over NVLink; the quotient pieces, the evaluations at x and the per-column work of the opening
go back to one device per polynomial.  Commitments that sit alone on a dependency chain (the quotient pieces, the opening) are split by point range over all devices and their partial sums folded on device 0.  No collective: only peer copies.
The wall time is taken on the host around the whole sequence with every device synchronised at the end.
usage: python tools/proof_pipeline_multi.py [cfg ...]        (uses every visible GPU)
"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import halo2_scaffold_b200 as h2
from halo2_scaffold_b200 import _lib, evaluation as ev
from halo2_scaffold_b200.domain import EvaluationDomain, fr_to_words, FR_MODULUS

SHAPES = {
    "linear_regression_k20": dict(k=20, A=2, LK=1, d=4),
    "logistic_regression_k22": dict(k=22, A=6, LK=2, d=4),
    "logistic_regression_k22_wide": dict(k=22, A=14, LK=2, d=4),
}
DELTA = pow(7, 1 << 28, FR_MODULUS)


def main():
    L = h2.load(); L.init(0)
    D = L.device_count()
    devs = [torch.device("cuda", i) for i in range(D)]
    st = [torch.cuda.current_stream(i).cuda_stream for i in range(D)]
    W = fr_to_words
    for name in (sys.argv[1:] or ["logistic_regression_k22"]):
        s = SHAPES[name]
        k, A, LK, d = s["k"], s["A"], s["LK"], s["d"]
        n = 1 << k
        dom = EvaluationDomain(d, k, lib=L)
        ek, en, rot_scale = dom.extended_k, 1 << dom.extended_k, 1 << (dom.extended_k - k)
        chunk, n_adv, usable = d - 2, A + LK, n - 6
        sets = math.ceil(n_adv / chunk)
        seed = [7000 * k]

        def dcol(dv, rows):
            with torch.cuda.device(dv):
                t = torch.empty(rows * 4, dtype=torch.int64, device=devs[dv])
                seed[0] += 1
                L.gen_scalars_dev(dv, seed[0], rows, 0, t.data_ptr(), st[dv])
            return t
        # ---- resident before the proof ---------------------------------------------------------------------------------------
        with torch.cuda.device(0):
            pts = torch.empty(n * 8, dtype=torch.int64, device=devs[0])
            L.gen_points_dev(0, 99 + k, n, pts.data_ptr(), st[0])
            torch.cuda.synchronize(0)
            hp = pts.cpu().numpy().view(np.uint64).reshape(n, 8)
            del pts
        h_g, h_gl = L.register_bases(hp), L.register_bases(hp[::-1].copy())          # every device gets the points and its own tables
        del hp
        fixed_ext = [dcol(0, en) for _ in range(A + 1)]
        sigma_ext = [dcol(0, en) for _ in range(n_adv)]
        l0, l_last, l_active = dcol(0, en), dcol(0, en), dcol(0, en)
        halo = 6 * rot_scale if D > 1 else 0                       # |last_rotation| * rot_scale covers every rotated read of evaluate_h
        rows_d = en // D

        def slice_to(t, dv):
            """rows [dv * rows_d - halo, (dv + 1) * rows_d + halo) of an extended column (wrap-around included) as a tensor on device dv"""
            v = t.view(-1, 4)
            lo, hi = dv * rows_d - halo, (dv + 1) * rows_d + halo
            with torch.cuda.device(dv):
                if D == 1:
                    return t
                if lo < 0:
                    return torch.cat([v[lo + en:].to(devs[dv], non_blocking=True), v[:hi].to(devs[dv], non_blocking=True)]).view(-1)
                if hi > en:
                    return torch.cat([v[lo:].to(devs[dv], non_blocking=True), v[:hi - en].to(devs[dv], non_blocking=True)]).view(-1)
                return v[lo:hi].to(devs[dv], non_blocking=True).contiguous().view(-1)
        # proving-key columns of evaluate_h: every device keeps its row slice (with halo) from keygen on
        pk_slices = [dict(fixed=[slice_to(t, dv) for t in fixed_ext], sigma=[slice_to(t, dv) for t in sigma_ext], l0=slice_to(l0, dv),
                          l_last=slice_to(l_last, dv), l_active=slice_to(l_active, dv)) for dv in range(D)]
        host_table = L.gen_scalars(450 + k, n, 0)
        table_lagrange = [torch.from_numpy(host_table.view(np.int64).reshape(-1)).to(devs[i]) for i in range(D)]      # proving-key data, replicated
        set_dev = [si % D for si in range(sets)]
        sigma_lagrange = [dcol(set_dev[c // chunk], n) for c in range(n_adv)]
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
        weights = L.gen_scalars(900, 64)
        # the witness in PINNED host memory (one DMA per column, each GPU over its own PCIe link)
        host_adv = []
        for j in range(n_adv):
            a = L.gen_scalars(300 + j, n, 1) if j < A else host_table.copy()
            if j >= A:
                a[:usable] = host_table[:usable][::-1]
            host_adv.append(torch.from_numpy(a.view(np.int64).reshape(-1)).pin_memory())
        host_inst = torch.from_numpy(L.gen_scalars(400, n, 1).view(np.int64).reshape(-1)).pin_memory()
        blocks = [torch.empty(64 * 28, dtype=torch.int64, device=devs[i]) for i in range(D)]
        evals = [torch.empty(256 * 4, dtype=torch.int64, device=devs[i]) for i in range(D)]
        n_msm = [0] * D
        n_eval = [0] * D
        torch.cuda.synchronize()

        def on(t):
            return t.device.index

        def commit(t, handle):
            dv = on(t)
            with torch.cuda.device(dv):
                L.msm_dev_registered(dv, t.data_ptr(), handle, 0, n, blocks[dv].data_ptr() + 224 * (n_msm[dv] % 64), st[dv])
            n_msm[dv] += 1

        folded = [torch.empty(12, dtype=torch.int64, device=devs[0]) for _ in range(16)]
        n_sharded = [0]

        def commit_sharded(t, handle):
            """ONE commitment split by point range over all devices: device dv gets scalars [dv n / D, (dv + 1) n / D) over NVLink, runs a
            partial MSM over the matching rows of its resident tables, and the D partial sums (224-byte blocks) are folded on device 0.  For the
            commitments that sit alone on a dependency chain (the quotient pieces, the opening)."""
            if D == 1:
                return commit(t, handle)
            v = t.view(-1, 4)
            per = n // D
            parts = []
            for dv in range(D):
                with torch.cuda.device(dv):
                    sl = v[dv * per:(dv + 1) * per]
                    sl = sl if on(t) == dv else sl.to(devs[dv], non_blocking=True)
                    blk = torch.empty(28, dtype=torch.int64, device=devs[dv])
                    L.msm_dev_registered(dv, sl.data_ptr(), handle, dv * per, per, blk.data_ptr(), st[dv])
                    n_msm[dv] += 1
                    parts.append((sl, blk))
            with torch.cuda.device(0):
                allb = torch.cat([move(b, 0) for _, b in parts])
                L.msm_fold_partials_dev(0, allb.data_ptr(), D, folded[n_sharded[0] % 16].data_ptr(), st[0])
            n_sharded[0] += 1
            return parts, allb

        P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47

        def affine_of_jacobian(words):
            """x | y | z Montgomery limbs (12 x u64) -> canonical affine (x, y) or None; host big-int arithmetic, for the self-check only"""
            w = [int(v) & 0xFFFFFFFFFFFFFFFF for v in words]
            val = lambda i: sum(w[4 * i + j] << (64 * j) for j in range(4)) * pow(1 << 256, -1, P_MOD) % P_MOD
            X, Y, Z = val(0), val(1), val(2)
            if Z == 0:
                return None
            zi = pow(Z, -1, P_MOD)
            return (X * zi * zi % P_MOD, Y * zi * zi * zi % P_MOD)

        def self_check():
            """a sharded commitment equals the same commitment on one device"""
            t = dcol(0, n)
            commit(t, h_g)
            commit_sharded(t, h_g)
            for i in range(D):
                torch.cuda.synchronize(i)
            one_dev = blocks[0][((n_msm[0] - 2 if D > 1 else n_msm[0] - 2) % 64) * 28:][:12].cpu().numpy()
            shard = folded[(n_sharded[0] - 1) % 16].cpu().numpy() if D > 1 else blocks[0][((n_msm[0] - 1) % 64) * 28:][:12].cpu().numpy()
            assert affine_of_jacobian(one_dev) == affine_of_jacobian(shard), "sharded commitment differs"

        def to_coeff(t):
            dv = on(t)
            with torch.cuda.device(dv):
                c = t.clone()
                L.lagrange_to_coeff_dev(dv, c.data_ptr(), k, W(dom.omega_inv), W(dom.ifft_divisor), st[dv])
            return c

        def to_ext(c):
            dv = on(c)
            with torch.cuda.device(dv):
                e = torch.empty(en * 4, dtype=torch.int64, device=devs[dv])
                e[: n * 4] = c
                L.coeff_to_extended_dev(dv, e.data_ptr(), k, ek, W(dom.extended_omega), zs, st[dv])
            return e

        def move(t, dv):
            return t if on(t) == dv else t.to(devs[dv], non_blocking=True)

        def run():
            for i in range(D):
                n_msm[i] = n_eval[i] = 0
            keep = []                                           # slices and partial blocks of the sharded commitments stay alive until the end
            phase, t_prev = {}, [time.perf_counter()]

            def mark(label):
                for i in range(D):
                    torch.cuda.synchronize(i)
                t = time.perf_counter()
                phase[label] = round((t - t_prev[0]) * 1e3, 3)
                t_prev[0] = t
            # columns: upload, commit_lagrange, coefficient form, extended coset -- column j on device j mod D
            adv, adv_c, adv_e = [], [], []
            for j in range(n_adv):
                dv = j % D
                with torch.cuda.device(dv):
                    t = torch.empty(n * 4, dtype=torch.int64, device=devs[dv])
                    t.copy_(host_adv[j], non_blocking=True)
                commit(t, h_gl)
                adv.append(t)
            with torch.cuda.device(0):
                d_inst = torch.empty(n * 4, dtype=torch.int64, device=devs[0])
                d_inst.copy_(host_inst, non_blocking=True)
            inst_c = to_coeff(d_inst)
            for t in adv:
                adv_c.append(to_coeff(t))
            for c in adv_c:
                adv_e.append(to_ext(c))
            inst_e = to_ext(inst_c)
            # permutation sets: the set's columns are copied to the set's device; z, commit, coefficient + extended form
            z_l, last_z = [], one
            for si in range(sets):
                dv = set_dev[si]
                cols = list(range(si * chunk, min((si + 1) * chunk, n_adv)))
                with torch.cuda.device(dv):
                    local = [move(adv[c], dv) for c in cols]
                    z = torch.empty(n * 4, dtype=torch.int64, device=devs[dv])
                    L.permutation_product_dev(dv, [t.data_ptr() for t in local], [sigma_lagrange[c].data_ptr() for c in cols], n, beta, gamma, W(DELTA),
                                              W(pow(DELTA, cols[0], FR_MODULUS)), W(dom.omega), last_z, z.data_ptr(), st[dv])
                commit(z, h_gl)
                z_l.append(z)
                last_z = gamma                                  # stands in for z[n - (blinding_factors + 1)] of the previous set (a 32-byte read-back)
            z_c = [to_coeff(t) for t in z_l]
            z_e = [to_ext(c) for c in z_c]
            mark("columns_and_permutation")
            # lookups on the column's device; the table-membership verdict of each permutation is read at the end (asynchronous variant)
            perm_l, zl_l, status = [], [], []
            busy = {(A + j) % D for j in range(LK)} | {0}
            idle = [i for i in range(D) if i not in busy] or list(range(D))
            spare = [idle[i % len(idle)] for i in range(2 * LK + 1)]     # where the permuted columns of each lookup are committed and transformed
            for j in range(LK):
                dv = on(adv[A + j])
                with torch.cuda.device(dv):
                    a, s_ = adv[A + j].clone(), table_lagrange[dv].clone()
                    status.append(torch.zeros(1, dtype=torch.int32, device=devs[dv]))
                    L.lookup_permute_async_dev(dv, adv[A + j].data_ptr(), table_lagrange[dv].data_ptr(), usable, a.data_ptr(), s_.data_ptr(),
                                               status[-1].data_ptr(), st[dv])
                    z = torch.empty(n * 4, dtype=torch.int64, device=devs[dv])
                    L.lookup_product_dev(dv, adv[A + j].data_ptr(), table_lagrange[dv].data_ptr(), a.data_ptr(), s_.data_ptr(), n, beta, gamma, z.data_ptr(), st[dv])
                commit(z, h_gl)                                    # (sharding this one over busy devices measured slower: 54 vs 42 ms for the phase)
                # the two permuted columns leave for other devices: only z stays on the lookup's critical path
                with torch.cuda.device(spare[2 * j]):
                    a2 = move(a, spare[2 * j])
                with torch.cuda.device(spare[2 * j + 1]):
                    s2 = move(s_, spare[2 * j + 1])
                commit(a2, h_gl); commit(s2, h_gl)
                perm_l.append((a2, s2)); zl_l.append(z)
            zl_c = [to_coeff(t) for t in zl_l]
            perm_c = [(to_coeff(a), to_coeff(s_)) for a, s_ in perm_l]
            lk_e = [(to_ext(zc), to_ext(pc[0]), to_ext(pc[1])) for zc, pc in zip(zl_c, perm_c)]
            with torch.cuda.device(spare[2 * LK]):
                rnd = dcol(spare[2 * LK], n)
            commit(rnd, h_g)
            mark("lookups")
            # evaluate_h sharded by rows: device dv evaluates rows [dv * en / D, (dv + 1) * en / D); it receives that slice (+ halo) of every
            # witness-dependent extended column over NVLink and already holds its slice of the proving-key columns
            # all slice exchanges are queued before any evaluation kernel: a peer copy orders the destination stream after everything already
            # queued on the source stream, so an evaluation queued on device 0 first would hold back every copy out of device 0
            slices = []
            for dv in range(D):
                with torch.cuda.device(dv):
                    slices.append(([slice_to(t, dv) for t in adv_e], slice_to(inst_e, dv), [slice_to(t, dv) for t in z_e],
                                   [tuple(slice_to(t, dv) for t in tr) for tr in lk_e]))
            mark("exchange_row_slices")
            shard_vals = []
            for dv in range(D):
                with torch.cuda.device(dv):
                    s_adv, s_inst, s_z, s_lk = slices[dv]
                    vals = torch.zeros(rows_d * 4, dtype=torch.int64, device=devs[dv])
                    pk = pk_slices[dv]
                    shard = (dv * rows_d, rows_d, halo)
                    cols = _lib.EvalColumns([t.data_ptr() for t in pk["fixed"]], [t.data_ptr() for t in s_adv], [s_inst.data_ptr()],
                                            np.zeros((0, 4), dtype=np.uint64), beta, gamma, theta, y)
                    L.evaluate_graph_dev(dv, g_gates, cols, vals.data_ptr(), en, rot_scale, st[dv], shard=shard)
                    L.evaluate_h_permutation_dev(dv, vals.data_ptr(), en, rot_scale, [t.data_ptr() for t in s_z], [t.data_ptr() for t in s_adv],
                                                 [t.data_ptr() for t in pk["sigma"]], chunk, -6, pk["l0"].data_ptr(), pk["l_last"].data_ptr(),
                                                 pk["l_active"].data_ptr(), beta, gamma, y, W(DELTA), W(dom.g_coset), W(dom.extended_omega), st[dv], shard=shard)
                    for g, (ze, ae, se) in zip(g_lk, s_lk):
                        L.evaluate_h_lookup_dev(dv, g, cols, vals.data_ptr(), en, rot_scale, ze.data_ptr(), ae.data_ptr(), se.data_ptr(), pk["l0"].data_ptr(),
                                                pk["l_last"].data_ptr(), pk["l_active"].data_ptr(), st[dv], shard=shard)
                    shard_vals.append((vals,))
            mark("evaluate_h_row_sharded")
            with torch.cuda.device(0):
                values = torch.cat([move(v[0], 0) for v in shard_vals])
                L.fr_scale_dev(0, values.data_ptr(), en, tev, st[0])
                L.extended_to_coeff_dev(0, values.data_ptr(), ek, W(dom.extended_omega_inv), e2c, st[0])
                pieces = [values[p * n * 4:(p + 1) * n * 4] for p in range(d - 1)]
            mark("gather_values_and_quotient")
            # quotient pieces: one device each
            h_pieces = []
            for p, t in enumerate(pieces):
                dv = p % D
                with torch.cuda.device(dv):
                    t2 = move(t, dv) if dv else t
                h_pieces.append(t2)
                keep.append(commit_sharded(t2, h_g))
            # evaluations at x: on the device that holds the coefficients
            queried = [(c, 4) for c in adv_c] + [(c, 3) for c in z_c] + [(c, 2) for c in zl_c] + [(p_, 1) for pc in perm_c for p_ in pc] + \
                      [(t, 1) for t in h_pieces] + [(rnd, 1)]
            for c, rotations in queried:
                dv = on(c)
                with torch.cuda.device(dv):
                    for r in range(rotations):
                        L.check(L.L.h2b_fr_eval_polynomial_dev(dv, c.data_ptr(), n, x.ctypes.data, evals[dv].data_ptr() + 32 * (n_eval[dv] % 256), st[dv]))
                        n_eval[dv] += 1
            mark("quotient_commit_and_evaluations")
            # SHPLONK: each rotation set on its own device (its polynomials are copied there), the final quotient on device 0
            rot_sets = [(npts, [c for c, r in queried if r == npts]) for npts in (4, 3, 2, 1)]
            quot = []
            for qi, (npts, members) in enumerate(rot_sets):
                if not members:
                    continue
                dv = qi % D
                with torch.cuda.device(dv):
                    local = [move(m_, dv) for m_ in members]
                    comb = torch.empty(n * 4, dtype=torch.int64, device=devs[dv])
                    L.fr_lincomb_dev(dv, [m_.data_ptr() for m_ in local], weights[:len(local)], n, comb.data_ptr(), st[dv])
                    q = comb
                    for _ in range(npts):
                        nxt = torch.empty(n * 4, dtype=torch.int64, device=devs[dv])
                        L.check(L.L.h2b_fr_kate_division_dev(dv, q.data_ptr(), n, x.ctypes.data, nxt.data_ptr(), st[dv]))
                        q = nxt
                quot.append(q)
            with torch.cuda.device(0):
                quot0 = [move(q, 0) for q in quot]
                hq = torch.empty(n * 4, dtype=torch.int64, device=devs[0])
                L.fr_lincomb_dev(0, [q.data_ptr() for q in quot0], weights[:len(quot0)], n, hq.data_ptr(), st[0])
                fin = torch.empty(n * 4, dtype=torch.int64, device=devs[0])
                L.check(L.L.h2b_fr_kate_division_dev(0, hq.data_ptr(), n, v.ctypes.data, fin.data_ptr(), st[0]))
                keep.append(commit_sharded(hq, h_g))
            with torch.cuda.device(D - 1):                               # the two opening commitments on two devices
                fin2 = move(fin, D - 1)
            keep.append(commit_sharded(fin2, h_g))
            out = [b.cpu() for b in blocks] + [e.cpu() for e in evals]
            assert all(int(w.cpu()[0]) == 0 for w in status), "a lookup input value is not in the table"
            mark("multiopen")
            return phase

        self_check()
        run()
        for i in range(D):
            torch.cuda.synchronize(i)
        t0 = time.perf_counter()
        phase = run()
        total_ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"config": name, "devices": D, "k": k, "extended_k": ek, "gate_advice": A, "lookup_advice": LK, "permutation_sets": sets,
                          "device_resident_hot_path_ms": round(total_ms, 2), "phases_ms": phase, "msm_per_device": list(n_msm),
                          "note": "one process, %d x B200; witness in pinned host memory; columns round-robin over the devices, evaluate_h sharded by rows over the devices "
                                  "(slices + halo exchanged over NVLink); host-side prover work not included; column counts are estimates" % D}), flush=True)
        L.unregister_bases(h_g); L.unregister_bases(h_gl)
        del fixed_ext, sigma_ext, sigma_lagrange, table_lagrange
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

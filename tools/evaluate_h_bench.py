#!/usr/bin/env python3
"""Quotient evaluation (evaluate_h) on device-resident extended-coset columns for a halo2-lib-shaped circuit
(what scaffold::run builds: per advice column one gate q * (a + b * c - d) on rotations 0..3 of that column, selector-less
lookups of the lookup-advice columns into one table column, a permutation over every advice column):
ms per loop, micro-operations and multiplications per row, achieved Fr multiplications/s against the measured 68.06 G/s
of this B200 (profiles/r01_field_primitives.jsonl) and the column traffic.  The CPU restatement (one thread, upstream
is rayon) is timed on a bounded sample of rows.
usage: python tools/evaluate_h_bench.py [--advice A] [--lookup-advice L] [--cpu-rows R] [k:extended_k ...]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import halo2_scaffold_b200 as h2
from halo2_scaffold_b200 import _lib, evaluation as ev
from halo2_scaffold_b200.domain import fr_to_words

FR_MUL_PEAK = 68.06e9


def halo2_lib_shape(n_advice, n_lookup):
    polys = []
    for c in range(n_advice):            # FlexGateConfig: q * (a + b * c - d), a..d = rotations 0..3 of one advice column
        a, b, cc, d = (ev.Advice(c, r) for r in range(4))
        polys.append(ev.Product(ev.Fixed(c), ev.Sum(ev.Sum(a, ev.Product(b, cc)), ev.Negated(d))))
    lookups = [([ev.Advice(n_advice + j)], [ev.Fixed(n_advice)]) for j in range(n_lookup)]
    return polys, lookups


def count_muls(graph):
    """Fr multiplications per row of a GraphEvaluator as upstream walks it (Mul, Square, Horner steps)"""
    n = 0
    for (op, a, b, ps), _ in graph.calculations:
        n += 1 if op in (ev.MUL, ev.SQUARE) else len(ps) if op == ev.HORNER else 0
    return n


def measure(L, k, ek, A=8, LK=2, reps=5, cpu_rows=1 << 15):
    """evaluate_h on device-resident columns of 2^ek rows (domain 2^k): dict of timings; parity of a sample against the oracle"""
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    polys, lookups = halo2_lib_shape(A, LK)
    E = ev.Evaluator(polys, lookups)
    sc = L.gen_scalars(99, 8)
    beta, gamma, theta, y, delta, zeta = sc[:6]
    none = np.zeros((0, 4), dtype=np.uint64)

    class _A:
        pass
    args = _A()
    args.reps, args.cpu_rows = reps, cpu_rows
    size, rot_scale = 1 << ek, 1 << (ek - k)
    seed = [1000 * ek]

    def col():
        t = torch.empty(size * 4, dtype=torch.int64, device=dev)
        seed[0] += 1
        L.gen_scalars_dev(0, seed[0], size, 0, t.data_ptr(), st)
        return t
    fixed = [col() for _ in range(A + 1)]
    advice = [col() for _ in range(A + LK)]
    l0, l_last, l_active = col(), col(), col()
    values = col()
    perm_cols = advice                                      # every advice column takes part in the permutation
    chunk_len = 2                                           # cs.degree() = 4
    n_sets = (len(perm_cols) + chunk_len - 1) // chunk_len
    z = [col() for _ in range(n_sets)]
    sigma = [col() for _ in perm_cols]
    lk_cols = [(col(), col(), col()) for _ in range(LK)]
    cols = _lib.EvalColumns([t.data_ptr() for t in fixed], [t.data_ptr() for t in advice], [], none, beta, gamma, theta, y)
    omega = fr_to_words(pow(7, (ev.FR_MODULUS - 1) >> ek, ev.FR_MODULUS))
    g_gates = E.custom_gates.arrays()
    g_lk = [g.arrays() for g in E.lookups]

    def gates():
        L.evaluate_graph_dev(0, g_gates, cols, values.data_ptr(), size, rot_scale, st)

    def perm():
        L.evaluate_h_permutation_dev(0, values.data_ptr(), size, rot_scale, [t.data_ptr() for t in z], [t.data_ptr() for t in perm_cols],
                                     [t.data_ptr() for t in sigma], chunk_len, -6, l0.data_ptr(), l_last.data_ptr(), l_active.data_ptr(), beta, gamma, y,
                                     delta, zeta, omega, st)

    def lks():
        for g, (p, a, s) in zip(g_lk, lk_cols):
            L.evaluate_h_lookup_dev(0, g, cols, values.data_ptr(), size, rot_scale, p.data_ptr(), a.data_ptr(), s.data_ptr(), l0.data_ptr(),
                                    l_last.data_ptr(), l_active.data_ptr(), st)
    out = {"k": k, "extended_k": ek, "advice": A, "lookup_advice": LK, "permutation_sets": n_sets,
           "device_columns": len(fixed) + len(advice) + 4 + n_sets + len(sigma) + 3 * LK, "column_bytes": size * 32}
    gates()
    out["custom_gates_slots"], out["custom_gates_micro_ops"] = L.evaluate_graph_info()
    out["custom_gates_calculations"] = len(E.custom_gates.calculations)
    # multiplications per row, as upstream's walk performs them
    m_gates = count_muls(E.custom_gates)
    m_perm = 4 + 2 * (n_sets - 1) + 1 + n_sets * 2 + 4 * len(perm_cols)        # y-steps and terms, delta_start, per column: 2 products + beta*s + delta step
    m_lk = LK * (count_muls(E.lookups[0]) + 16)
    muls = {"custom_gates": m_gates, "permutation": m_perm, "lookups": m_lk}
    reads = {"custom_gates": 5 * A + 1, "permutation": 4 + 2 * n_sets + (n_sets - 1) + 2 * len(perm_cols), "lookups": LK * (10 + 3)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for name, fn in (("custom_gates", gates), ("permutation", perm), ("lookups", lks)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        total += ms
        out[name + "_ms"] = round(ms, 4)
        out[name + "_fr_mul_per_row"] = muls[name]
        out[name + "_fr_mul_frac_of_peak"] = round(muls[name] * size / (ms * 1e-3) / FR_MUL_PEAK, 3)
        out[name + "_column_gb_s"] = round((reads[name] + 1) * size * 32 / ms / 1e6, 1)
    out["evaluate_h_ms"] = round(total, 4)
    out["rows_per_s"] = size / total * 1e3
    # parity on a sample + the CPU restatement's time (one thread) on the first R rows of the same columns... the rotated
    # reads wrap inside the sample, so the sample is evaluated as its own small domain on both sides
    R = min(args.cpu_rows, size)
    if R:
        import oracle_c as oc
        hf = [oc.random_fr(5000 + j, R) for j in range(A + 1)]
        ha = [oc.random_fr(6000 + j, R) for j in range(A + LK)]
        hl = [oc.random_fr(7000 + j, R) for j in range(3)]
        hz = [oc.random_fr(7100 + j, R) for j in range(n_sets)]
        hs = [oc.random_fr(7200 + j, R) for j in range(len(perm_cols))]
        hk = [[oc.random_fr(7300 + 3 * j + i, R) for i in range(3)] for j in range(LK)]
        om = fr_to_words(pow(7, (ev.FR_MODULUS - 1) >> (R.bit_length() - 1), ev.FR_MODULUS))

        def tup(g):
            return (g.constants, g.rotations, g.calculations, g.parts, g.n_intermediates)
        t0 = time.perf_counter()
        want = oc.evaluate_graph(tup(g_gates), hf, ha, [], none, beta, gamma, theta, y, np.zeros((R, 4), dtype=np.uint64), rot_scale)
        want = oc.evaluate_h_permutation(want, rot_scale, hz, ha, hs, chunk_len, -6, *hl, beta, gamma, y, delta, zeta, om)
        for g, (p, a, s) in zip(g_lk, hk):
            want = oc.evaluate_h_lookup(tup(g), hf, ha, [], none, beta, gamma, theta, y, want, rot_scale, p, a, s, *hl)
        cpu_s = time.perf_counter() - t0
        got = E.evaluate_h(size=R, rot_scale=rot_scale, fixed=hf, advice=ha, instance=[], challenges=none, y=y, beta=beta, gamma=gamma, theta=theta,
                           l0=hl[0], l_last=hl[1], l_active_row=hl[2],
                           permutation=dict(product_cosets=hz, columns=[("advice", j) for j in range(len(ha))], cosets=hs, chunk_len=chunk_len,
                                            last_rotation=-6, delta=delta, zeta=zeta, extended_omega=om),
                           lookups=[dict(product_coset=p, permuted_input_coset=a, permuted_table_coset=s) for p, a, s in hk], lib=L)
        out["parity_rows"] = R
        out["parity"] = bool((got == want).all())
        out["cpu_restatement_rows_per_s_1_thread"] = R / cpu_s
    del fixed, advice, z, sigma, lk_cols, values
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--advice", type=int, default=8)
    ap.add_argument("--lookup-advice", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=1 << 15)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("shapes", nargs="*", default=["16:18", "20:22"])
    args = ap.parse_args()
    L = h2.load(); L.init_device(0)
    for shape in args.shapes:
        k, ek = (int(v) for v in shape.split(":"))
        print(json.dumps(measure(L, k, ek, args.advice, args.lookup_advice, args.reps, args.cpu_rows)), flush=True)


if __name__ == "__main__":
    main()

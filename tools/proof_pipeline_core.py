"""
The data-parallel part of ONE create_proof with every column resident on the device, in the order of
halo2_proofs' plonk/prover.rs [UP]: witness columns are uploaded once, everything else -- commitments,
(i)NTTs, the lookup permutation, grand products, the quotient evaluation, the evaluations at x and the opening quotients -- runs
through the device-pointer entry points of include/h2b200.h and only 224-byte commitment blocks / 32-byte evaluations come back.

This is a TOOL (benchmark + test harness), not prover orchestration: witness generation, the transcript, blinding rows and the
real challenge derivation stay with the caller (the Rust prover); here challenges are fixed pseudo-random field elements and the
column counts are the estimates of SURVEY.md section 3.3.  It uses nothing but `Lib` (no torch), so the same code runs on the
CUDA library and -- at small k -- on the kernel-logic emulator, and with `record` set it hands every input and every output to
tests/pipeline_oracle.py, which recomputes the whole flow with the CPU oracle.

`batched=True` issues the independent commitments / conversions of one prover phase as ONE batched call each
(h2b_msm_bn254_g1_dev_batch_registered, h2b_lagrange_to_coeff_dev_batch, h2b_coeff_to_extended_dev_batch).
"""
import math
import time

import numpy as np

from halo2_scaffold_b200 import _lib, evaluation as ev
from halo2_scaffold_b200.domain import EvaluationDomain, fr_to_words, FR_MODULUS

# A = gate advice columns, LK = lookup-advice columns (one lookup argument each), d = cs.degree()
SHAPES = {
    "halo2_lib_k16": dict(k=16, A=1, LK=0, d=3),
    "linear_regression_k20": dict(k=20, A=2, LK=1, d=4),
    "logistic_regression_k22": dict(k=22, A=6, LK=2, d=4),
}
DELTA = pow(7, 1 << 28, FR_MODULUS)
BLINDING = 6            # blinding_factors + 1 rows at the end of every column


class Arena:
    """bump allocator over one device allocation (a proof allocates dozens of columns; cudaMalloc per column would dominate at k = 16)"""

    def __init__(self, L, device, nbytes):
        self.L, self.device, self.cap, self.off = L, device, nbytes, 0
        self.base = L.dev_alloc(device, nbytes)

    def alloc(self, nbytes):
        p = self.base + self.off
        self.off += (nbytes + 255) // 256 * 256
        if self.off > self.cap:
            raise MemoryError("proof pipeline arena exhausted (%d of %d bytes)" % (self.off, self.cap))
        return p

    def free(self):
        self.L.dev_free(self.device, self.base)


class Pipeline:
    def __init__(self, L, shape, device=0, record=None, batched=True, seed=0):
        self.L, self.dev, self.rec, self.batched = L, device, record, batched
        s = dict(shape)
        self.k, self.A, self.LK, self.d = s["k"], s["A"], s["LK"], s["d"]
        k, A, LK, d = self.k, self.A, self.LK, self.d
        self.n = n = 1 << k
        self.dom = dom = EvaluationDomain(d, k, lib=L)
        self.ek, self.en = dom.extended_k, 1 << dom.extended_k
        self.rot_scale = 1 << (dom.extended_k - k)
        self.chunk = d - 2
        self.n_adv = A + LK
        self.sets = math.ceil(self.n_adv / self.chunk)
        self.usable = n - BLINDING
        self.seed = 7000 * k + 100000 * seed
        en = self.en
        # one arena for the proving key, one (reset per proof) for the proof's own columns
        n_ext_cols_pk = (A + 1) + self.n_adv + 3
        self.pk_arena = Arena(L, device, n_ext_cols_pk * en * 32 + (self.n_adv + 1) * n * 32 + (1 << 20))
        proof_cols_n = 3 * self.n_adv + 4 * self.sets + 12 * LK + 16 + 3 * d
        proof_cols_en = self.n_adv + 1 + self.sets + 3 * LK + 2
        self.arena = Arena(L, device, proof_cols_n * n * 32 + proof_cols_en * en * 32 + (4 << 20))
        W = fr_to_words
        self.W = W
        # ---- resident before the proof: SRS (two vectors with window tables) and the proving-key columns ---------------------
        pts = L.gen_points(99 + k + seed, n, device)
        self.g, self.g_lagrange = pts, np.ascontiguousarray(pts[::-1])
        self.h_g, self.h_gl = L.register_bases(self.g), L.register_bases(self.g_lagrange)
        self._input("g", self.g)
        self._input("g_lagrange", self.g_lagrange)
        self.fixed_ext = [self._pk_col("fixed_ext%d" % j, en) for j in range(A + 1)]      # one selector per gate column + the lookup table column
        self.host_table = L.gen_scalars(450 + k + seed, n, 0)
        self.table_lagrange = self.pk_arena.alloc(n * 32)
        L.h2d(device, self.table_lagrange, self.host_table)
        self._input("table_lagrange", self.host_table)
        self.sigma_lagrange = [self._pk_col("sigma_lagrange%d" % j, n) for j in range(self.n_adv)]
        self.sigma_ext = [self._pk_col("sigma_ext%d" % j, en) for j in range(self.n_adv)]
        self.l0, self.l_last, self.l_active = (self._pk_col(nm, en) for nm in ("l0", "l_last", "l_active"))
        polys = [ev.Product(ev.Fixed(c), ev.Sum(ev.Sum(ev.Advice(c, 0), ev.Product(ev.Advice(c, 1), ev.Advice(c, 2))), ev.Negated(ev.Advice(c, 3))))
                 for c in range(A)]
        self.E = ev.Evaluator(polys, [([ev.Advice(A + j)], [ev.Fixed(A)]) for j in range(LK)])
        self.g_gates, self.g_lk = self.E.custom_gates.arrays(), [g.arrays() for g in self.E.lookups]
        sc = L.gen_scalars(5 + seed, 8)
        self.theta, self.beta, self.gamma, self.y, self.x, self.v = sc[:6]
        self._input("challenges", sc)
        self.one = W(1)
        self.zs = np.stack([W(1), W(dom.g_coset), W(dom.g_coset_inv)])
        self.e2c = np.stack([W(dom.extended_ifft_divisor), W(dom.extended_ifft_divisor * dom.g_coset_inv), W(dom.extended_ifft_divisor * dom.g_coset)])
        self.tev = np.stack([W(t) for t in dom.t_evaluations])
        # the witness as the host holds it: pageable arrays
        self.host_adv = [L.gen_scalars(300 + j + seed, n, 1) for j in range(self.n_adv)]
        for j in range(LK):     # lookup-advice columns hold table values (a permutation of the usable rows, with repeats)
            col = self.host_table.copy()
            col[:self.usable] = self.host_table[:self.usable][::-1]
            col[1:self.usable:7] = self.host_table[5]
            self.host_adv[A + j] = col
        self.host_inst = L.gen_scalars(400 + seed, n, 1)
        self.host_rnd = L.gen_scalars(401 + seed, n, 0)       # the vanishing argument's random polynomial (host randomness)
        self.weights = L.gen_scalars(900 + seed, 64)           # powers of y / v of the multi-open argument (host scalars)
        for j, c in enumerate(self.host_adv):
            self._input("advice%d" % j, c)
        self._input("instance", self.host_inst)
        self._input("random_poly", self.host_rnd)
        self._input("weights", self.weights)
        self.blocks = self.pk_arena.alloc(64 * 224)
        self.evals = self.pk_arena.alloc(256 * 32)
        self.counts = {}

    # ---- helpers ---------------------------------------------------------------------------------------------------------------
    def _input(self, name, arr):
        if self.rec is not None:
            self.rec.setdefault("inputs", {})[name] = np.array(arr, copy=True)

    def _pk_col(self, name, rows):
        p = self.pk_arena.alloc(rows * 32)
        self.seed += 1
        self.L.gen_scalars_dev(self.dev, self.seed, rows, 0, p)
        if self.rec is not None:
            self.L.dev_sync(self.dev)
            a = np.empty((rows, 4), dtype=np.uint64)
            self.L.d2h(self.dev, a, p)
            self._input(name, a)
        return p

    def close(self):
        self.L.dev_sync(self.dev)
        self.L.unregister_bases(self.h_g)
        self.L.unregister_bases(self.h_gl)
        self.arena.free()
        self.pk_arena.free()

    def commit(self, cols, handle, rows=None):
        """commit a list of device columns over the same SRS vector (one prover phase): -> nothing; blocks land in the ring"""
        L, rows = self.L, rows or self.n
        cols = list(cols)
        if not cols:
            return
        first = self.counts["msm"]
        assert first + len(cols) <= 64
        if self.batched and len(cols) > 1:
            L.msm_dev_batch_registered(self.dev, cols, [rows] * len(cols), handle, self.blocks + 224 * first)
        else:
            for i, c in enumerate(cols):
                L.msm_dev_registered(self.dev, c, handle, 0, rows, self.blocks + 224 * (first + i))
        self.counts["msm"] += len(cols)

    def to_coeff(self, cols):
        """copies of the Lagrange-basis columns, converted to coefficient form"""
        L, n = self.L, self.n
        out = []
        for c in cols:
            p = self.arena.alloc(n * 32)
            _d2d(L, self.dev, p, c, n * 32)
            out.append(p)
        if self.batched and len(out) > 1:
            L.lagrange_to_coeff_dev_batch(self.dev, out, self.k, self.W(self.dom.omega_inv), self.W(self.dom.ifft_divisor))
        else:
            for p in out:
                L.lagrange_to_coeff_dev(self.dev, p, self.k, self.W(self.dom.omega_inv), self.W(self.dom.ifft_divisor))
        self.counts["intt"] += len(out)
        return out

    def to_ext(self, cols):
        L, n, en = self.L, self.n, self.en
        out = []
        for c in cols:
            p = self.arena.alloc(en * 32)
            _d2d(L, self.dev, p, c, n * 32)
            out.append(p)
        if self.batched and len(out) > 1:
            L.coeff_to_extended_dev_batch(self.dev, out, self.k, self.ek, self.W(self.dom.extended_omega), self.zs)
        else:
            for p in out:
                L.coeff_to_extended_dev(self.dev, p, self.k, self.ek, self.W(self.dom.extended_omega), self.zs)
        self.counts["coset_ntt"] += len(out)
        return out

    # ---- one proof -----------------------------------------------------------------------------------------------------------------
    def run(self):
        L, dev, n, en, k = self.L, self.dev, self.n, self.en, self.k
        A, LK, d, n_adv, chunk = self.A, self.LK, self.d, self.n_adv, self.chunk
        W, dom = self.W, self.dom
        self.arena.off = 0
        self.counts = {"msm": 0, "intt": 0, "coset_ntt": 0, "eval": 0, "kate": 0}
        counts = self.counts
        phase, t_prev = {}, [time.perf_counter()]

        def mark(label):
            L.dev_sync(dev)
            t = time.perf_counter()
            phase[label] = round((t - t_prev[0]) * 1e3, 3)
            t_prev[0] = t
        # instance -> coefficient form
        d_inst = self.arena.alloc(n * 32)
        L.h2d_async(dev, d_inst, self.host_inst)
        # advice: upload, commit_lagrange
        adv = []
        for j in range(n_adv):
            p = self.arena.alloc(n * 32)
            L.h2d_async(dev, p, self.host_adv[j])
            adv.append(p)
        self.commit(adv, self.h_gl)
        mark("advice_upload_commit")
        # lookups: permute_expression_pair on the device (sort + table matching), commit both permuted columns
        perm_l = []
        for j in range(LK):
            a, s_ = self.arena.alloc(n * 32), self.arena.alloc(n * 32)
            _d2d(L, dev, a, adv[A + j], n * 32)                 # rows >= usable: the blinding rows (host randomness) stay as uploaded
            _d2d(L, dev, s_, self.table_lagrange, n * 32)
            L.lookup_permute_dev(dev, adv[A + j], self.table_lagrange, self.usable, a, s_)
            perm_l.append((a, s_))
        self.commit([c for pair in perm_l for c in pair], self.h_gl)
        mark("lookup_permuted_commit")
        # permutation grand products: z per set, chained through z[n - (blinding_factors + 1)] of the previous set
        z_l, last_z = [], self.one
        for si in range(self.sets):
            cols = list(range(si * chunk, min((si + 1) * chunk, n_adv)))
            z = self.arena.alloc(n * 32)
            L.permutation_product_dev(dev, [adv[c] for c in cols], [self.sigma_lagrange[c] for c in cols], n, self.beta, self.gamma, W(DELTA),
                                      W(pow(DELTA, cols[0], FR_MODULUS)), W(dom.omega), last_z, z)
            z_l.append(z)
            if si + 1 < self.sets:
                last_z = np.zeros(4, dtype=np.uint64)
                L.d2h(dev, last_z, z + 32 * (n - BLINDING))      # a 32-byte read-back per set
        # lookup grand products
        zl_l = []
        for j in range(LK):
            z = self.arena.alloc(n * 32)
            L.lookup_product_dev(dev, adv[A + j], self.table_lagrange, perm_l[j][0], perm_l[j][1], n, self.beta, self.gamma, z)
            zl_l.append(z)
        # the vanishing argument's random polynomial is committed in the same phase
        rnd = self.arena.alloc(n * 32)
        L.h2d_async(dev, rnd, self.host_rnd)
        self.commit(z_l + zl_l, self.h_gl)
        self.commit([rnd], self.h_g)
        mark("grand_products_commit")
        # everything to coefficient form (kept for the evaluations and the opening) and to the extended coset
        flat = [d_inst] + adv + z_l + zl_l + [c for pair in perm_l for c in pair]
        coeff = self.to_coeff(flat)
        inst_c, adv_c = coeff[0], coeff[1:1 + n_adv]
        z_c = coeff[1 + n_adv:1 + n_adv + self.sets]
        zl_c = coeff[1 + n_adv + self.sets:1 + n_adv + self.sets + LK]
        perm_c = coeff[1 + n_adv + self.sets + LK:]
        perm_c = [(perm_c[2 * j], perm_c[2 * j + 1]) for j in range(LK)]
        mark("lagrange_to_coeff")
        ext = self.to_ext(adv_c + [inst_c] + z_c + [c for zc, pc in zip(zl_c, perm_c) for c in (zc, pc[0], pc[1])])
        adv_e, inst_e = ext[:n_adv], ext[n_adv]
        z_e = ext[n_adv + 1:n_adv + 1 + self.sets]
        rest = ext[n_adv + 1 + self.sets:]
        lk_e = [tuple(rest[3 * j:3 * j + 3]) for j in range(LK)]
        mark("coeff_to_extended")
        # evaluate_h
        values = self.arena.alloc(en * 32)
        _zero(L, dev, values, en * 32)
        cols = _lib.EvalColumns(self.fixed_ext, adv_e, [inst_e], np.zeros((0, 4), dtype=np.uint64), self.beta, self.gamma, self.theta, self.y)
        L.evaluate_graph_dev(dev, self.g_gates, cols, values, en, self.rot_scale)
        L.evaluate_h_permutation_dev(dev, values, en, self.rot_scale, z_e, adv_e, self.sigma_ext, chunk, -BLINDING, self.l0, self.l_last, self.l_active,
                                     self.beta, self.gamma, self.y, W(DELTA), W(dom.g_coset), W(dom.extended_omega))
        for g, (ze, ae, se) in zip(self.g_lk, lk_e):
            L.evaluate_h_lookup_dev(dev, g, cols, values, en, self.rot_scale, ze, ae, se, self.l0, self.l_last, self.l_active)
        mark("evaluate_h")
        # h = values / (X^n - 1) -> coefficients -> d - 1 pieces of n -> commit each
        L.fr_scale_dev(dev, values, en, self.tev)
        L.extended_to_coeff_dev(dev, values, self.ek, W(dom.extended_omega_inv), self.e2c)
        counts["coset_ntt"] += 1
        pieces = [values + piece * n * 32 for piece in range(d - 1)]
        self.commit(pieces, self.h_g)
        mark("quotient_commit")
        # evaluations at x: Horner over every queried polynomial, once per rotation (all at the same point here)
        queried = [(c, 4) for c in adv_c] + [(c, 3) for c in z_c] + [(c, 2) for c in zl_c] + [(p, 1) for pc in perm_c for p in pc] + \
                  [(p, 1) for p in pieces] + [(rnd, 1)]
        for c, rotations in queried:
            for r in range(rotations):
                assert counts["eval"] < 256
                L.check(L.L.h2b_fr_eval_polynomial_dev(dev, c, n, self.x.ctypes.data, self.evals + 32 * counts["eval"], 0))
                counts["eval"] += 1
        mark("evaluations")
        # SHPLONK: per rotation set a y-weighted sum of its polynomials, divided by (X - point) for every point of the set; the
        # v-weighted sum of the quotients is committed; then the final quotient at v is committed
        rot_sets = {4: [c for c, r in queried if r == 4], 3: [c for c, r in queried if r == 3], 2: [c for c, r in queried if r == 2],
                    1: [c for c, r in queried if r == 1]}
        quot = []
        for npts, members in rot_sets.items():
            if not members:
                continue
            comb = self.arena.alloc(n * 32)
            L.fr_lincomb_dev(dev, members, self.weights[:len(members)], n, comb)
            q = comb
            for _ in range(npts):
                nxt = self.arena.alloc(n * 32)
                _zero(L, dev, nxt + (n - 1) * 32, 32)            # the quotient has n - 1 coefficients; the top slot stays zero
                L.check(L.L.h2b_fr_kate_division_dev(dev, q, n, self.x.ctypes.data, nxt, 0))
                counts["kate"] += 1
                q = nxt
            quot.append(q)
        hq = self.arena.alloc(n * 32)
        L.fr_lincomb_dev(dev, quot, self.weights[:len(quot)], n, hq)
        self.commit([hq], self.h_g)
        fin = self.arena.alloc(n * 32)
        _zero(L, dev, fin + (n - 1) * 32, 32)
        L.check(L.L.h2b_fr_kate_division_dev(dev, hq, n, self.v.ctypes.data, fin, 0))
        counts["kate"] += 1
        self.commit([fin], self.h_g)
        L.dev_sync(dev)
        blocks = np.zeros((counts["msm"], 28), dtype=np.uint64)      # the commitments (224-byte blocks) come back
        L.d2h(dev, blocks, self.blocks)
        evals = np.zeros((counts["eval"], 4), dtype=np.uint64)
        L.d2h(dev, evals, self.evals)
        mark("multiopen")
        if self.rec is not None:
            self.rec["commitments"] = blocks[:, :12].copy()
            self.rec["evaluations"] = evals.copy()
        return phase, blocks, evals


def _d2d(L, dev, dst, src, nbytes):
    L.check(L.L.h2b_memcpy_d2d_async(dev, dst, src, nbytes, 0))


def _zero(L, dev, dst, nbytes):
    L.check(L.L.h2b_memset_zero_async(dev, dst, nbytes, 0))

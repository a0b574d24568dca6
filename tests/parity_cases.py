"""Parity checks shared by the emulator tests (CPU, small sizes) and the GPU tests (C ABI on a B200)."""
import numpy as np

import bn254 as o


def omega_words(oc, k, inverse=False):
    w = o.omega_for(k)
    if inverse:
        w = pow(w, -1, o.R_MOD)
    return oc.ints_to_words([o.to_mont(w, o.R_MOD)])[0]


def affine_of(oc, jac):
    return oc.g1_to_affine(np.ascontiguousarray(jac, dtype=np.uint64))


def check_field(L, oc, n=2048):
    a, b = oc.random_fr(1, n), oc.random_fr(2, n)
    a[0] = 0
    b[1] = 0
    a[2] = b[2]
    for f in ("fr", "fq"):
        for op in ("add", "sub", "mul"):
            assert (L.field_op(f, op, a, b) == oc.field_op(f, op, a, b)).all(), (f, op)
        assert (L.field_op(f, "sqr", a) == oc.field_op(f, "mul", a, a)).all(), (f, "sqr")
    assert (L.field_op("fr", "from_mont", a) == oc.fr_from_mont(a)).all()
    assert (L.field_op("fr", "to_mont", a) == oc.fr_to_mont(a)).all()
    m = min(n, 64)
    inv = L.field_op("fq", "inv", a[3:m])
    one = L.field_op("fq", "mul", inv, a[3:m])
    R1 = oc.ints_to_words([(1 << 256) % o.P_MOD])[0]
    assert (one == R1).all()


def check_group(L, oc, n=64):
    def aff(w):
        v = oc.words_to_ints(np.ascontiguousarray(w).reshape(-1, 4))
        return [None if (v[2 * i] == 0 and v[2 * i + 1] == 0) else (o.from_mont(v[2 * i], o.P_MOD), o.from_mont(v[2 * i + 1], o.P_MOD))
                for i in range(len(v) // 2)]
    P, Q = oc.gen_points(7, n), oc.gen_points(9, n)
    Q[5] = P[5]          # doubling branch
    Q[6] = 0             # identity operands
    P[7] = 0
    Q[8, :4] = P[8, :4]  # q = -p
    Q[8, 4:] = oc.field_op("fq", "sub", np.zeros((1, 4), dtype=np.uint64), P[8:9, 4:])[0]
    pa, qa = aff(P), aff(Q)
    r = [aff(L.ec_op(op, P, Q)) for op in range(4)]
    for i in range(n):
        s = o.g1_add(pa[i], qa[i])
        assert r[0][i] == s, (i, "add")
        assert r[3][i] == s, (i, "jacobian")
        assert r[1][i] == o.g1_add(pa[i], o.g1_neg(qa[i])), (i, "sub")
        assert r[2][i] == (o.g1_mul(s, 4) if s else None), (i, "x4")


def check_ntt(L, oc, k, seed=0):
    a = oc.random_fr(0xA000 + 97 * k + seed, 1 << k)
    for inverse in (False, True):
        w = omega_words(oc, k, inverse)
        got = L.ntt(a.copy(), w, k)
        want = oc.best_fft(a, w, k)
        assert (got == want).all(), "NTT mismatch at k=%d inverse=%s" % (k, inverse)


def edge_msm_inputs(L, oc, n, kind, seed):
    s = L.gen_scalars(seed, n, kind)
    P = oc.gen_points(seed + 1, n) if n <= (1 << 16) else L.gen_points(seed + 1, n)
    if n > 40:
        s[1] = 0
        P[3] = 0
        P[5] = P[4]
        s[5] = s[4]
        P[9] = P[8]
    return s, P


def check_msm(L, oc, n, kind=0, windows=(0,), seed=1):
    s, P = edge_msm_inputs(L, oc, n, kind, seed + n)
    want = affine_of(oc, oc.best_multiexp(s, P))
    for cw in windows:
        L.set_msm_window(cw)
        try:
            got = affine_of(oc, L.msm(s, P))
        finally:
            L.set_msm_window(0)
        assert (got == want).all(), "MSM mismatch n=%d kind=%d window=%d" % (n, kind, cw)


def check_msm_tables(L, oc, n, spacing, kind=0, seed=5, windows=(0,), ranges=None):
    """Registered base set with precomputed window tables: full range and sub-ranges, forced table spacing."""
    s, P = edge_msm_inputs(L, oc, n, kind, seed + n)
    L.set_msm_precomp(spacing)
    try:
        h = L.register_bases(P)
    finally:
        L.set_msm_precomp(0)
    try:
        info = L.base_set_info(h)
        assert info["n_tables"] > 1 and info["spacing"] == (spacing or info["spacing"]), info
        for (off, m) in (ranges or [(0, n)]):
            want = affine_of(oc, oc.best_multiexp(s[off:off + m], P[off:off + m]))
            for cw in windows:
                L.set_msm_window(cw)
                try:
                    got = affine_of(oc, L.msm_registered(s[off:off + m], h, off))
                finally:
                    L.set_msm_window(0)
                assert (got == want).all(), "table MSM mismatch n=%d range=(%d,%d) spacing=%d window=%d" % (n, off, m, spacing, cw)
    finally:
        L.unregister_bases(h)


def check_msm_single_bucket(L, oc, n, scalar=1, tables=True):
    """Every scalar equal: the whole sorted list is one bucket, cut into many slices (multi-chunk tree combine)."""
    P = oc.gen_points(901, n) if n <= (1 << 16) else L.gen_points(901, n)
    one = oc.fr_to_mont(np.array([[scalar, 0, 0, 0]], dtype=np.uint64))
    s = np.repeat(one, n, axis=0)
    want = affine_of(oc, oc.best_multiexp(s, P))
    if tables:
        h = L.register_bases(P)
        try:
            got = affine_of(oc, L.msm_registered(s, h))
        finally:
            L.unregister_bases(h)
    else:
        got = affine_of(oc, L.msm(s, P))
    assert (got == want).all()


def check_msm_random(L, oc, examples, max_n, spacings, windows):
    """Property test over the MSM's shape space: random length, scalar mix (incl. 0, 1, r-1, duplicates, identity bases),
    table spacing / plain mode (-1), forced window, sub-range of a registered set -- always the oracle's affine result."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    one = oc.fr_to_mont(np.array([[1, 0, 0, 0]], dtype=np.uint64))[0]
    minus_one = oc.field_op("fr", "sub", np.zeros((1, 4), dtype=np.uint64), one.reshape(1, 4))[0]

    @settings(max_examples=examples, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(n=st.integers(1, max_n), seed=st.integers(0, 10 ** 6), spacing=st.sampled_from(list(spacings)),
           window=st.sampled_from(list(windows)), kind=st.integers(0, 1), off_frac=st.floats(0, 0.9))
    def run(n, seed, spacing, window, kind, off_frac):
        s = L.gen_scalars(seed, n, kind)
        P = oc.gen_points(seed + 1, n)
        rng = np.random.default_rng(seed)
        for i in rng.integers(0, n, size=min(n, 6)):
            choice = rng.integers(0, 5)
            if choice == 0:
                s[i] = 0
            elif choice == 1:
                s[i] = one
            elif choice == 2:
                s[i] = minus_one
            elif choice == 3:
                P[i] = 0
            else:
                P[i] = P[rng.integers(0, n)]
        if spacing < 0:
            L.set_msm_window(window)
            try:
                got = affine_of(oc, L.msm(s, P))
            finally:
                L.set_msm_window(0)
            want = affine_of(oc, oc.best_multiexp(s, P))
        else:
            off = int(off_frac * n)
            m = n - off
            L.set_msm_precomp(spacing)
            try:
                h = L.register_bases(P)
            finally:
                L.set_msm_precomp(0)
            try:
                L.set_msm_window(window)
                got = affine_of(oc, L.msm_registered(s[:m], h, off))
            finally:
                L.set_msm_window(0)
                L.unregister_bases(h)
            want = affine_of(oc, oc.best_multiexp(s[:m], P[off:off + m]))
        assert (got == want).all(), (n, seed, spacing, window, kind, off_frac)

    run()


def check_grand_product_blocks(L, oc, sizes):
    """batch inversion (zeros stay zero) and exclusive prefix product against the oracle, ragged sizes"""
    for n in sizes:
        a = oc.random_fr(0xD000 + n, n)
        if n > 3:
            a[1] = 0
            a[n - 1] = 0
            a[n // 2] = a[n // 2 - 1]
        assert (L.fr_batch_invert(a) == oc.fr_batch_invert(a)).all(), ("batch_invert", n)
        assert (L.fr_prefix_product(a) == oc.fr_prefix_product(a)).all(), ("prefix_product", n)
        if n > 8:
            b = a.copy()
            b[b.sum(axis=1) == 0] = a[0]                                   # no zeros: the product never collapses
            assert (L.fr_prefix_product(b) == oc.fr_prefix_product(b)).all(), ("prefix_product nz", n)


def check_poly_eval_and_division(L, oc, sizes):
    """eval_polynomial and kate_division against the oracle, plus the identity a(X) = q(X) (X - b) + a(b) at a random point"""
    for n in sizes:
        a = oc.random_fr(0xE000 + n, n)
        x, b = oc.random_fr(0xE100 + n, 1)[0], oc.random_fr(0xE200 + n, 1)[0]
        assert (L.fr_eval_polynomial(a, x) == oc.fr_eval_polynomial(a, x)).all(), ("eval", n)
        q = L.fr_kate_division(a, b)
        assert (q == oc.fr_kate_division(a, b)).all(), ("kate", n)
        if n > 1:
            # a(x) == q(x) * (x - b) + a(b), all through the device evaluation
            ax, qx, ab = L.fr_eval_polynomial(a, x), L.fr_eval_polynomial(q, x), L.fr_eval_polynomial(a, b)
            xmb = oc.field_op("fr", "sub", x.reshape(1, 4), b.reshape(1, 4))
            rhs = oc.field_op("fr", "add", oc.field_op("fr", "mul", qx.reshape(1, 4), xmb), ab.reshape(1, 4))
            assert (rhs[0] == ax).all(), ("identity", n)


def check_golden_ntt(L, g):
    for k in (1, 2, 3, 5, 8):
        for tag, wkey in (("fwd", "omega"), ("inv", "omega_inv")):
            a = np.array(g["k%d_in" % k], dtype=np.uint64, order="C")
            got = L.ntt(a, g["k%d_%s" % (k, wkey)], k)
            assert (got == g["k%d_%s" % (k, tag)]).all(), (k, tag)


def check_golden_msm(L, oc, g):
    for tag in ("n1", "n2", "n8", "n64", "n200", "cancel", "anchor"):
        got = affine_of(oc, L.msm(g[tag + "_scalars"], g[tag + "_bases"]))
        assert (got == g[tag + "_result"]).all(), tag

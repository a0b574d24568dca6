"""
ORACLE (test infrastructure only). ctypes binding of oracle/libh2oracle.so (built from
h2_oracle.cpp by oracle/Makefile) plus numpy <-> big-int packing helpers.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libh2oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "h2_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libh2oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u64p = ctypes.c_void_p
        L.orc_hardware_threads.restype = ctypes.c_int
        L.orc_best_multiexp.argtypes = [u64p, u64p, ctypes.c_size_t, ctypes.c_int, u64p]
        L.orc_best_fft.argtypes = [u64p, u64p, ctypes.c_uint32, ctypes.c_int]
        L.orc_g1_to_affine.argtypes = [u64p, u64p]
        L.orc_g1_sum.argtypes = [u64p, ctypes.c_size_t, u64p]
        L.orc_field_op.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, ctypes.c_size_t, u64p]
        L.orc_fr_to_mont.argtypes = [u64p, ctypes.c_size_t, u64p]
        L.orc_fr_from_mont.argtypes = [u64p, ctypes.c_size_t, u64p]
        L.orc_fr_scale.argtypes = [u64p, ctypes.c_size_t, u64p]
        L.orc_fr_batch_invert.argtypes = [u64p, ctypes.c_size_t]
        L.orc_fr_prefix_product.argtypes = [u64p, ctypes.c_size_t, u64p]
        L.orc_fr_eval_polynomial.argtypes = [u64p, ctypes.c_size_t, u64p, u64p]
        L.orc_fr_kate_division.argtypes = [u64p, ctypes.c_size_t, u64p, u64p]
        L.orc_random_fr.argtypes = [ctypes.c_uint64, ctypes.c_size_t, u64p]
        L.orc_permutation_product.argtypes = [u64p, u64p, ctypes.c_uint32, ctypes.c_size_t] + [u64p] * 7
        L.orc_lookup_product.argtypes = [u64p] * 4 + [ctypes.c_size_t] + [u64p] * 3
        L.orc_lookup_permute.argtypes = [u64p, u64p, ctypes.c_size_t, u64p, u64p]
        L.orc_g1_from_bytes.argtypes = [u64p, ctypes.c_size_t, ctypes.c_int, u64p]
        L.orc_g1_from_bytes.restype = ctypes.c_size_t
        L.orc_g1_to_bytes.argtypes = [u64p, ctypes.c_size_t, u64p]
        u32, i32 = ctypes.c_uint32, ctypes.c_int32
        graph = [u64p, u32, u64p, u32, u64p, u32, u64p, u32, u32]           # constants, rotations, calculations, parts, n_intermediates
        columns = [u64p, u64p, u64p, u64p, u64p, u64p, u64p, u64p]          # fixed, advice, instance, challenges, beta, gamma, theta, y
        L.orc_evaluate_graph.argtypes = graph + columns + [u64p, u32, i32]
        L.orc_evaluate_h_lookup.argtypes = graph + columns + [u64p, u32, i32, u64p, u64p, u64p, u64p, u64p, u64p]
        L.orc_evaluate_h_permutation.argtypes = [u64p, u32, i32, u64p, u32, u64p, u64p, u32, u32, i32] + [u64p] * 9
        L.orc_gen_points.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int, u64p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def hardware_threads() -> int:
    return lib().orc_hardware_threads()


# ---- packing: python ints <-> (n,4) uint64 little-endian limb arrays -------------------------
def ints_to_words(vals, words: int = 4) -> np.ndarray:
    out = np.empty((len(vals), words), dtype=np.uint64)
    buf = b"".join(int(v).to_bytes(8 * words, "little") for v in vals)
    out[:] = np.frombuffer(buf, dtype="<u8").reshape(len(vals), words)
    return out


def words_to_ints(arr: np.ndarray):
    arr = np.ascontiguousarray(arr, dtype=np.uint64)
    w = arr.shape[-1]
    flat = arr.reshape(-1, w)
    raw = flat.astype("<u8").tobytes()
    return [int.from_bytes(raw[i * 8 * w:(i + 1) * 8 * w], "little") for i in range(flat.shape[0])]


# ---- thin wrappers ----------------------------------------------------------------------------
def random_fr(seed: int, n: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_random_fr(seed, n, _p(out))
    return out


def gen_points(seed: int, n: int, threads: int = 0) -> np.ndarray:
    out = np.empty((n, 8), dtype=np.uint64)
    lib().orc_gen_points(seed, n, threads or hardware_threads(), _p(out))
    return out


def fr_to_mont(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fr_to_mont(_p(a), a.shape[0], _p(out))
    return out


def fr_from_mont(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fr_from_mont(_p(a), a.shape[0], _p(out))
    return out


def fr_scale(a: np.ndarray, s: np.ndarray) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, order="C", copy=True)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    lib().orc_fr_scale(_p(a), a.shape[0], _p(s))
    return a


def fr_batch_invert(a: np.ndarray) -> np.ndarray:
    """a[i] -> 1 / a[i], zeros stay zero (ff::BatchInvert semantics)"""
    a = np.array(a, dtype=np.uint64, order="C", copy=True)
    lib().orc_fr_batch_invert(_p(a), a.shape[0])
    return a


def fr_prefix_product(a: np.ndarray) -> np.ndarray:
    """out[0] = 1, out[i] = a[0] * ... * a[i-1]"""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_fr_prefix_product(_p(a), a.shape[0], _p(out))
    return out


def fr_eval_polynomial(a: np.ndarray, x: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    x = np.ascontiguousarray(x, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_fr_eval_polynomial(_p(a), a.shape[0], _p(x), _p(out))
    return out


def fr_kate_division(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    q = np.zeros((max(a.shape[0] - 1, 0), 4), dtype=np.uint64)
    if a.shape[0] > 1:
        lib().orc_fr_kate_division(_p(a), a.shape[0], _p(b), _p(q))
    return q


# ---- SRS point encodings ([UP] halo2curves GroupEncoding for G1Affine) -------------------------------------------------
def g1_from_bytes(b: np.ndarray, threads: int = 0):
    """(n, 32) uint8 compressed points -> ((n, 8) uint64 affine Montgomery, index of the first invalid encoding or n)"""
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32)
    out = np.empty((b.shape[0], 8), dtype=np.uint64)
    first = lib().orc_g1_from_bytes(b.ctypes.data, b.shape[0], threads or hardware_threads(), out.ctypes.data)
    return out, int(first)


def g1_to_bytes(aff: np.ndarray) -> np.ndarray:
    aff = np.ascontiguousarray(aff, dtype=np.uint64).reshape(-1, 8)
    out = np.empty((aff.shape[0], 32), dtype=np.uint8)
    lib().orc_g1_to_bytes(aff.ctypes.data, aff.shape[0], out.ctypes.data)
    return out


# ---- quotient evaluation ([UP] halo2_proofs/src/plonk/evaluation.rs) ------------------------------------------------
def _col(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)


def _ptrs(cols):
    cols = [_col(c) for c in cols]
    return cols, np.array([c.ctypes.data for c in cols], dtype=np.uint64)


def _graph_args(graph):
    """graph: (constants (n,4) u64, rotations int32, calculations (n,10) u32, parts (m,3) u32, n_intermediates)"""
    constants, rotations, calcs, parts, n_inter = graph
    constants = _col(constants)
    rotations = np.ascontiguousarray(rotations, dtype=np.int32).reshape(-1)
    calcs = np.ascontiguousarray(calcs, dtype=np.uint32).reshape(-1, 10)
    parts = np.ascontiguousarray(parts, dtype=np.uint32).reshape(-1, 3)
    keep = (constants, rotations, calcs, parts)
    return keep, [constants.ctypes.data, constants.shape[0], rotations.ctypes.data, rotations.shape[0], calcs.ctypes.data, calcs.shape[0],
                  parts.ctypes.data, parts.shape[0], int(n_inter)]


def _column_args(fixed, advice, instance, challenges, beta, gamma, theta, y):
    kf, pf = _ptrs(fixed)
    ka, pa = _ptrs(advice)
    ki, pi = _ptrs(instance)
    scal = [_col(v) for v in (challenges, beta, gamma, theta, y)]
    return (kf, ka, ki, pf, pa, pi, scal), [pf.ctypes.data, pa.ctypes.data, pi.ctypes.data] + [v.ctypes.data for v in scal]


def evaluate_graph(graph, fixed, advice, instance, challenges, beta, gamma, theta, y, values, rot_scale: int) -> np.ndarray:
    """the "Custom gates" loop of evaluate_h, sequentially as upstream walks it -> new values"""
    values = np.array(values, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    k1, ga = _graph_args(graph)
    k2, ca = _column_args(fixed, advice, instance, challenges, beta, gamma, theta, y)
    lib().orc_evaluate_graph(*ga, *ca, values.ctypes.data, values.shape[0], rot_scale)
    return values


def evaluate_h_lookup(graph, fixed, advice, instance, challenges, beta, gamma, theta, y, values, rot_scale: int, product_coset,
                      permuted_input_coset, permuted_table_coset, l0, l_last, l_active_row) -> np.ndarray:
    values = np.array(values, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    k1, ga = _graph_args(graph)
    k2, ca = _column_args(fixed, advice, instance, challenges, beta, gamma, theta, y)
    extra = [_col(v) for v in (product_coset, permuted_input_coset, permuted_table_coset, l0, l_last, l_active_row)]
    lib().orc_evaluate_h_lookup(*ga, *ca, values.ctypes.data, values.shape[0], rot_scale, *[v.ctypes.data for v in extra])
    return values


def evaluate_h_permutation(values, rot_scale: int, product_cosets, columns, perm_cosets, chunk_len: int, last_rotation: int, l0, l_last,
                           l_active_row, beta, gamma, y, delta, zeta, extended_omega) -> np.ndarray:
    values = np.array(values, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    kp, pp = _ptrs(product_cosets)
    kc, pc = _ptrs(columns)
    ks, ps = _ptrs(perm_cosets)
    extra = [_col(v) for v in (l0, l_last, l_active_row, beta, gamma, y, delta, zeta, extended_omega)]
    lib().orc_evaluate_h_permutation(values.ctypes.data, values.shape[0], rot_scale, pp.ctypes.data, len(kp), pc.ctypes.data, ps.ctypes.data, len(kc),
                                     chunk_len, last_rotation, *[v.ctypes.data for v in extra])
    return values


def permutation_product(values, sigma, beta, gamma, delta, deltaomega, omega, last_z) -> np.ndarray:
    """[UP] permutation::Argument::commit for one set of columns -> z (n x 4), blinding rows not applied"""
    kv, pv = _ptrs(values)
    ks, ps = _ptrs(sigma)
    n = kv[0].shape[0]
    z = np.empty((n, 4), dtype=np.uint64)
    sc = [_col(v) for v in (beta, gamma, delta, deltaomega, omega, last_z)]
    lib().orc_permutation_product(pv.ctypes.data, ps.ctypes.data, len(kv), n, *[v.ctypes.data for v in sc], z.ctypes.data)
    return z


def lookup_product(compressed_input, compressed_table, permuted_input, permuted_table, beta, gamma) -> np.ndarray:
    cols = [_col(v) for v in (compressed_input, compressed_table, permuted_input, permuted_table)]
    n = cols[0].shape[0]
    z = np.empty((n, 4), dtype=np.uint64)
    sc = [_col(v) for v in (beta, gamma)]
    lib().orc_lookup_product(*[c.ctypes.data for c in cols], n, *[v.ctypes.data for v in sc], z.ctypes.data)
    return z


def lookup_permute(input_expression, table_expression, usable_rows: int):
    """[UP] permute_expression_pair -> (permuted_input, permuted_table) of usable_rows rows; raises ValueError when an input value
    is not in the table"""
    a, t = _col(input_expression), _col(table_expression)
    pa, pt = np.empty((usable_rows, 4), dtype=np.uint64), np.empty((usable_rows, 4), dtype=np.uint64)
    if lib().orc_lookup_permute(a.ctypes.data, t.ctypes.data, usable_rows, pa.ctypes.data, pt.ctypes.data):
        raise ValueError("ConstraintSystemFailure: input value not in the table")
    return pa, pt


def field_op(field: str, op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.empty_like(a)
    lib().orc_field_op({"fr": 0, "fq": 1}[field], {"add": 0, "sub": 1, "mul": 2}[op], _p(a), _p(b), a.shape[0], _p(out))
    return out


def best_multiexp(scalars: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
    """-> Jacobian (12,) uint64, Montgomery. Restates halo2_proofs::arithmetic::best_multiexp."""
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    assert scalars.shape[0] == bases.shape[0]
    out = np.empty(12, dtype=np.uint64)
    lib().orc_best_multiexp(_p(scalars), _p(bases), scalars.shape[0], threads or hardware_threads(), _p(out))
    return out


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int = 0) -> np.ndarray:
    """-> new array; restates halo2_proofs::arithmetic::best_fft (natural order, no scaling)."""
    a = np.array(a, dtype=np.uint64, order="C", copy=True)
    omega = np.ascontiguousarray(omega, dtype=np.uint64)
    assert a.shape[0] == 1 << log_n
    lib().orc_best_fft(_p(a), _p(omega), log_n, threads or hardware_threads())
    return a


def g1_to_affine(jac: np.ndarray) -> np.ndarray:
    jac = np.ascontiguousarray(jac, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_g1_to_affine(_p(jac), _p(out))
    return out


def g1_sum(jacs: np.ndarray) -> np.ndarray:
    jacs = np.ascontiguousarray(jacs, dtype=np.uint64).reshape(-1, 12)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_sum(_p(jacs), jacs.shape[0], _p(out))
    return out

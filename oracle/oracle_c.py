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

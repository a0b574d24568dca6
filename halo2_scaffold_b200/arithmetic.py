"""
Host-side mirror of halo2_proofs::arithmetic for the two functions on the hot path
([UP] halo2_proofs/src/arithmetic.rs @ v2023_02_02).  Same names, argument
meaning and error behaviour as the reference: infallible apart from the length assertions, which
raise AssertionError where the Rust code panics.  Arrays are numpy uint64 in halo2curves' memory
layout: Fr (n,4), G1Affine (n,8), G1 Jacobian (12,).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _lib_ready(lib=None):
    L = lib or _lib.load()
    if L.device_count() == 0:
        L.init(0)
    return L


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray, lib=None) -> np.ndarray:
    """sum_i coeffs[i] * bases[i]  ->  G1 Jacobian (x|y|z Montgomery); `assert_eq!(coeffs.len(), bases.len())`."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert coeffs.shape[0] == bases.shape[0], "best_multiexp: coeffs.len() != bases.len()"
    return _lib_ready(lib).msm(coeffs, bases)


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, lib=None) -> np.ndarray:
    """In-place forward DFT of `a` (2^log_n Fr elements) with root `omega`; natural order, no scaling."""
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "a must be a contiguous uint64 array"
    assert a.size == 4 << log_n, "best_fft: a.len() != 1 << log_n"
    return _lib_ready(lib).ntt(a, np.ascontiguousarray(omega, dtype=np.uint64), log_n)

"""
Host-side mirror of halo2_proofs::poly::kzg::commitment::ParamsKZG for the callers of
best_multiexp ([UP] halo2_proofs/src/poly/kzg/commitment.rs): the two SRS vectors
`g` and `g_lagrange` are registered once (device resident on every GPU), and `commit` /
`commit_lagrange` are MSMs over their prefixes.  The `Blind` argument of the reference is ignored
by KZG and therefore absent here.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class ParamsKZG:
    def __init__(self, k: int, g: np.ndarray, g_lagrange: np.ndarray, lib=None):
        self.lib = lib or _lib.load()
        if self.lib.device_count() == 0:
            self.lib.init(0)
        self.k = k
        self.n = 1 << k
        g = np.ascontiguousarray(g, dtype=np.uint64).reshape(-1, 8)
        g_lagrange = np.ascontiguousarray(g_lagrange, dtype=np.uint64).reshape(-1, 8)
        assert g.shape[0] == self.n and g_lagrange.shape[0] == self.n
        self._g = self.lib.register_bases(g)
        self._g_lagrange = self.lib.register_bases(g_lagrange)

    def commit(self, poly: np.ndarray) -> np.ndarray:
        """best_multiexp(poly, g[..poly.len()]) -> G1 Jacobian"""
        poly = np.ascontiguousarray(poly, dtype=np.uint64).reshape(-1, 4)
        assert poly.shape[0] <= self.n
        return self.lib.msm_registered(poly, self._g, 0)

    def commit_lagrange(self, poly: np.ndarray) -> np.ndarray:
        """best_multiexp(poly, g_lagrange[..poly.len()]) -> G1 Jacobian"""
        poly = np.ascontiguousarray(poly, dtype=np.uint64).reshape(-1, 4)
        assert poly.shape[0] <= self.n
        return self.lib.msm_registered(poly, self._g_lagrange, 0)

    def close(self):
        for h in (self._g, self._g_lagrange):
            if h:
                self.lib.unregister_bases(h)
        self._g = self._g_lagrange = 0

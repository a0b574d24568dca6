"""
Host-side mirror of halo2_proofs::poly::EvaluationDomain for the callers of best_fft
([UP] halo2_proofs/src/poly/domain.rs).  The domain constants
(omega, extended omega, divisors, zeta) are a handful of scalar modular operations done once per
domain on the host, exactly as the Rust code does; every vector operation runs on the GPU: the
column is uploaded once, scaled / padded / transformed on the device, and downloaded once.
"""
from __future__ import annotations

import numpy as np

from . import _lib

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FR_S = 28
FR_ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
_R = 1 << 256


def fr_to_words(x: int) -> np.ndarray:
    """canonical integer -> Montgomery 4 x u64"""
    m = (x % FR_MODULUS) * _R % FR_MODULUS
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class EvaluationDomain:
    """EvaluationDomain::new(j, k): j = cs.degree(), n = 2^k."""

    def __init__(self, j: int, k: int, lib=None, device: int = 0):
        self.lib = lib or _lib.load()
        if self.lib.device_count() == 0:
            self.lib.init(0)
        self.device = device
        self.k = k
        self.n = 1 << k
        self.quotient_poly_degree = j - 1
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        assert ek <= FR_S, "extended domain exceeds the two-adicity of Fr"
        self.extended_k = ek
        w = FR_ROOT_OF_UNITY
        for _ in range(ek, FR_S):
            w = w * w % FR_MODULUS
        self.extended_omega = w
        for _ in range(k, ek):
            w = w * w % FR_MODULUS
        self.omega = w
        self.omega_inv = pow(self.omega, -1, FR_MODULUS)
        self.extended_omega_inv = pow(self.extended_omega, -1, FR_MODULUS)
        self.ifft_divisor = pow(1 << k, -1, FR_MODULUS)
        self.extended_ifft_divisor = pow(1 << ek, -1, FR_MODULUS)
        self.g_coset = FR_ZETA
        self.g_coset_inv = FR_ZETA * FR_ZETA % FR_MODULUS
        # t_evaluations[i] = 1 / t(zeta * extended_omega^i), t(X) = X^n - 1: it repeats with period 2^(extended_k - k)
        orig, step = pow(FR_ZETA, self.n, FR_MODULUS), pow(self.extended_omega, self.n, FR_MODULUS)
        cur, t = orig, []
        while True:
            t.append(cur)
            cur = cur * step % FR_MODULUS
            if cur == orig:
                break
        assert len(t) == 1 << (ek - k)
        self.t_evaluations = [pow(v - 1, -1, FR_MODULUS) for v in t]

    # -- helpers ----------------------------------------------------------------------------------
    def _run(self, a: np.ndarray, out_len: int, work_len: int, steps, alloc_len: int | None = None):
        """upload `work_len` elements into a device buffer of `alloc_len` (default work_len), run `steps`, download `out_len`"""
        L, dev = self.lib, self.device
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        d = L.dev_alloc(dev, (alloc_len or work_len) * 32)
        try:
            if work_len > a.shape[0]:
                pad = np.zeros((work_len, 4), dtype=np.uint64)
                pad[: a.shape[0]] = a
                L.h2d(dev, d, pad)
            else:
                L.h2d(dev, d, a[:work_len])
            for step in steps:
                step(d)
            L.dev_sync(dev)
            out = np.empty((out_len, 4), dtype=np.uint64)
            L.d2h(dev, out, d)
        finally:
            L.dev_free(dev, d)
        return out

    def lagrange_to_coeff(self, a: np.ndarray) -> np.ndarray:
        """best_fft(a, omega_inv, k); a[i] *= 1/n"""
        assert a.size == 4 * self.n
        L, dev = self.lib, self.device
        return self._run(a, self.n, self.n, [
            lambda d: L.lagrange_to_coeff_dev(dev, d, self.k, fr_to_words(self.omega_inv), fr_to_words(self.ifft_divisor)),
        ])

    def coeff_to_extended(self, a: np.ndarray) -> np.ndarray:
        """a[i] *= zeta^(i mod 3); zero-pad to 2^extended_k; best_fft(a, extended_omega, extended_k)"""
        assert a.size == 4 * self.n
        L, dev = self.lib, self.device
        en = 1 << self.extended_k
        zs = np.stack([fr_to_words(1), fr_to_words(self.g_coset), fr_to_words(self.g_coset_inv)])
        return self._run(a, en, self.n, [
            lambda d: L.coeff_to_extended_dev(dev, d, self.k, self.extended_k, fr_to_words(self.extended_omega), zs),
        ], alloc_len=en)

    def extended_to_coeff(self, a: np.ndarray) -> np.ndarray:
        """best_fft(a, extended_omega_inv, extended_k); scale by 1/2^ek and un-zeta; truncate to n*(j-1)"""
        en = 1 << self.extended_k
        assert a.size == 4 * en
        L, dev = self.lib, self.device
        zs = np.stack([fr_to_words(self.extended_ifft_divisor),
                       fr_to_words(self.extended_ifft_divisor * self.g_coset_inv),
                       fr_to_words(self.extended_ifft_divisor * self.g_coset)])
        return self._run(a, self.n * self.quotient_poly_degree, en, [
            lambda d: L.extended_to_coeff_dev(dev, d, self.extended_k, fr_to_words(self.extended_omega_inv), zs),
        ])

    def divide_by_vanishing_poly(self, a: np.ndarray) -> np.ndarray:
        """a[i] *= t_evaluations[i % len]: the extended-coset evaluations of h(X) = (gate combination) / (X^n - 1)"""
        en = 1 << self.extended_k
        assert a.size == 4 * en
        L, dev = self.lib, self.device
        t = np.stack([fr_to_words(v) for v in self.t_evaluations])
        return self._run(a, en, en, [lambda d: L.fr_scale_dev(dev, d, en, t)])

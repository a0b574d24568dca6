"""
Host-side mirrors of the two argument provers that sit between the commitments and evaluate_h
([UP] halo2_proofs/src/plonk/permutation/prover.rs `Argument::commit`, plonk/lookup/prover.rs
`Argument::commit_permuted` / `Permuted::commit_product`).  The per-row work runs on the GPU
(h2b_permutation_product_dev, h2b_lookup_permute_dev, h2b_lookup_product_dev); what stays on the host is what the
Rust code keeps on the host too: the chaining of the sets through `last_z`, the blinding rows (randomness) and the
commitments' bookkeeping.  Names and argument meaning follow upstream.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from . import _lib
from .domain import FR_MODULUS, fr_to_words

FR_DELTA = pow(7, 1 << 28, FR_MODULUS)           # Fr::DELTA = GENERATOR^(2^S)


def _rows(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)


def permutation_commit(columns: Sequence[np.ndarray], permutations: Sequence[np.ndarray], *, chunk_len: int, blinding_factors: int, beta, gamma, omega,
                       blind: Callable[[int], np.ndarray], params=None, lib=None, device: int = 0):
    """permutation::Argument::commit: one product polynomial z per chunk of `chunk_len` columns.
    columns / permutations: the Lagrange-basis values of every column of the argument and of its sigma polynomial (n x 4 words each);
    blind(count) -> (count, 4) random field elements for the blinding rows; params: a ParamsKZG to commit with (optional).
    -> (list of z columns, list of commitments or None).  z[0] of the first set is one; every later set starts at the previous set's
    z[n - (blinding_factors + 1)]; rows n - blinding_factors .. n - 1 are overwritten with blinding values, as upstream does."""
    L = lib or _lib.load()
    assert len(columns) == len(permutations) and chunk_len >= 1
    n = _rows(columns[0]).shape[0]
    last_z = fr_to_words(1)
    zs, commitments = [], []
    for lo in range(0, len(columns), chunk_len):
        hi = min(lo + chunk_len, len(columns))
        z = L.permutation_product([_rows(c) for c in columns[lo:hi]], [_rows(p) for p in permutations[lo:hi]], beta, gamma, fr_to_words(FR_DELTA),
                                  fr_to_words(pow(FR_DELTA, lo, FR_MODULUS)), omega, last_z, device=device)
        if blinding_factors:
            z[n - blinding_factors:] = _rows(blind(blinding_factors))
        last_z = z[n - (blinding_factors + 1)].copy()
        zs.append(z)
        commitments.append(params.commit_lagrange(z) if params is not None else None)
    return zs, commitments


def lookup_compress_expressions(expressions: Sequence, *, fixed, advice, instance, challenges, theta, lib=None, device: int = 0) -> np.ndarray:
    """lookup::Argument::commit_permuted's `compress_expressions`: theta^(m-1) e_0 + ... + e_(m-1) over the n Lagrange rows, every expression
    evaluated at the row (rotations wrap inside the n rows, rot_scale = 1).  The expressions go through GraphEvaluator::add_expression and
    one Horner in theta -- the same graph Evaluator::new builds for the extended-coset side -- and run on the device."""
    from . import evaluation as ev
    L = lib or _lib.load()
    g = ev.GraphEvaluator()
    parts = [g.add_expression(e) for e in expressions]
    g.add_calculation(ev.HORNER, ev.ValueSource(ev.CONSTANT, 0), ev.ValueSource(ev.THETA), parts)
    n = _rows((list(fixed) + list(advice) + list(instance))[0]).shape[0]
    zero = fr_to_words(0)
    return ev.evaluate_graph(L, g.arrays(), fixed, advice, instance, challenges, zero, zero, theta, zero, np.zeros((n, 4), dtype=np.uint64), 1,
                             device=device)


def lookup_commit_permuted(compressed_input: np.ndarray, compressed_table: np.ndarray, *, blinding_factors: int, blind: Callable[[int], np.ndarray],
                           params=None, lib=None, device: int = 0):
    """lookup::Argument::commit_permuted after compress_expressions: permute_expression_pair on the usable rows, blinding rows appended.
    -> (permuted_input, permuted_table, commitments or None).  Raises H2BError (upstream: Error::ConstraintSystemFailure) when an input
    value is not in the table."""
    L = lib or _lib.load()
    a, s = _rows(compressed_input), _rows(compressed_table)
    n = a.shape[0]
    usable = n - (blinding_factors + 1)
    pa, pt = L.lookup_permute(a, s, usable, device=device)
    pa = np.concatenate([pa, _rows(blind(blinding_factors + 1))]) if n > usable else pa
    pt = np.concatenate([pt, _rows(blind(blinding_factors + 1))]) if n > usable else pt
    com = (params.commit_lagrange(pa), params.commit_lagrange(pt)) if params is not None else None
    return pa, pt, com


def lookup_commit_product(compressed_input, compressed_table, permuted_input, permuted_table, *, blinding_factors: int, beta, gamma,
                          blind: Callable[[int], np.ndarray], params=None, lib=None, device: int = 0):
    """lookup::Permuted::commit_product: z[0] = 1, z[i + 1] = z[i] (a + beta)(s + gamma) / ((a' + beta)(s' + gamma)) over the first
    n - blinding_factors rows, blinding values after them.  -> (z, commitment or None)"""
    L = lib or _lib.load()
    z = L.lookup_product(_rows(compressed_input), _rows(compressed_table), _rows(permuted_input), _rows(permuted_table), beta, gamma, device=device)
    n = z.shape[0]
    if blinding_factors:
        z[n - blinding_factors:] = _rows(blind(blinding_factors))
    return z, (params.commit_lagrange(z) if params is not None else None)

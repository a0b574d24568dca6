"""
halo2_scaffold_b200 -- B200-native (sm_100a) hot path of the Halo2 KZG prover as
DCMMC/halo2-scaffold exercises it: BN254 G1 MSM (`best_multiexp`) and the Fr NTT (`best_fft`),
hand-written CUDA behind the C ABI of include/h2b200.h.

This python package is the thin host-side mirror of the reference's operator interface for this
path (same names and argument meaning as halo2_proofs::arithmetic / poly::domain / poly::kzg) that
the parity tests drive; the product is csrc/ -> lib/libh2b200.so.  There is no CPU fallback.
"""
from ._lib import H2BError, Lib, load, exported_symbols  # noqa: F401
from .arithmetic import best_fft, best_multiexp  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .kzg import ParamsKZG  # noqa: F401

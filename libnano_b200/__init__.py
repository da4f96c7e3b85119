"""libnano_b200 — B200-native (sm_100a) value-and-gradient evaluation of libnano's linear-model objective.

Only what the hot path needs: the CUDA kernels + C ABI (csrc/, include/nb200.h) and the host-side
mirror of the reference interface (`Function`, `LinearFunction`, `SyntheticFunction`, `Loss`, `Dataset`, and the
caller around them, `LinearModel` = linear_t::fit).
There is no CPU fallback: every evaluation is a CUDA launch and a missing libnb200.so raises.
"""
from .function import (FLAVOUR_LINEAR, FLAVOUR_SYNTH, SYNTH_CLASSIFICATION, SYNTH_MULTICLASS, SYNTH_REGRESSION,
                       Context, Dataset, Function, LinearFunction, SyntheticFunction, critical)
from .gboost import GBoostBiasFunction, GBoostGradsFunction, GBoostScaleFunction
from .linear import LinearModel
from .loss import LOSS_IDS, Loss
from ._lib import LIB_PATH, Nb200Error

__version__ = "0.1.0"

__all__ = [
    "Context", "Dataset", "Function", "LinearFunction", "SyntheticFunction", "LinearModel", "Loss", "LOSS_IDS", "critical",
    "Nb200Error", "LIB_PATH", "FLAVOUR_LINEAR", "FLAVOUR_SYNTH", "SYNTH_CLASSIFICATION", "SYNTH_REGRESSION",
    "SYNTH_MULTICLASS", "GBoostBiasFunction", "GBoostScaleFunction", "GBoostGradsFunction",
]

"""pda_b200: B200-native (sm_100a) implementation of PDA's BPR-MF train step and all-items scoring.

Only the hot path lives here: csrc/ holds the CUDA kernels and the C ABI
(include/pda_b200.h), the Python modules mirror the reference's model / recommender interface.
"""
from ._lib import PdaError, load  # noqa: F401
from .model import PDAModel  # noqa: F401

__all__ = ["PDAModel", "PdaError", "load"]

"""rotmv_b200 -- B200-native (sm_100a) implementation of the Rot-MVGaze multi-view hot path.

Host side mirrors the reference operator surface (models/rot_mv.py::FeatRotationSymm); compute runs
in librotmv_sm100.so (hand-written CUDA: tcgen05/TMEM/TMA implicit GEMM + HBM-bound fusion kernels).
"""
from . import _lib  # noqa: F401
from . import functional  # noqa: F401
from .module import FeatRotationSymm  # noqa: F401

__all__ = ["FeatRotationSymm", "functional", "_lib"]

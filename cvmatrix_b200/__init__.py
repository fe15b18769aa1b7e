"""cvmatrix_b200: B200-native fold-wise training-matrix engine with the CVMatrix / Partitioner API of
sm00thix/cvmatrix (reference cvmatrix/__init__.py:1-4).  CUDA only - there is no CPU fallback."""

from .cvmatrix import CVMatrix
from .partitioner import Partitioner

__version__ = "0.1.0"
__all__ = ["CVMatrix", "Partitioner"]

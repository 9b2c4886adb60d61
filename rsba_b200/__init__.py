"""rsba_b200 -- B200-native (sm_100a) rolling-shutter bundle-adjustment inner loop.

Drop-in for the Evaluator + LinearSolver that henrique/rsba obtains from Ceres
(``CeresHandler.h:245-255, 394-426``).  The product is the C-ABI shared library
``rsba_b200/lib/librsba_cuda.so`` (``include/rsba_cuda.h``); this Python package is only
the ctypes mirror used by tests and ``bench.py`` plus the synthetic scene generator.
"""
from .scene import Scene, make_scene, make_config, CONFIGS  # noqa: F401

__all__ = ["Scene", "make_scene", "make_config", "CONFIGS"]

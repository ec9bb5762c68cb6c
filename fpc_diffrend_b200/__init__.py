"""fpc_diffrend_b200 — B200-native (sm_100a) fit hot path of fpc-diffrend.

    import fpc_diffrend_b200.ops as dr      # drop-in for `import nvdiffrast.torch as dr` (reference fit.py:13)
    from fpc_diffrend_b200.fit import FitSession, FitConfig

The compute path is libfpc_b200.so (hand-written CUDA behind the C-ABI of include/fpc_b200.h); importing the
package does not need a GPU, calling any op does — there is no CPU or PyTorch fallback.
"""
from . import _lib, camera, rig  # noqa: F401

__all__ = ['ops', 'fit', 'camera', 'rig', '_lib', 'build']

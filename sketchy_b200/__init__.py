"""sketchy_b200 — B200 (sm_100a) implementation of sketchy's MinHash hot path behind a C ABI.

The package holds only what the path needs: ``csrc/`` (CUDA kernels + the C ABI, built into
``libsketchy_b200.so``), the ctypes binding (``_lib``) and the host-side mirror of the reference's
``Sketchy`` interface (``api``). There is no CPU fallback: without the built extension and a B200 every
compute call raises.
"""
from .api import PredictConfig, Sketchy, SketchyError, SkbError  # noqa: F401
from ._lib import Context, Batch, load_library  # noqa: F401

__version__ = "0.1.0"

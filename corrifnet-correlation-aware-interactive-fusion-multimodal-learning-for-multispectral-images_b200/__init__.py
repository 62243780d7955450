"""corrif_b200: B200-native (sm_100a) kernels for CorrIFNet's correlation-aware interactive fusion
hot path (reference mmvit4.py:456-529, F5_JACCARD2.py, F4_TRAIN.py:52-71).

The directory name is fixed by the project layout and is not a Python identifier; import it through
the ``corrif_b200`` alias package at the repo root.

There is no CPU fallback and no alternative backend: compute entry points raise when
libcorrif_b200.so is missing or the device is not compute capability 10.x.
"""
from . import _lib  # noqa: F401
from ._lib import CorrifError, LIB_PATH  # noqa: F401

__all__ = ["CorrifError", "LIB_PATH", "build_library", "library_loaded"]


def build_library(force: bool = False, verbose: bool = False) -> str:
    from .build import build_library as _b
    return _b(force=force, verbose=verbose)


def library_loaded() -> bool:
    return _lib._lib is not None

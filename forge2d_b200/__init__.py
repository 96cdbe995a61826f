"""forge2d_b200 — B200-native implementation of everything behind forge2d's ``World.step``.

The product is the C-ABI shared library ``forge2d_b200/csrc/libforge2d_b200.so`` (CUDA, sm_100a) declared in
``include/forge2d_b200.h``; this package is a thin ctypes loader plus a Python mirror of forge2d's Dart API for the
step path (``forge2d_b200.api``) and the synthetic scene generators of the benchmark configurations.
"""
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.path.join(_HERE, "csrc", "libforge2d_b200.so")
_lib = None


def load_library():
    """Loads the CUDA library. Raises loudly when it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBRARY_PATH):
            raise RuntimeError(
                "forge2d_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % LIBRARY_PATH)
        _lib = _abi.Library(LIBRARY_PATH, "product")
        if _lib.missing:
            raise RuntimeError("forge2d_b200: library lacks symbols: %s" % _lib.missing)
    return _lib

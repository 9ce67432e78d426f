"""n-body-simulation_b200: B200-native gravity hot path (naive all-pairs + Barnes-Hut) behind a C ABI.

The product is libnbody_b200.so (csrc/, hand-written sm_100a CUDA) and the host executable (host/, C++ mirror of the
reference's driver).  This Python package only holds the build script, a ctypes binding used by tests / bench, and the
synthetic body generators.  Import with importlib (the directory name is not a Python identifier):

    nb = importlib.import_module("n-body-simulation_b200")
"""
from . import build as _build  # noqa: F401
from .binding import (Context, NBodyError, default_config, library_path, load_library, slice_bounds,  # noqa: F401
                      comm_unique_id, TIMER_NAMES)
from . import generators  # noqa: F401


def build(force=False, verbose=False):
    """Compile libnbody_b200.so (+ the host executable when its sources exist)."""
    return _build.build_all(force=force, verbose=verbose)

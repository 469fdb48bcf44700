"""patmatchdocker_b200 -- B200-native drop-in for the PatMatch search path.

Host side (this package) mirrors the reference's www/FlaskApp/FlaskApp/patmatch.py
for the search request; the search itself runs in hand-written sm_100a CUDA kernels
behind the C ABI declared in include/patmatch_b200.h (lib/libpatmatch_b200.so).
There is no CPU fallback: importing the native layer without the built library, or
creating an engine without a CUDA device, raises.
"""
from .pattern import convert, reverse_complement, PatternError            # noqa: F401
from ._native import Engine, Dataset, NativeError, plan, lib_path, load, set_compat_deployed_glibc, jit_wait    # noqa: F401
from . import patmatch                                                     # noqa: F401

__all__ = ["Engine", "Dataset", "NativeError", "plan", "convert", "reverse_complement",
           "PatternError", "patmatch", "lib_path", "load"]

"""debigulator_b200 -- B200-native batched inflate / gzip / PNG decode.

The product is the C-ABI shared library built from csrc/ (hand-written sm_100a
CUDA kernels + a C/C++ host shim, see include/*.h). This Python package is only
a ctypes binding for tests, benchmarks and scripting; it has no decode logic of
its own and no CPU fallback: loading fails loudly if the library is missing and
`Context()` raises if no CUDA device is usable.
"""
from .api import (Context, MultiContext, Pipe, DebigulatorError, STATUS_NAMES, load_library, library_path,
                  png_get_width_height, bmp_get_width_height)
from . import api

__all__ = ["Context", "MultiContext", "Pipe", "DebigulatorError", "STATUS_NAMES", "load_library", "library_path",
           "png_get_width_height", "bmp_get_width_height", "api"]

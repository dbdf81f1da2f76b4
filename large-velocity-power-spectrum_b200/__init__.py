"""B200-native particles -> P(k) hot path of vpower (YujieH3/large-velocity-power-spectrum).

Layout
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/vpower_b200.h) -> libvpower_b200.so
  vpower/    host-side mirror of the reference call surface (vpower.interp, vpower.spctrm) over ctypes
  scripts/   drop-in for scripts/parallel_optimized.py (same command line)

The directory name is not a Python identifier; import it with
    importlib.import_module("large-velocity-power-spectrum_b200")
or put it on sys.path and `import vpower`.
"""
import os
import sys

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
if PACKAGE_DIR not in sys.path:
    sys.path.insert(0, PACKAGE_DIR)

import vpower  # noqa: E402,F401
from vpower import _lib, interp, spctrm  # noqa: E402,F401

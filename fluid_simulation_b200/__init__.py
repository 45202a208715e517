"""B200-native PIC/FLIP hot path of kbladin/Fluid_Simulation.

The product is `lib/libfsb.so` (hand-written CUDA for sm_100a behind the C ABI
of include/fsb.h) and the C++ host classes of include/fsb/.  This Python
package is only harness glue for tests and bench.py: a ctypes binding.  There
is no CPU fallback; importing `capi` fails loudly when the library is missing.
"""
from . import capi  # noqa: F401

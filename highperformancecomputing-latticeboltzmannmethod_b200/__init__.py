"""B200-native D2Q9 BGK collide-stream engine behind the LBMSolver / LBMGrid / LBMIO interface.

Contents: csrc/ (sm_100a CUDA kernels + the C-ABI library liblbm_b200.so) and binding.py (ctypes,
for tests and bench.py).  The C++ drop-in headers live in ../include.
"""
from .binding import *  # noqa: F401,F403
from .binding import Solver, SimulationParams, LbmError, load, LIB_PATH  # noqa: F401
from . import slabs  # noqa: F401,E402
from .slabs import Slab, create_slab_solver  # noqa: F401,E402

"""robio_2021_dexopt_b200 -- B200-native execution engine for tractor/dexopt's recorded-objective hot path.

Only what the path needs lives here:
    csrc/        CUDA kernels, tape lowering and the C ABI (include/trx_b200.h)
    engine.py    host mirror of tractor's Engine / Executable / Memory interface
    trxfile.py   reader for serialised tape programs (include/trx_file.h)
"""
from .engine import CudaEngine, Executable, Memory, kernel_launch_count  # noqa: F401
from ._native import TrxError  # noqa: F401
from . import trxfile  # noqa: F401
from . import collision  # noqa: F401
from .builder import build_dexsyn  # noqa: F401
from .solver import GradientDescentSolver, LeastSquaresSolver, Objective, RolloutObjective  # noqa: F401

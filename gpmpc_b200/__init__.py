"""Importable alias of the package directory `gaussian-process-mpc_b200/` (hyphens are not valid in a
module name): `import gpmpc_b200` loads the modules from that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "gaussian-process-mpc_b200"))
from . import _lib, backend  # noqa: E402,F401
from ._lib import GpmpcError, LIB_PATH, SIGNATURES  # noqa: E402,F401
from .gpr import GaussianProcessRegression  # noqa: E402,F401
from .dynamics import Dynamics  # noqa: E402,F401
from .mpc import RiskSensitiveMPC  # noqa: E402,F401
from .batched import BatchedRollouts, BatchedSolver, BatchedSimulator, shard_range, shard_indices  # noqa: E402,F401

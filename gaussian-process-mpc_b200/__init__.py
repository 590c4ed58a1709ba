"""gaussian-process-mpc on B200: drop-in replacements for the reference's hot-path classes, backed by the
hand-written sm_100a library libgpmpc.so (include/gpmpc.h).  Import as `gpmpc_b200` (the directory name
contains hyphens)."""
from ._lib import GpmpcError, LIB_PATH, SIGNATURES  # noqa: F401
from .gpr import GaussianProcessRegression  # noqa: F401
from .dynamics import Dynamics  # noqa: F401
from .mpc import RiskSensitiveMPC  # noqa: F401
from .batched import BatchedRollouts, BatchedSolver, BatchedSimulator, shard_range, shard_indices  # noqa: F401

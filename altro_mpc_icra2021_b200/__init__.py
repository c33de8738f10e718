"""altro_mpc_icra2021_b200 -- B200-native batched ALTRO (AL-iLQR) solve path.

Host-side mirror of the Julia API the reference drives (Problem / ConstraintList / ALTROSolver /
SolverOptions / solve!) over a C-ABI CUDA library (include/altro_b200.h).
"""
from .problem import *  # noqa: F401,F403
from . import problems  # noqa: F401

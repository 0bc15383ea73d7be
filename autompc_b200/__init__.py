"""autompc_b200 -- B200-native MPC solve engine behind AutoMPC's plugin API.

Controllers (`MPPI`, `IterativeLQR`) and the dynamics model (`B200MLP`) subclass
the reference's `Controller` / `Model` ABCs and call hand-written sm_100a CUDA
through the C ABI in include/ampc_b200.h (libampc_b200.so).  No CPU fallback.
"""
from .mlp import B200MLP, MLPWeights  # noqa: F401
from .mppi import MPPI, MPPIFactory  # noqa: F401
from .ilqr import IterativeLQR, IterativeLQRFactory  # noqa: F401
from .closed_loop import simulate, evaluate_candidates  # noqa: F401
from .evaluation import get_model_rmse  # noqa: F401
from .linear import B200Linear  # noqa: F401
from .nmpc import NonLinearMPCProblem  # noqa: F401

__all__ = ["MPPI", "MPPIFactory", "IterativeLQR", "IterativeLQRFactory", "B200MLP", "MLPWeights", "simulate",
           "evaluate_candidates", "get_model_rmse", "B200Linear", "NonLinearMPCProblem"]

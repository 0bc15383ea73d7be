"""Shared helpers for the ``-m gpu`` parity tests: build engine-side objects
(``autompc_b200``) from the same plain data the oracle uses."""
import numpy as np

from autompc_b200 import B200MLP, MLPWeights
from autompc_b200.plugin import QuadCost, System, Task


def weights_of(p):
    """oracle ``MLPParams`` -> engine ``MLPWeights`` (same float64 arrays)."""
    return MLPWeights(p.weights, p.biases, p.act, p.xu_mean, p.xu_std, p.dy_mean, p.dy_std, p.nx, p.nu)


def problem_of(p, cost, umin, umax, dt=0.05):
    """(system, task, model) on the plugin surface for oracle-side (MLPParams, QuadCostParams, bounds)."""
    system = System(["x%d" % i for i in range(p.nx)], ["u%d" % i for i in range(p.nu)])
    system.dt = dt
    task = Task(system)
    task.set_ctrl_bounds(np.asarray(umin, dtype=np.float64), np.asarray(umax, dtype=np.float64))
    task.set_cost(QuadCost(system, cost.Q, cost.R, cost.F, goal=cost.goal))
    model = B200MLP(system, weights_of(p))
    return system, task, model


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300)))

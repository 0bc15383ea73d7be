"""Synthetic problem definitions for the configurations BASELINE.json names
(SURVEY.md 8(d)): plain data, no reference code and no oracle."""
import numpy as np

from .mlp import MLPWeights
from .plugin import QuadCost, System, Task


def torch_default_mlp(nx, nu, hidden, act="relu", seed=100):
    """Random-init MLP with torch.nn.Linear's default init under ``torch.manual_seed(seed)``
    (what an untrained reference MLP holds, mlp.py:160-161) and the synthetic normalisers of
    SURVEY.md 8(d): xu_mean~N(0,1), xu_std~U(0.5,2), dy_mean~0.01 N(0,1), dy_std~U(0.01,0.1)."""
    import torch
    torch.manual_seed(seed)
    dims = [nx + nu] + list(hidden) + [nx]
    W, b = [], []
    for i in range(len(dims) - 1):
        lin = torch.nn.Linear(dims[i], dims[i + 1]).double()
        W.append(lin.weight.detach().numpy().copy())
        b.append(lin.bias.detach().numpy().copy())
    rng = np.random.default_rng(0)
    return MLPWeights(W, b, act, rng.normal(size=nx + nu), rng.uniform(0.5, 2.0, size=nx + nu),
                      0.01 * rng.normal(size=nx), rng.uniform(0.01, 0.1, size=nx), nx, nu)


def halfcheetah_dim_problem():
    """C3: nx=17, nu=6, u in [-1,1]^6, MLP[3x256] ReLU, Q=I, R=0.01 I, F=10 I, goal 0."""
    nx, nu = 17, 6
    system = System(["o%d" % i for i in range(nx)], ["u%d" % i for i in range(nu)])
    system.dt = 0.05
    task = Task(system)
    for name in system.controls:
        task.set_ctrl_bound(name, -1.0, 1.0)
    task.set_cost(QuadCost(system, np.eye(nx), 0.01 * np.eye(nu), 10.0 * np.eye(nx), goal=np.zeros(nx)))
    weights = torch_default_mlp(nx, nu, [256, 256, 256])
    x0 = np.random.default_rng(0).normal(size=nx)
    return system, task, weights, x0


def cartpole_problem(weights=None):
    """C1/C2/C4 dims: nx=4, nu=1, u in [-20,20], dt 0.05, x0=[3.1,0,0,0]; QuadCost of
    examples/3_Controllers_and_Tasks.ipynb cell 6.  ``weights``: a trained 2x64 MLP
    (tests use tests/golden/cartpole_mlp.npz); random-init otherwise."""
    system = System(["theta", "omega", "x", "dx"], ["u"])
    system.dt = 0.05
    task = Task(system)
    task.set_ctrl_bound("u", -20.0, 20.0)
    task.set_cost(QuadCost(system, np.diag([30.0, 3.0, 0.005, 0.1]), np.diag([0.02]),
                           np.diag([2.0, 3000.0, 0.15, 0.3]), goal=np.zeros(4)))
    if weights is None:
        weights = torch_default_mlp(4, 1, [64, 64])
    x0 = np.array([3.1, 0.0, 0.0, 0.0])
    return system, task, weights, x0

"""Shared test helpers (oracle side)."""
import os

import numpy as np

from oracle.mppi_oracle import MLPParams, MPPIOracle, QuadCostParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cartpole():
    z = np.load(os.path.join(GOLDEN, "cartpole_mlp.npz"))
    mlp = MLPParams.from_npz(z)
    cost = QuadCostParams(z["Q"], z["R"], z["F"], z["goal"])
    return mlp, cost, z["umin"], z["umax"], z["x0"], float(z["dt"])


def cartpole_step(y, u, dt=0.05, g=9.8, m=1.0, L=1.0, b=1.0):
    theta, omega, x, dx = y
    f = np.array([omega, g * np.sin(theta) / L - b * omega / (m * L ** 2) + u[0] * np.cos(theta) / L, dx, u[0]])
    return y + dt * f


def synthetic_mlp(nx, nu, hidden, act="relu", seed=0, scale=1.0):
    """Random MLP in the same spirit as SURVEY.md 8(d) (numpy-only, test use)."""
    rng = np.random.default_rng(seed)
    dims = [nx + nu] + list(hidden) + [nx]
    ws, bs = [], []
    for i in range(len(dims) - 1):
        bound = 1.0 / np.sqrt(dims[i])
        ws.append(rng.uniform(-bound, bound, size=(dims[i + 1], dims[i])) * scale)
        bs.append(rng.uniform(-bound, bound, size=dims[i + 1]))
    return MLPParams(ws, bs, act, rng.normal(size=nx + nu), rng.uniform(0.5, 2.0, size=nx + nu),
                     0.01 * rng.normal(size=nx), rng.uniform(0.01, 0.1, size=nx), nx, nu)

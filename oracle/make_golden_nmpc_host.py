"""Golden fixture for the HOST-side callbacks of the direct-transcription problem, from the UNMODIFIED reference --
TEST INFRASTRUCTURE.

    python -m oracle.make_golden_nmpc_host          (build container only: needs /root/reference)

``nmpc_host_cartpole_H8.npz``: ``NonLinearMPCProblem.get_cost`` / ``get_gradient`` (autompc/control/nmpc.py:72-100),
``get_constr_bounds`` / ``get_variable_bounds`` (:112-131) for the cartpole problem with a dense quadratic cost whose
goal is NOT the origin (the terminal gradient of the reference ignores the goal, autompc/costs/cost.py:194-199 -- the
engine's host code has to reproduce that, so the fixture must see it), observation bounds set on two dimensions.
The model only enters through ``state_dim``; the callbacks that run it are pinned by ``nmpc_cartpole_H8.npz``.
"""
import os
import sys

import numpy as np
import torch

from . import ref_loader
from .make_golden import GOLD, make_cartpole
from .make_golden_f import SUM_F2, SUM_G2, SUM_Q2, reference_mlp_from_npz
from .make_golden_r2 import load_extra

R2 = np.array([[0.07]])
OBS_BOUNDS = {"theta": (-4.0, 4.5), "x": (-7.0, 6.0)}


def main():
    ns = load_extra(ref_loader.load())
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "cartpole_mlp.npz"))
    system, task = make_cartpole(ns)
    mlp = reference_mlp_from_npz(ns, system, z)
    task.set_cost(ns.QuadCost(system, SUM_Q2, R2, SUM_F2, goal=SUM_G2))
    for name, (lo, hi) in OBS_BOUNDS.items():
        task.set_obs_bound(name, lo, hi)
    H = 8
    np.random.seed(9)
    prob = ns.NonLinearMPCProblem(system, mlp, task, H)
    rng = np.random.default_rng(17)
    x = rng.normal(size=prob.dimx)
    x[(H + 1) * 4:] *= 4.0
    xlb, xub = prob.get_variable_bounds()
    clb, cub = prob.get_constr_bounds()
    out = {"H": H, "x": x, "Q": SUM_Q2, "R": R2, "F": SUM_F2, "goal": SUM_G2, "dt": system.dt,
           "obs_bounds": np.asarray(task.get_obs_bounds(), dtype=np.float64),
           "cost": float(prob.get_cost(x)), "gradient": prob.get_gradient(x).copy(),
           "xlb": xlb, "xub": xub, "clb": clb, "cub": cub}
    path = os.path.join(GOLD, "nmpc_host_cartpole_H8.npz")
    np.savez_compressed(path, **out)
    print("  %-44s %8d B   cost %.12g" % (os.path.basename(path), os.path.getsize(path), out["cost"]))


if __name__ == "__main__":
    sys.exit(main())

"""Round-2 golden fixtures from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

    python -m oracle.make_golden_r2          (build container only: needs /root/reference)

* ``mppi_cartpole_thresh_K256_H20.npz``: reference ``MPPI`` driven by ``QuadCost + ThresholdCost + BoxThresholdCost``
  (a ``SumCost``, evaluated term by term: autompc/costs/sum_cost.py:52-81, thresh_cost.py:27-32, :73-77);
  ``mppi_cartpole_threshonly_K128_H15.npz``: the same with a lone ``ThresholdCost`` as the task cost.
* ``cost_call_thresh.npz``: ``Cost.__call__`` (cost.py:27-41) of those costs on a recorded trajectory.
* ``linear_models.npz``: ``ARX`` (trained by the reference's own ``ARX.train``) and ``Koopman`` (parameters injected with
  ``set_parameters``): ``pred_batch`` / ``pred`` / ``traj_to_state`` (arx.py:151-154, koopman.py:170-173).
* ``nmpc_cartpole_H8.npz``: ``NonLinearMPCProblem.get_constraint`` / ``get_jacobian`` (nmpc.py:102-110, :148-187) with
  the cartpole MLP.

``nmpc.py`` does ``from collections import Iterable`` (removed in Python 3.10) and ``stable_koopman.py`` imports
``scipy.linalg.pinv2`` (removed in SciPy 1.9): both names are aliased on the live modules before the import; the
reference files are executed as they are.
"""
import collections
import collections.abc
import importlib
import os
import sys

import numpy as np
import scipy.linalg
import torch

from . import ref_loader
from .make_golden import CART_F, CART_Q, CART_R, GOLD, gen_trajs, golden_mppi, make_cartpole
from .make_golden_f import reference_mlp_from_npz

THRESH = dict(goal=np.zeros(4), obs_range=(0, 2), threshold=0.55)
BOX_LIMITS = np.array([[-np.inf, np.inf], [-1.2, 1.2], [-0.45, 0.1], [-np.inf, 0.9]])


def load_extra(ns):
    """thresh_cost, arx, koopman, nmpc from the reference tree (after ref_loader.load())."""
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable          # nmpc.py:1
    if not hasattr(scipy.linalg, "pinv2"):
        scipy.linalg.pinv2 = scipy.linalg.pinv                   # stable_koopman.py:8
    with ref_loader.quiet():
        tc = importlib.import_module("autompc.costs.thresh_cost")
        ns.ThresholdCost, ns.BoxThresholdCost = tc.ThresholdCost, tc.BoxThresholdCost
        ns.ARX = importlib.import_module("autompc.sysid.arx").ARX
        ns.Koopman = importlib.import_module("autompc.sysid.koopman").Koopman
        ns.NonLinearMPCProblem = importlib.import_module("autompc.control.nmpc").NonLinearMPCProblem
    return ns


def main():
    ns = load_extra(ref_loader.load())
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "cartpole_mlp.npz"))
    system, task = make_cartpole(ns)
    mlp = reference_mlp_from_npz(ns, system, z)
    quad = ns.QuadCost(system, CART_Q, CART_R, np.diag([2.0, 30.0, 0.15, 0.3]), goal=np.zeros(4))
    thr = ns.ThresholdCost(system, THRESH["goal"], THRESH["obs_range"], THRESH["threshold"])
    box = ns.BoxThresholdCost(system, BOX_LIMITS)
    # --- MPPI with threshold terms, through the reference's own MPPI
    task.set_cost(quad + thr + box)
    g = golden_mppi(ns, system, task, mlp, K=256, H=20, seed=21, n_steps=3, sigma=0.6, lmda=0.8,
                    x_init=[0.5, 0.2, -0.4, 0.1])
    g.update(Q=CART_Q, R=CART_R, F=np.diag([2.0, 30.0, 0.15, 0.3]), thr_goal=THRESH["goal"],
             thr_range=np.array(THRESH["obs_range"]), thr_threshold=THRESH["threshold"], box_limits=BOX_LIMITS)
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_thresh_K256_H20.npz"), **g)
    task.set_cost(thr)
    g = golden_mppi(ns, system, task, mlp, K=128, H=15, seed=22, n_steps=3, sigma=0.8, lmda=0.5,
                    x_init=[0.45, 0.3, 0.0, 0.0])
    g.update(thr_goal=THRESH["goal"], thr_range=np.array(THRESH["obs_range"]), thr_threshold=THRESH["threshold"])
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_threshonly_K128_H15.npz"), **g)
    # --- Cost.__call__ on a trajectory
    traj = gen_trajs(ns, system, 1, 40, seed=5)[0]
    traj.obs[:, 0] = 0.9 * np.sin(np.linspace(0, 6, 40))          # wander in and out of the threshold / the box
    traj.obs[:, 1] = 1.5 * np.cos(np.linspace(0, 9, 40))
    traj.obs[:, 2] = np.linspace(-0.6, 0.3, 40)
    traj.obs[:, 3] = np.linspace(-1.0, 1.2, 40)
    out = {"obs": traj.obs.copy(), "ctrls": traj.ctrls.copy(), "call_thr": thr(traj), "call_box": box(traj),
           "call_quad": quad(traj), "call_sum": (quad + thr + box)(traj)}
    out.update(Q=CART_Q, R=CART_R, F=np.diag([2.0, 30.0, 0.15, 0.3]), thr_goal=THRESH["goal"],
               thr_range=np.array(THRESH["obs_range"]), thr_threshold=THRESH["threshold"], box_limits=BOX_LIMITS)
    np.savez_compressed(os.path.join(GOLD, "cost_call_thresh.npz"), **out)
    # --- linear models
    trajs = gen_trajs(ns, system, 8, 60, seed=9)
    arx = ns.ARX(system, history=3)
    arx.train(trajs)
    rng = np.random.default_rng(3)
    Xa, Ua = rng.normal(size=(9, arx.state_dim)), rng.normal(size=(9, 1)) * 5.0
    lin = {"arx_A": arx.A, "arx_B": arx.B, "arx_history": 3, "arx_X": Xa, "arx_U": Ua,
           "arx_pred_batch": arx.pred_batch(Xa, Ua), "arx_pred0": arx.pred(Xa[0], Ua[0]),
           "arx_traj_obs": trajs[0].obs[:7].copy(), "arx_traj_ctrls": trajs[0].ctrls[:7].copy(),
           "arx_traj_to_state": arx.traj_to_state(trajs[0][:7])}
    st = arx.traj_to_state(trajs[0][:7])
    lin["arx_update_state"] = arx.update_state(st, trajs[0][6].ctrl, trajs[0][7].obs)
    with ref_loader.quiet():
        koop = ns.Koopman(system, method="lstsq", poly_basis=True, poly_degree=3)
    nk = koop.state_dim
    koop.set_parameters({"A": rng.normal(size=(nk, nk)) / np.sqrt(nk), "B": rng.normal(size=(nk, 1))})
    Xk, Uk = rng.normal(size=(11, nk)), rng.normal(size=(11, 1))
    lin.update(koop_A=koop.A, koop_B=koop.B, koop_X=Xk, koop_U=Uk, koop_pred_batch=koop.pred_batch(Xk, Uk),
               koop_pred0=koop.pred(Xk[0], Uk[0]), koop_state_dim=nk)
    np.savez_compressed(os.path.join(GOLD, "linear_models.npz"), **lin)
    # --- NMPC callbacks
    H = 8
    task.set_cost(quad)
    np.random.seed(4)
    prob = ns.NonLinearMPCProblem(system, mlp, task, H)
    x = rng.normal(size=prob.dimx)
    x[(H + 1) * 4:] *= 4.0
    row, col = prob.get_jacobian(x, True)
    nm = {"H": H, "x": x, "constraint": prob.get_constraint(x).copy(), "row": row, "col": col,
          "jac": prob.get_jacobian(x, False).copy(), "dimx": prob.dimx, "dimc": prob.dimc}
    np.savez_compressed(os.path.join(GOLD, "nmpc_cartpole_H8.npz"), **nm)
    for f in ("mppi_cartpole_thresh_K256_H20.npz", "mppi_cartpole_threshonly_K128_H15.npz", "cost_call_thresh.npz",
              "linear_models.npz", "nmpc_cartpole_H8.npz"):
        print("  %-44s %8d B" % (f, os.path.getsize(os.path.join(GOLD, f))))
    print({k: float(v) for k, v in out.items() if k.startswith("call_")})


if __name__ == "__main__":
    sys.exit(main())

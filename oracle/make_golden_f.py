"""Golden fixtures for the SURVEY.md section 8(f) rows, from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

    python -m oracle.make_golden_f          (build container only: needs /root/reference)

* ``mppi_cartpole_sumcost_K256_H20.npz``: the reference ``MPPI`` driven by a ``SumCost`` of two ``QuadCost`` terms
  with DIFFERENT goals (``autompc/costs/sum_cost.py``; such a sum is not ``is_quad``, the reference evaluates it
  term by term through ``eval_obs_cost`` / ``eval_ctrl_cost`` / ``eval_term_obs_cost``).
* ``model_rmse_cartpole.npz``: ``autompc.evaluation.model_metrics.get_model_rmse`` (``model_metrics.py:12-43``) of the
  trained cartpole MLP on fresh trajectories at several horizons.

The MLP is the one frozen in ``cartpole_mlp.npz`` (trained by the reference's ``MLP.train`` in ``make_golden.py``),
re-injected into a reference ``MLP`` object.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import ref_loader
from .make_golden import CART_F, CART_Q, CART_R, GOLD, gen_trajs, golden_mppi, make_cartpole


def reference_mlp_from_npz(ns, system, z):
    n_hidden = int(z["n_layers"]) - 1
    hidden = int(z["W0"].shape[0])
    with ref_loader.quiet():
        mlp = ns.MLP(system, n_hidden_layers=n_hidden, hidden_size=hidden, nonlintype=str(z["act"]), use_cuda=False)
    sd = {}
    for i in range(n_hidden):
        sd["layers.layer%d.weight" % i] = torch.from_numpy(z["W%d" % i].copy())
        sd["layers.layer%d.bias" % i] = torch.from_numpy(z["b%d" % i].copy())
    sd["output_layer.weight"] = torch.from_numpy(z["W%d" % n_hidden].copy())
    sd["output_layer.bias"] = torch.from_numpy(z["b%d" % n_hidden].copy())
    mlp.net.load_state_dict(sd)
    mlp.net.double().eval()
    for p in mlp.net.parameters():
        p.requires_grad_(False)
    mlp.xu_means, mlp.xu_std = z["xu_mean"].copy(), z["xu_std"].copy()
    mlp.dy_means, mlp.dy_std = z["dy_mean"].copy(), z["dy_std"].copy()
    return mlp


SUM_Q2 = np.array([[2.0, 0.3, 0.0, 0.1], [0.3, 1.0, 0.2, 0.0], [0.0, 0.2, 0.5, 0.1], [0.1, 0.0, 0.1, 0.4]])
SUM_F2 = np.diag([5.0, 1.0, 0.0, 2.0])
SUM_R2 = np.diag([0.05])
SUM_G2 = np.array([0.4, -0.3, 0.8, 0.1])


def main():
    ns = ref_loader.load()
    torch.set_num_threads(1)
    z = np.load(os.path.join(GOLD, "cartpole_mlp.npz"))
    system, task = make_cartpole(ns)
    mlp = reference_mlp_from_npz(ns, system, z)
    # --- SumCost of two quadratics with different goals, through the reference's own MPPI
    c1 = ns.QuadCost(system, CART_Q, CART_R, CART_F, goal=np.zeros(4))
    c2 = ns.QuadCost(system, SUM_Q2, SUM_R2, SUM_F2, goal=SUM_G2)
    task.set_cost(c1 + c2)
    assert isinstance(task.get_cost(), ns.SumCost) and not task.get_cost().is_quad
    g = golden_mppi(ns, system, task, mlp, K=256, H=20, seed=11, n_steps=3, sigma=0.6, lmda=0.8,
                    x_init=[0.5, 0.2, -0.4, 0.1])
    g.update(Q1=CART_Q, R1=CART_R, F1=CART_F, g1=np.zeros(4), Q2=SUM_Q2, R2=SUM_R2, F2=SUM_F2, g2=SUM_G2)
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_sumcost_K256_H20.npz"), **g)
    # --- k-step model error (get_model_rmse) through the reference's own evaluation code
    ev = types.ModuleType("autompc.evaluation")
    ev.__path__ = [os.path.join(ref_loader.REF_ROOT, "autompc", "evaluation")]
    sys.modules["autompc.evaluation"] = ev
    mm = importlib.import_module("autompc.evaluation.model_metrics")
    trajs = gen_trajs(ns, system, 6, 60, seed=7)
    out = {"obs": np.stack([t.obs for t in trajs]), "ctrls": np.stack([t.ctrls for t in trajs]),
           "horizons": np.array([1, 3, 10, 25])}
    for h in out["horizons"]:
        out["rmse_h%d" % h] = mm.get_model_rmse(mlp, trajs, horizon=int(h))
    # ragged set: trajectories of different lengths
    ragged = [trajs[0][:60], trajs[1][:33], trajs[2][:12]]
    out["ragged_lens"] = np.array([60, 33, 12])
    out["rmse_ragged_h5"] = mm.get_model_rmse(mlp, ragged, horizon=5)
    np.savez_compressed(os.path.join(GOLD, "model_rmse_cartpole.npz"), **out)
    for f in ("mppi_cartpole_sumcost_K256_H20.npz", "model_rmse_cartpole.npz"):
        print("  %-44s %8d B" % (f, os.path.getsize(os.path.join(GOLD, f))))
    print({k: float(v) for k, v in out.items() if k.startswith("rmse")})


if __name__ == "__main__":
    sys.exit(main())

import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
import tests.test_thresh_gpu as T
import pytest
# monkeypatch assert-free measurement: re-run the synthetic test body with prints
from autompc_b200 import MPPI, B200MLP
from autompc_b200.plugin import BoxThresholdCost, QuadCost, System, Task, ThresholdCost
from oracle.mppi_oracle import MPPIOracle, SumQuadCostParams, QuadCostParams, BoxThresholdCostParams, ThresholdCostParams
from tests.gpu_helpers import weights_of
from tests.helpers import synthetic_mlp, GOLDEN, load_cartpole
for precision in ("fp32","fp16","bf16"):
    nx, nu, K, H = 6, 3, 600, 12
    p = synthetic_mlp(nx, nu, [64, 64], seed=8)
    rng = np.random.default_rng(3)
    A = rng.normal(size=(nx, nx))
    Q, R, F, g = A @ A.T / nx, 0.05 * np.eye(nu), 2.0 * np.eye(nx), 0.1 * rng.normal(size=nx)
    lim = np.stack([np.full(nx, -np.inf), np.full(nx, np.inf)], axis=1)
    lim[0], lim[3] = [-0.4, 0.9], [-np.inf, 0.3]
    thr_goal = 0.2 * rng.normal(size=nx)
    system = System(["x%d" % i for i in range(nx)], ["u%d" % i for i in range(nu)]); system.dt = 0.05
    task = Task(system); task.set_ctrl_bounds(-np.ones(nu), 1.5 * np.ones(nu))
    task.set_cost(QuadCost(system, Q, R, F, goal=g) + BoxThresholdCost(system, lim) + ThresholdCost(system, thr_goal, [1, 4], 0.8))
    ocost = SumQuadCostParams([QuadCostParams(Q, R, F, g), BoxThresholdCostParams(lim), ThresholdCostParams(thr_goal, (1, 4), 0.8)])
    np.random.seed(2)
    ctl = MPPI(system, task, B200MLP(system, weights_of(p)), horizon=H, num_path=K, sigma=0.7, lmda=1.3, noise="numpy", precision=precision)
    np.random.seed(2)
    o = MPPIOracle(p, ocost, -np.ones(nu), 1.5 * np.ones(nu), horizon=H, num_path=K, sigma=0.7, lmda=1.3)
    x0 = 0.3 * rng.normal(size=nx)
    eps = o.sample_eps(); ctl.act_sequence = o.act_sequence
    u = ctl.solve(x0, eps=eps); uo = o.solve(x0, eps=eps.copy())
    costs, _ = ctl.last_costs(); ref = o.last_costs - o.term_const
    d = costs - ref
    off = np.abs(d) > T.COST_RTOL[precision] * np.abs(ref).max()
    nonflip = np.abs(d - np.round(d))
    print(precision, "flip frac", off.mean(), "limit", T.MAX_FLIP_FRAC[precision], "max nonint", nonflip.max(), "tol", T.COST_RTOL[precision]*np.abs(ref).max(), "act err", np.abs(ctl.act_sequence-o.act_sequence).max())
    ctl.close()
# fixtures
for name, kind in [("mppi_cartpole_thresh_K256_H20", "sum"), ("mppi_cartpole_threshonly_K128_H15", "lone")]:
    for precision in ("fp32", "fp16"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        system, task, model = T._problem(z, kind)
        np.random.seed(int(z["seed"]))
        ctl = MPPI(system, task, model, horizon=int(z["H"]), num_path=int(z["K"]), sigma=float(z["sigma"]), lmda=float(z["lmda"]), noise="numpy", precision=precision)
        constate = np.zeros(5); worst = 0
        for s in range(int(z["n_steps"])):
            if s > 0: ctl.act_sequence = z["act_%d" % (s - 1)]
            u, constate = ctl.run(constate, z["x0_%d" % s])
            costs, term = ctl.last_costs(); ref = z["costs_%d" % s]
            scale = max(np.abs(ref).max(), 1.0)
            d = (costs - costs.min()) - (ref - ref.min())
            off = np.abs(d) > T.COST_RTOL[precision] * scale
            worst = max(worst, off.mean())
            nz = np.abs(d[~off]).max() / scale
        print(name, precision, "worst flip frac", worst, "limit", T.MAX_FLIP_FRAC[precision], "max rel diff of non-flipped", nz, "tol", T.COST_RTOL[precision])
        ctl.close()

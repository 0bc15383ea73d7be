"""GPU parity: the device-resident closed loop (ampc_mppi_closed_loop_*) against
 (1) the same loop stepped from the host through the plugin surface (``MPPI.run`` + ``B200MLP.pred`` per step, the
     reference's ``simulate`` structure, utils/simulation.py:45-63): identical kernels, so bit-identical controls;
 (2) the float64 oracle in closed loop fed the device's Philox noise: fp32 tolerance, growing mildly with the step;
 (3) ``Cost.__call__`` (costs/cost.py:27-41) recomputed with the oracle's QuadCost on the returned trajectory."""
import os

import numpy as np
import pytest

from oracle.mppi_oracle import MLPParams, MPPIOracle, mlp_pred
from tests.helpers import GOLDEN, load_cartpole

pytestmark = pytest.mark.gpu


def _cartpole(**kw):
    from autompc_b200 import MPPI, B200MLP
    from autompc_b200.mlp import MLPWeights
    from autompc_b200.problems import cartpole_problem
    z = np.load(os.path.join(GOLDEN, "cartpole_mlp.npz"))
    system, task, w, x0 = cartpole_problem(MLPWeights.from_npz(z))
    model = B200MLP(system, w)
    np.random.seed(0)
    return MPPI(system, task, model, **kw), model, x0


def _traj_cost(cost, obs, ctrls):
    c = 0.0
    for i in range(len(obs)):                                  # cost.py:36-40
        c += cost.eval_obs_cost(obs[i]) + cost.eval_ctrl_cost(ctrls[i])
    return c + cost.eval_term_obs_cost(obs[-1])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_closed_loop_equals_host_stepped_loop(precision):
    from autompc_b200 import simulate
    T = 25
    kw = dict(horizon=15, num_path=400, seed=11, precision=precision)
    a, model, x0 = _cartpole(**kw)
    res = simulate(a, x0, sim_model=model, max_steps=T)
    b, model_b, _ = _cartpole(**kw)
    x, constate = x0.copy(), np.concatenate([x0, np.zeros(1)])
    obs, ctrls = [x.copy()], []
    for _ in range(T):
        u, constate = b.run(constate, x)
        x = model_b.pred(x, u)
        ctrls.append(u)
        obs.append(x.copy())
    # the host loop hands float64 observations to float32 kernels exactly like the device loop does
    np.testing.assert_allclose(res.ctrls[:T], np.array(ctrls), rtol=0, atol=1e-6)
    np.testing.assert_allclose(res.obs, np.array(obs), rtol=0, atol=1e-6)
    assert np.all(res.ctrls[T] == 0) and a.cur_step == T
    _, cost, _, _, _, _ = load_cartpole()
    np.testing.assert_allclose(res.cost, _traj_cost(cost, res.obs, res.ctrls), rtol=1e-12)
    a.close()
    b.close()


def test_closed_loop_matches_oracle_closed_loop():
    from autompc_b200 import simulate
    T, K, H = 8, 300, 12
    a, model, x0 = _cartpole(horizon=H, num_path=K, seed=5, precision="fp32")
    act0 = a.act_sequence
    noise = [a.philox_noise(counter=t).astype(np.float64) for t in range(T)]
    res = simulate(a, x0, sim_model=model, max_steps=T)
    mlp, cost, umin, umax, _, _ = load_cartpole()
    o = MPPIOracle(mlp, cost, umin, umax, horizon=H, num_path=K, draw_init=False)
    o.act_sequence = act0.copy()
    x = x0.copy()
    for t in range(T):
        u = o.solve(x, eps=noise[t])
        # unstable learned dynamics + softmax: the closed-loop trajectories separate slowly (fp32 vs float64)
        np.testing.assert_allclose(res.ctrls[t], u, rtol=0, atol=20.0 * 2e-3 * (1 + t))
        np.testing.assert_allclose(res.obs[t], x, rtol=0, atol=2e-3 * (1 + t))
        x = mlp_pred(mlp, x, u)
    a.close()


def test_many_closed_loops_in_flight():
    """evaluate_candidates: independent controllers with different hyper-parameters (mppi.py:50-64 ranges), all
    closed loops enqueued before any is awaited; each equals its own stand-alone run."""
    from autompc_b200 import simulate, evaluate_candidates
    rng = np.random.default_rng(100)
    cfgs = [dict(horizon=int(rng.integers(5, 31)), num_path=int(rng.integers(100, 1001)),
                 sigma=float(rng.uniform(1e-2, 2.0)), lmda=float(rng.uniform(0.1, 2.0)), seed=i, precision="auto")
            for i in range(6)]
    ctls, model, x0 = [], None, None
    for c in cfgs:
        ctl, model, x0 = _cartpole(**c)
        ctls.append(ctl)
    np.random.seed(123)
    costs, results = evaluate_candidates(ctls, x0, 20, model)
    for i, c in enumerate(cfgs):
        solo, m2, _ = _cartpole(**c)
        np.random.seed(123)                        # evaluate_candidates resets the controllers in order; each reset
        np.random.normal(size=sum(cc["horizon"] for cc in cfgs[:i]))   # draws H normals (mppi.py:99, ctrl_dim 1)
        solo.reset()
        r = simulate(solo, x0, sim_model=m2, max_steps=20)
        np.testing.assert_allclose(results[i].ctrls, r.ctrls, rtol=0, atol=1e-6)
        assert costs[i] == pytest.approx(r.cost, rel=1e-9)
        solo.close()
    for ctl in ctls:
        ctl.close()


def test_closed_loop_cost_of_folded_sumcost():
    """A SumCost of two quadratics with different goals (folded on the host): the device-side trajectory cost plus the
    fold's constants equals Cost.__call__ (cost.py:27-41) summed over the terms on the returned trajectory."""
    from autompc_b200 import MPPI, simulate
    from oracle.mppi_oracle import QuadCostParams, SumQuadCostParams
    from tests.test_mppi_gpu import _sumcost_problem
    z = np.load(os.path.join(GOLDEN, "mppi_cartpole_sumcost_K256_H20.npz"))
    system, task, model = _sumcost_problem(z)
    np.random.seed(0)
    ctl = MPPI(system, task, model, horizon=12, num_path=256, seed=3, precision="fp32")
    res = simulate(ctl, np.array([0.4, 0.0, 0.1, 0.0]), sim_model=model, max_steps=15)
    cost = SumQuadCostParams([QuadCostParams(z["Q1"], z["R1"], z["F1"], z["g1"]),
                              QuadCostParams(z["Q2"], z["R2"], z["F2"], z["g2"])])
    np.testing.assert_allclose(res.cost, _traj_cost(cost, res.obs, res.ctrls), rtol=1e-10)
    ctl.close()

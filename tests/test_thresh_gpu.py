"""GPU parity for SURVEY.md 8(f) row 2, second half: ThresholdCost / BoxThresholdCost (autompc/costs/thresh_cost.py)
as per-step predicates inside both rollout kernels and inside the closed loop's trajectory cost.

Checker: fixtures recorded from the UNMODIFIED reference MPPI (oracle/make_golden_r2.py) and the oracle's
``traj_cost`` (pinned to the reference's ``Cost.__call__`` in tests/test_oracle_golden_r2.py).

A threshold term is an indicator: a sample whose state lies within the arithmetic's rounding of a box limit may be
counted on the other side (the state is float32 on the device, float64 in the reference).  The tests therefore allow a
stated small fraction of samples to differ by (near-)integer amounts and hold all others to the usual tolerance.
"""
import os

import numpy as np
import pytest

from oracle.mppi_oracle import (BoxThresholdCostParams, QuadCostParams, SumQuadCostParams, ThresholdCostParams,
                                traj_cost)
from tests.helpers import GOLDEN, load_cartpole

pytestmark = pytest.mark.gpu

# of the cost scale, on cost differences.  The fixtures use the trained cartpole MLP, whose unstable dynamics amplify
# operand rounding along the rollout; with bf16 operands (2.5 % of the cost scale on the round-1 fixtures,
# profiles/r02_precision.jsonl) most samples that hover near a limit are counted differently at some step, so the
# fixtures are run in fp32 and fp16 only; bf16 is checked on a stable synthetic model below.
COST_RTOL = {"fp32": 2e-5, "fp16": 2e-3, "bf16": 4e-3}
ACT_ATOL = {"fp32": 2e-3, "fp16": 1e-2, "bf16": 5e-2}
MAX_FLIP_FRAC = {"fp32": 0.005, "fp16": 0.03, "bf16": 0.05}


def _problem(z, kind):
    from autompc_b200 import B200MLP
    from autompc_b200.plugin import BoxThresholdCost, QuadCost, System, Task, ThresholdCost
    from tests.gpu_helpers import weights_of
    mlp, _, umin, umax, _, _ = load_cartpole()
    system = System(["theta", "omega", "x", "dx"], ["u"])
    system.dt = 0.05
    task = Task(system)
    task.set_ctrl_bounds(np.asarray(umin, dtype=np.float64), np.asarray(umax, dtype=np.float64))
    thr = ThresholdCost(system, z["thr_goal"], [int(v) for v in z["thr_range"]], float(z["thr_threshold"]))
    if kind == "sum":
        cost = QuadCost(system, z["Q"], z["R"], z["F"], goal=np.zeros(4)) + thr + BoxThresholdCost(system, z["box_limits"])
    else:
        cost = thr
    task.set_cost(cost)
    return system, task, B200MLP(system, weights_of(mlp))


@pytest.mark.parametrize("name,kind", [("mppi_cartpole_thresh_K256_H20", "sum"),
                                       ("mppi_cartpole_threshonly_K128_H15", "lone")])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_mppi_threshold_costs_match_unmodified_reference(name, kind, precision):
    from autompc_b200 import MPPI
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    system, task, model = _problem(z, kind)
    np.random.seed(int(z["seed"]))
    ctl = MPPI(system, task, model, horizon=int(z["H"]), num_path=int(z["K"]), sigma=float(z["sigma"]),
               lmda=float(z["lmda"]), noise="numpy", precision=precision)
    assert ctl.precision == precision
    np.testing.assert_allclose(ctl.act_sequence, z["act0"], rtol=0, atol=1e-7)
    constate = np.zeros(5)
    for s in range(int(z["n_steps"])):
        if s > 0:
            ctl.act_sequence = z["act_%d" % (s - 1)]
        u, constate = ctl.run(constate, z["x0_%d" % s])
        costs, term = ctl.last_costs()
        ref = z["costs_%d" % s]
        scale = max(np.abs(ref).max(), 1.0)
        d = (costs - costs.min()) - (ref - ref.min())
        off = np.abs(d) > COST_RTOL[precision] * scale
        # samples counted on the other side of a limit differ by ~ +-1, +-2, ...: few, and (near-)integers
        assert off.mean() <= MAX_FLIP_FRAC[precision], "%.3f of the samples differ" % off.mean()
        if precision == "fp32" and off.any():
            assert np.all(np.abs(d[off] - np.round(d[off])) < 0.05)
        # The softmax weights are exp(-(c - min) / lmda) (mppi.py:115): the updated sequence can only be held to the
        # action tolerance where the cost deviations are small against lmda (the fixtures' costs are O(1e3), lmda < 1).
        if not off.any() and np.abs(d).max() < ACT_ATOL[precision] * float(z["lmda"]):
            assert int(np.argmin(costs)) == int(z["argmin_%d" % s])
            np.testing.assert_allclose(ctl.act_sequence, z["act_%d" % s], rtol=0, atol=ACT_ATOL[precision])
            np.testing.assert_allclose(u, z["u_%d" % s], rtol=0, atol=ACT_ATOL[precision] * 20.0)
    ctl.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_mppi_threshold_costs_match_oracle_synthetic(precision):
    """nu > 1, two box terms with one-sided limits, dense Q, on a synthetic (stable) MLP: all three arithmetic modes
    against the float64 oracle (restated thresh_cost.py:27-32, :73-77) on the same noise."""
    from autompc_b200 import MPPI, B200MLP
    from autompc_b200.plugin import BoxThresholdCost, QuadCost, System, Task, ThresholdCost
    from oracle.mppi_oracle import MPPIOracle
    from tests.gpu_helpers import weights_of
    from tests.helpers import synthetic_mlp
    nx, nu, K, H = 6, 3, 600, 12
    p = synthetic_mlp(nx, nu, [64, 64], seed=8)
    rng = np.random.default_rng(3)
    A = rng.normal(size=(nx, nx))
    Q, R, F, g = A @ A.T / nx, 0.05 * np.eye(nu), 2.0 * np.eye(nx), 0.1 * rng.normal(size=nx)
    lim = np.stack([np.full(nx, -np.inf), np.full(nx, np.inf)], axis=1)
    lim[0], lim[3] = [-0.4, 0.9], [-np.inf, 0.3]
    thr_goal = 0.2 * rng.normal(size=nx)
    system = System(["x%d" % i for i in range(nx)], ["u%d" % i for i in range(nu)])
    system.dt = 0.05
    task = Task(system)
    task.set_ctrl_bounds(-np.ones(nu), 1.5 * np.ones(nu))
    task.set_cost(QuadCost(system, Q, R, F, goal=g) + BoxThresholdCost(system, lim) + ThresholdCost(system, thr_goal, [1, 4], 0.8))
    ocost = SumQuadCostParams([QuadCostParams(Q, R, F, g), BoxThresholdCostParams(lim), ThresholdCostParams(thr_goal, (1, 4), 0.8)])
    np.random.seed(2)
    ctl = MPPI(system, task, B200MLP(system, weights_of(p)), horizon=H, num_path=K, sigma=0.7, lmda=1.3, noise="numpy",
               precision=precision)
    np.random.seed(2)
    o = MPPIOracle(p, ocost, -np.ones(nu), 1.5 * np.ones(nu), horizon=H, num_path=K, sigma=0.7, lmda=1.3)
    x0 = 0.3 * rng.normal(size=nx)
    eps = o.sample_eps()
    ctl.act_sequence = o.act_sequence
    u = ctl.solve(x0, eps=eps)
    uo = o.solve(x0, eps=eps.copy())
    costs, _ = ctl.last_costs()
    ref = o.last_costs - o.term_const
    d = costs - ref
    off = np.abs(d) > COST_RTOL[precision] * np.abs(ref).max()
    assert off.mean() <= MAX_FLIP_FRAC[precision], "%.3f of the samples differ" % off.mean()
    assert np.all(np.abs(d[off] - np.round(d[off])) < 0.1)               # whole violations, nothing else
    counted = (o.cost.terms[1].obs_cost_batch(np.repeat(x0[None], 2, 0)).sum() >= 0)   # oracle terms are live
    assert counted and ref.max() - ref.min() > 1.0
    if not off.any():
        np.testing.assert_allclose(ctl.act_sequence, o.act_sequence, rtol=0, atol=ACT_ATOL[precision])
        np.testing.assert_allclose(u, uo, rtol=0, atol=ACT_ATOL[precision] * 1.5)
    ctl.close()


def test_threshold_terms_are_counted_per_step():
    """One sample-level identity that holds in every arithmetic: with Q = R = F = 0 and lmda/sigma -> the action cost
    only, cost(threshold task) - cost(zero-cost task) is an integer in [0, H] for every sample."""
    from autompc_b200 import MPPI, B200MLP
    from autompc_b200.plugin import BoxThresholdCost, QuadCost, System, Task
    from tests.gpu_helpers import weights_of
    mlp, _, umin, umax, _, _ = load_cartpole()
    system = System(["theta", "omega", "x", "dx"], ["u"])
    system.dt = 0.05
    model = B200MLP(system, weights_of(mlp))
    z4, z1 = np.zeros((4, 4)), np.zeros((1, 1))
    lim = np.array([[-0.3, 0.3], [-np.inf, np.inf], [-np.inf, 0.05], [-np.inf, np.inf]])
    x0 = np.array([0.2, 0.1, 0.0, 0.0])
    out = {}
    for kind in ("zero", "box"):
        task = Task(system)
        task.set_ctrl_bounds(np.asarray(umin, dtype=np.float64), np.asarray(umax, dtype=np.float64))
        cost = QuadCost(system, z4, z1, z4, goal=np.zeros(4))
        task.set_cost(cost + BoxThresholdCost(system, lim) if kind == "box" else cost)
        np.random.seed(5)
        ctl = MPPI(system, task, model, horizon=12, num_path=700, seed=3, precision="fp32")
        ctl.solve(x0)
        out[kind] = ctl.last_costs()[0]
        ctl.close()
    d = out["box"] - out["zero"]
    assert np.all(np.abs(d - np.round(d)) < 1e-3) and d.min() >= -1e-3 and d.max() <= 12 + 1e-3 and d.max() >= 1


def test_closed_loop_eval_cost_with_threshold_terms():
    """The tuner scores a candidate with the TASK's cost (pipeline_tuner.py:230-231), e.g. the cartpole benchmark's
    ThresholdCost: ``simulate(..., cost=)`` accumulates Cost.__call__ (cost.py:27-41) of that cost on the device; it
    must equal the oracle's value on the very trajectory the closed loop returns (integers: exactly)."""
    from autompc_b200 import MPPI, simulate
    from autompc_b200.plugin import BoxThresholdCost, QuadCost, ThresholdCost
    zc = np.load(os.path.join(GOLDEN, "cost_call_thresh.npz"))
    z = {k: zc[k] for k in zc.files}
    system, task, model = _problem(z, "lone")
    # the controller optimises a quadratic; the score is something else
    task.set_cost(QuadCost(system, z["Q"], z["R"], z["F"], goal=np.zeros(4)))
    np.random.seed(1)
    ctl = MPPI(system, task, model, horizon=15, num_path=512, seed=4, precision="fp32")
    x0 = np.array([0.6, 0.0, 0.0, 0.0])
    thr = ThresholdCost(system, z["thr_goal"], [int(v) for v in z["thr_range"]], float(z["thr_threshold"]))
    box = BoxThresholdCost(system, z["box_limits"])
    quad2 = QuadCost(system, np.diag([1.0, 2.0, 3.0, 4.0]), np.diag([0.5]), np.diag([4.0, 3.0, 2.0, 1.0]), goal=np.full(4, 0.1))
    o_thr = ThresholdCostParams(z["thr_goal"], z["thr_range"], float(z["thr_threshold"]))
    o_box = BoxThresholdCostParams(z["box_limits"])
    o_quad2 = QuadCostParams(np.diag([1.0, 2.0, 3.0, 4.0]), np.diag([0.5]), np.diag([4.0, 3.0, 2.0, 1.0]), np.full(4, 0.1))
    own = QuadCostParams(z["Q"], z["R"], z["F"], np.zeros(4))
    T = 25
    for cost, ocost, exact in [(thr, o_thr, 1), (box, o_box, 1), (thr + box, SumQuadCostParams([o_thr, o_box]), 2),
                               (quad2 + thr, SumQuadCostParams([o_quad2, o_thr]), 0), (None, own, 0)]:
        np.random.seed(1)
        ctl.reset()
        r = simulate(ctl, x0, sim_model=model, max_steps=T, cost=cost)
        want = traj_cost(ocost, r.obs, r.ctrls)
        if exact:
            assert r.cost == want and 0 < want <= exact * (T + 1)      # a count of (term, step) violations
        else:
            np.testing.assert_allclose(r.cost, want, rtol=1e-12)
    ctl.close()

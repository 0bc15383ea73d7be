"""CPU: pin the round-2 oracle restatements (threshold costs, Cost.__call__, linear models, NMPC callbacks) to fixtures
produced by the UNMODIFIED reference (oracle/make_golden_r2.py), and the engine's HOST-side cost parsing to the same
reference semantics (no CUDA call is made here)."""
import os

import numpy as np

from oracle.mppi_oracle import (BoxThresholdCostParams, MPPIOracle, QuadCostParams, SumQuadCostParams,
                                ThresholdCostParams, linear_pred_batch, nmpc_constraint, nmpc_jacobian, traj_cost)
from tests.helpers import GOLDEN, load_cartpole


def thresh_terms(z):
    thr = ThresholdCostParams(z["thr_goal"], z["thr_range"], float(z["thr_threshold"]))
    box = BoxThresholdCostParams(z["box_limits"]) if "box_limits" in z.files else None
    quad = QuadCostParams(z["Q"], z["R"], z["F"], np.zeros(4)) if "Q" in z.files else None
    return quad, thr, box


def _replay(name, cost):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mlp, _, umin, umax, _, _ = load_cartpole()
    np.random.seed(int(z["seed"]))
    o = MPPIOracle(mlp, cost(z), umin, umax, horizon=int(z["H"]), num_path=int(z["K"]), sigma=float(z["sigma"]),
                   lmda=float(z["lmda"]))
    np.testing.assert_array_equal(o.act_sequence, z["act0"])
    constate = np.zeros(5)
    for s in range(int(z["n_steps"])):
        u, constate = o.run(constate, z["x0_%d" % s])
        np.testing.assert_allclose(o.last_costs, z["costs_%d" % s], rtol=1e-11, atol=1e-9)
        assert int(np.argmin(o.last_costs)) == int(z["argmin_%d" % s])
        np.testing.assert_allclose(o.act_sequence, z["act_%d" % s], rtol=0, atol=1e-9)
        np.testing.assert_allclose(u, z["u_%d" % s], rtol=0, atol=1e-8)


def test_threshold_sumcost_oracle_matches_reference():
    """Reference MPPI under QuadCost + ThresholdCost + BoxThresholdCost (thresh_cost.py:27-32, :73-77)."""
    _replay("mppi_cartpole_thresh_K256_H20", lambda z: SumQuadCostParams(list(thresh_terms(z))))


def test_threshold_only_oracle_matches_reference():
    _replay("mppi_cartpole_threshonly_K128_H15", lambda z: thresh_terms(z)[1])


def test_traj_cost_oracle_matches_reference_call():
    """Cost.__call__ (cost.py:27-41) of every cost kind on a recorded trajectory; the counts are integers: exact."""
    z = np.load(os.path.join(GOLDEN, "cost_call_thresh.npz"))
    quad, thr, box = thresh_terms(z)
    assert traj_cost(thr, z["obs"], z["ctrls"]) == float(z["call_thr"])
    assert traj_cost(box, z["obs"], z["ctrls"]) == float(z["call_box"])
    np.testing.assert_allclose(traj_cost(quad, z["obs"], z["ctrls"]), float(z["call_quad"]), rtol=1e-13)
    np.testing.assert_allclose(traj_cost(SumQuadCostParams([quad, thr, box]), z["obs"], z["ctrls"]),
                               float(z["call_sum"]), rtol=1e-13)
    # batch forms == scalar forms
    assert np.array_equal(thr.obs_cost_batch(z["obs"]), [thr.eval_obs_cost(x) for x in z["obs"]])
    assert np.array_equal(box.obs_cost_batch(z["obs"]), [box.eval_obs_cost(x) for x in z["obs"]])


def test_linear_model_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "linear_models.npz"))
    for k in ("arx", "koop"):
        A, B, X, U = z[k + "_A"], z[k + "_B"], z[k + "_X"], z[k + "_U"]
        np.testing.assert_allclose(linear_pred_batch(A, B, X, U), z[k + "_pred_batch"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(linear_pred_batch(A, B, X[:1], U[:1])[0], z[k + "_pred0"], rtol=0, atol=1e-12)


def test_nmpc_callbacks_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "nmpc_cartpole_H8.npz"))
    mlp = load_cartpole()[0]
    H = int(z["H"])
    np.testing.assert_allclose(nmpc_constraint(mlp, H, z["x"]), z["constraint"], rtol=0, atol=1e-12)
    rows, cols, vals = nmpc_jacobian(mlp, H, z["x"])
    assert np.array_equal(rows, z["row"]) and np.array_equal(cols, z["col"])          # sparsity pattern: exact
    np.testing.assert_allclose(vals, z["jac"], rtol=0, atol=1e-11)


def test_engine_cost_parsing_matches_reference_semantics():
    """autompc_b200.mppi.cost_spec_of (host side, no CUDA): threshold terms become boxes whose indicator equals the
    reference's eval_obs_cost on the recorded trajectory; quadratic terms are kept / folded."""
    from autompc_b200.mppi import cost_spec_of
    from autompc_b200.plugin import BoxThresholdCost, QuadCost, System, ThresholdCost
    z = np.load(os.path.join(GOLDEN, "cost_call_thresh.npz"))
    system = System(["theta", "omega", "x", "dx"], ["u"])
    quad = QuadCost(system, z["Q"], z["R"], z["F"], goal=np.zeros(4))
    thr = ThresholdCost(system, z["thr_goal"], list(z["thr_range"]), float(z["thr_threshold"]))
    box = BoxThresholdCost(system, z["box_limits"])
    bounds = np.array([[-20.0, 20.0]])
    spec = cost_spec_of(quad + thr + box, bounds, 4, 1)
    assert spec.quad and spec.n_box == 2 and spec.stage_const == 0.0
    obs = z["obs"]
    viol = [((obs < spec.box_lo[b]) | (obs > spec.box_hi[b])).any(axis=1) for b in range(2)]
    assert viol[0].sum() == float(z["call_thr"]) and viol[1].sum() == float(z["call_box"])
    assert np.array_equal(viol[0], [thr.eval_obs_cost(x) == 1.0 for x in obs])
    assert np.array_equal(viol[1], [box.eval_obs_cost(x) == 1.0 for x in obs])
    lone = cost_spec_of(thr, bounds, 4, 1)
    assert not lone.quad and lone.n_box == 1
    # the mirror classes evaluate Cost.__call__ like the reference (used where autompc is absent)
    class _Step:
        def __init__(self, o, c):
            self.obs, self.ctrl = o, c
    traj = [_Step(o, c) for o, c in zip(z["obs"], z["ctrls"])]
    if not hasattr(thr, "system") or type(thr).__module__.startswith("autompc_b200"):
        assert thr(traj) == float(z["call_thr"]) and box(traj) == float(z["call_box"])

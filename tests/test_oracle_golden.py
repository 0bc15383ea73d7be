"""CPU: pin the NumPy oracle to fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/make_golden.py) and to the reference's
own known answers for QuadCost (reference tests/test_costs.py:192-205)."""
import os

import numpy as np
import pytest

from oracle import ref_loader
from oracle.ilqr_oracle import ilqr_solve
from oracle.mppi_oracle import (MLPParams, MPPIOracle, QuadCostParams, mlp_pred, mlp_pred_batch,
                                mlp_pred_diff_batch)
from tests.helpers import GOLDEN, cartpole_step, load_cartpole


def test_quadcost_known_answers():
    # reference tests/test_costs.py:160-205: three QuadCosts summed at obs=[-1,1]
    # Q1=I, Q2=diag(1,2)... the reference KAT: cost 8, jac [-4,12], hess diag(4,12)
    obs = np.array([-1.0, 1.0])
    qs = [(np.eye(2), np.zeros(2)), (np.diag([1.0, 2.0]), np.zeros(2)), (np.diag([0.0, 3.0]), np.array([1.0, 0.0]))]
    total, jac, hess = 0.0, np.zeros(2), np.zeros((2, 2))
    for Q, g in qs:
        c = QuadCostParams(Q, np.eye(1), goal=g)
        total += c.eval_obs_cost(obs)
        jac += (Q + Q.T) @ (obs - g)
        hess += Q + Q.T
    assert total == 8.0
    assert np.array_equal(jac, [-4.0, 12.0])
    assert np.array_equal(hess, np.diag([4.0, 12.0]))
    # batch form == scalar form
    X = np.random.default_rng(0).normal(size=(7, 2))
    c = QuadCostParams(np.array([[2.0, 0.5], [0.1, 3.0]]), np.eye(1), goal=np.array([0.3, -0.2]))
    assert np.allclose(c.obs_cost_batch(X), [c.eval_obs_cost(x) for x in X], rtol=1e-14)


def test_mlp_cases_match_reference():
    z = np.load(os.path.join(GOLDEN, "mlp_cases.npz"))
    for c in range(int(z["n_cases"])):
        pre = "c%d_" % c
        p = MLPParams.from_npz(z, pre)
        X, U = z[pre + "X"], z[pre + "U"]
        np.testing.assert_allclose(mlp_pred_batch(p, X, U), z[pre + "pred_batch"], rtol=0, atol=2e-14)
        np.testing.assert_allclose(mlp_pred(p, X[0], U[0]), z[pre + "pred0"], rtol=0, atol=2e-14)
        xn, jx, ju = mlp_pred_diff_batch(p, X, U)
        np.testing.assert_allclose(xn, z[pre + "diff_xn"], rtol=0, atol=2e-14)
        np.testing.assert_allclose(jx, z[pre + "diff_jx"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(ju, z[pre + "diff_ju"], rtol=0, atol=1e-13)
        xn1, jx1, ju1 = mlp_pred_diff_batch(p, X[1:2], U[1:2])
        np.testing.assert_allclose(jx1[0], z[pre + "diff1_jx"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(ju1[0], z[pre + "diff1_ju"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name,faithful", [("mppi_cartpole_K256_H20", False), ("mppi_cartpole_K256_H20", True),
                                           ("mppi_cartpole_K100_H5", False), ("mppi_cartpole_K512_H30", False),
                                           ("mppi_cartpole_K4096_H30", False)])
def test_mppi_oracle_matches_reference(name, faithful):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mlp, cost, umin, umax, _, _ = load_cartpole()
    np.random.seed(int(z["seed"]))
    o = MPPIOracle(mlp, cost, umin, umax, horizon=int(z["H"]), num_path=int(z["K"]),
                   sigma=float(z["sigma"]), lmda=float(z["lmda"]), faithful_loop=faithful)
    np.testing.assert_array_equal(o.act_sequence, z["act0"])       # same RNG draws, same order
    constate = np.zeros(5)
    for s in range(int(z["n_steps"])):
        x = z["x0_%d" % s]
        u, constate = o.run(constate, x)
        costs = o.last_costs
        np.testing.assert_allclose(costs, z["costs_%d" % s], rtol=1e-11, atol=1e-9)
        assert int(np.argmin(costs)) == int(z["argmin_%d" % s])    # bit-exact index
        np.testing.assert_allclose(o.last_eps.sum(axis=1), z["eps_clip_sum_%d" % s], rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(o.act_sequence, z["act_%d" % s], rtol=0, atol=1e-9)
        np.testing.assert_allclose(u, z["u_%d" % s], rtol=0, atol=1e-8)
    if "run_api_us" in z.files:
        np.testing.assert_allclose([z["u_%d" % s] for s in range(int(z["n_steps"]))], z["run_api_us"], atol=0)


def test_reference_raises_for_multi_dim_ctrl_fixture():
    z = np.load(os.path.join(GOLDEN, "mppi_nu6_reference_raises.npz"))
    assert "could not broadcast" in str(z["message"])


@pytest.mark.parametrize("name", ["ilqr_cartpole_H50", "ilqr_cartpole_H10"])
def test_ilqr_oracle_matches_reference(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mlp, cost, umin, umax, _, dt = load_cartpole()
    H = int(z["H"])
    for i in range(int(z["n"])):
        pre = "p%d_" % i
        r = ilqr_solve(mlp, cost, dt, z[pre + "x0"], H, (umin, umax))
        assert r["converged"] == bool(z[pre + "converged"])
        assert r["n_iter"] == int(z[pre + "n_iter"])
        assert r["ls_fail"] == bool(z[pre + "ls_fail"])
        assert r["alpha_idx"] == list(z[pre + "alpha_idx"])         # bit-exact integer trace
        np.testing.assert_allclose(r["states"], z[pre + "states"], rtol=0, atol=1e-8)
        np.testing.assert_allclose(r["ctrls"], z[pre + "ctrls"], rtol=0, atol=1e-8)
        np.testing.assert_allclose(r["Ks"], z[pre + "Ks"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(r["ks"], z[pre + "ks"], rtol=1e-8, atol=1e-8)
        u = r["ctrls"][0] + r["Ks"][0] @ (z[pre + "x0"] - r["states"][0])
        np.testing.assert_allclose(u, z[pre + "run_u"], rtol=0, atol=1e-8)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_oracle_live_against_reference_three_solves():
    """Belt and braces: run the unmodified reference side by side (not just fixtures)."""
    ns = ref_loader.load()
    from oracle.make_golden import make_cartpole, random_mlp
    system, task = make_cartpole(ns)
    mlp_ref = random_mlp(ns, system, 2, 32, "tanh", seed=7)
    p = MLPParams.from_reference_mlp(mlp_ref)
    cost = QuadCostParams(*task.get_cost().get_cost_matrices(), task.get_cost().get_goal())
    np.random.seed(11)
    with ref_loader.quiet():
        ref = ns.MPPI(system, task, mlp_ref, horizon=12, num_path=96, sigma=0.8, lmda=0.7)
    np.random.seed(11)
    o = MPPIOracle(p, cost, [-20.0], [20.0], horizon=12, num_path=96, sigma=0.8, lmda=0.7)
    x = np.array([0.5, 0.1, -0.2, 0.3])
    cs_ref = cs_o = np.zeros(5)
    for _ in range(3):
        state = np.random.get_state()
        u_ref, cs_ref = ref.run(cs_ref, x)
        np.random.set_state(state)
        u_o, cs_o = o.run(cs_o, x)
        np.testing.assert_allclose(u_o, u_ref, rtol=0, atol=1e-10)
        np.testing.assert_allclose(o.act_sequence, ref.act_sequence, rtol=0, atol=1e-10)
        x = cartpole_step(x, u_ref)


def _sumcost_terms(z):
    return [QuadCostParams(z["Q1"], z["R1"], z["F1"], z["g1"]), QuadCostParams(z["Q2"], z["R2"], z["F2"], z["g2"])]


def test_sumcost_oracle_matches_reference():
    """the reference MPPI under a SumCost of two QuadCosts with different goals."""
    from oracle.mppi_oracle import SumQuadCostParams
    z = np.load(os.path.join(GOLDEN, "mppi_cartpole_sumcost_K256_H20.npz"))
    mlp, _, umin, umax, _, _ = load_cartpole()
    np.random.seed(int(z["seed"]))
    o = MPPIOracle(mlp, SumQuadCostParams(_sumcost_terms(z)), umin, umax, horizon=int(z["H"]), num_path=int(z["K"]),
                   sigma=float(z["sigma"]), lmda=float(z["lmda"]))
    np.testing.assert_array_equal(o.act_sequence, z["act0"])
    constate = np.zeros(5)
    for s in range(int(z["n_steps"])):
        u, constate = o.run(constate, z["x0_%d" % s])
        np.testing.assert_allclose(o.last_costs, z["costs_%d" % s], rtol=1e-11, atol=1e-9)
        assert int(np.argmin(o.last_costs)) == int(z["argmin_%d" % s])
        np.testing.assert_allclose(o.act_sequence, z["act_%d" % s], rtol=0, atol=1e-9)
        np.testing.assert_allclose(u, z["u_%d" % s], rtol=0, atol=1e-8)


def test_model_rmse_oracle_matches_reference():
    """get_model_rmse (model_metrics.py:12-43) at several horizons, equal and ragged lengths."""
    from oracle.mppi_oracle import model_rmse
    z = np.load(os.path.join(GOLDEN, "model_rmse_cartpole.npz"))
    mlp = load_cartpole()[0]
    obs, ctrls = list(z["obs"]), list(z["ctrls"])
    for h in z["horizons"]:
        assert abs(model_rmse(mlp, obs, ctrls, int(h)) - float(z["rmse_h%d" % h])) < 1e-12
    lens = z["ragged_lens"]
    r = model_rmse(mlp, [obs[i][:n] for i, n in enumerate(lens)], [ctrls[i][:n] for i, n in enumerate(lens)], 5)
    assert abs(r - float(z["rmse_ragged_h5"])) < 1e-12

"""GPU parity: the single-launch float64 iLQR solve (through the C ABI) against fixtures recorded
from the UNMODIFIED reference ``IterativeLQR.compute_ilqr_default`` (ilqr.py:100-265) / ``run``
(:267-295), and against the oracle on other problems.
Exact: adopted line-search index per iteration, iteration count, converged, ls_fail.
float64 both sides: trajectories atol 1e-7 (50 outer iterations of accumulated re-ordering)."""
import os

import numpy as np
import pytest

from oracle.ilqr_oracle import ilqr_solve
from oracle.mppi_oracle import QuadCostParams
from tests.helpers import GOLDEN, load_cartpole, synthetic_mlp

pytestmark = pytest.mark.gpu


def _ctl(p, cost, umin, umax, dt, H, **kw):
    from autompc_b200 import IterativeLQR
    from tests.gpu_helpers import problem_of
    system, task, model = problem_of(p, cost, umin, umax, dt=dt)
    return IterativeLQR(system, task, model, horizon=H, **kw)


@pytest.mark.parametrize("name", ["ilqr_cartpole_H50", "ilqr_cartpole_H10"])
def test_ilqr_matches_unmodified_reference_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mlp, cost, umin, umax, _, dt = load_cartpole()
    H = int(z["H"])
    ctl = _ctl(mlp, cost, umin, umax, dt, H)
    for i in range(int(z["n"])):
        pre = "p%d_" % i
        conv, states, ctrls, Ks, ks = ctl.compute_ilqr(z[pre + "x0"])
        info = ctl.last_info
        assert conv == bool(z[pre + "converged"])
        assert info["n_iter"] == int(z[pre + "n_iter"])
        assert info["ls_fail"] == bool(z[pre + "ls_fail"])
        assert info["alpha_idx"] == [int(a) for a in z[pre + "alpha_idx"]]      # bit-exact integer trace
        np.testing.assert_allclose(states, z[pre + "states"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(ctrls, z[pre + "ctrls"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(Ks, z[pre + "Ks"], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(ks, z[pre + "ks"], rtol=1e-7, atol=1e-7)
        ctl.reset()
        u, newstate = ctl.run(np.concatenate([z[pre + "x0"], np.zeros(1)]), z[pre + "x0"])
        np.testing.assert_allclose(u, z[pre + "run_u"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(newstate, np.concatenate([z[pre + "x0"], u]))
    ctl.close()


@pytest.mark.parametrize("nx,nu,hidden,act,H", [(6, 2, [32, 32], "tanh", 15), (17, 6, [64, 64], "relu", 8)])
def test_ilqr_matches_oracle_multi_dim_ctrl(nx, nu, hidden, act, H):
    rng = np.random.default_rng(2)
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=6)
    A = rng.normal(size=(nx, nx))
    cost = QuadCostParams(np.eye(nx) + 0.1 * A @ A.T, 0.05 * np.eye(nu), 5 * np.eye(nx), goal=0.1 * rng.normal(size=nx))
    umin, umax = -np.ones(nu), 1.5 * np.ones(nu)
    x0 = rng.normal(size=nx)
    r = ilqr_solve(p, cost, 0.05, x0, H, (umin, umax))
    ctl = _ctl(p, cost, umin, umax, 0.05, H)
    conv, states, ctrls, Ks, ks = ctl.compute_ilqr(x0)
    assert conv == r["converged"] and ctl.last_info["n_iter"] == r["n_iter"]
    assert ctl.last_info["alpha_idx"] == r["alpha_idx"]
    np.testing.assert_allclose(states, r["states"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(ctrls, r["ctrls"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(Ks, r["Ks"], rtol=1e-6, atol=1e-7)
    ctl.close()


def _solve(p, cost, umin, umax, dt, H, x0, **kw):
    ctl = _ctl(p, cost, umin, umax, dt, H, **kw)
    conv, states, ctrls, Ks, ks = ctl.compute_ilqr(x0)
    info = dict(ctl.last_info)
    ctl.close()
    return conv, states, ctrls, Ks, ks, info


def test_ilqr_line_search_paths_agree(monkeypatch):
    """The line search on FP64 tensor-core fragments (step sizes in tiles of eight, evaluated until the acceptance rule
    stops) against the CUDA-core line search that rolls all step sizes out: same decisions, same trajectories."""
    mlp, cost, umin, umax, _, dt = load_cartpole()
    z = np.load(os.path.join(GOLDEN, "ilqr_cartpole_H50.npz"))
    x0 = z["p0_x0"]
    monkeypatch.delenv("AMPC_ILQR_NO_MMA", raising=False)
    a = _solve(mlp, cost, umin, umax, dt, 50, x0)
    monkeypatch.setenv("AMPC_ILQR_NO_MMA", "1")
    b = _solve(mlp, cost, umin, umax, dt, 50, x0)
    assert a[0] == b[0] and a[5]["alpha_idx"] == b[5]["alpha_idx"] and a[5]["n_iter"] == b[5]["n_iter"]
    np.testing.assert_allclose(a[1], b[1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(a[2], b[2], rtol=0, atol=1e-9)


@pytest.mark.parametrize("ls_max_iter,thr", [(12, 0.97), (20, 0.999), (5, 0.3)])
def test_ilqr_later_step_size_tiles_match_oracle(ls_max_iter, thr):
    """A demanding acceptance threshold sends the rule past the first eight step sizes (second and third tile of the
    tensor-core line search, and the best-so-far fallback when none is accepted); ls_max_iter < 8 is a partial tile."""
    nx, nu, H = 6, 2, 12
    rng = np.random.default_rng(4)
    p = synthetic_mlp(nx, nu, [48, 40], act="relu", seed=9)
    A = rng.normal(size=(nx, nx))
    cost = QuadCostParams(np.eye(nx) + 0.1 * A @ A.T, 0.05 * np.eye(nu), 5 * np.eye(nx), goal=0.1 * rng.normal(size=nx))
    umin, umax = -np.ones(nu), 1.5 * np.ones(nu)
    x0 = rng.normal(size=nx)
    r = ilqr_solve(p, cost, 0.05, x0, H, (umin, umax), ls_max_iter=ls_max_iter, ls_discount=0.6, ls_cost_threshold=thr)
    conv, states, ctrls, Ks, ks, info = _solve(p, cost, umin, umax, 0.05, H, x0, ls_max_iter=ls_max_iter, ls_discount=0.6,
                                               ls_cost_threshold=thr)
    assert conv == r["converged"] and info["n_iter"] == r["n_iter"] and info["alpha_idx"] == r["alpha_idx"]
    if ls_max_iter > 8:
        assert max(info["alpha_idx"]) >= 8, "the case is meant to reach the second tile of step sizes"
    np.testing.assert_allclose(states, r["states"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(ctrls, r["ctrls"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("env", [{"AMPC_ILQR_NO_JAC_MMA": "1"}, {"AMPC_ILQR_NO_SMEM": "1"}])
def test_ilqr_fallback_paths_match_reference_fixture(env, monkeypatch):
    """The other routes through the solve kernel -- tensor-core line search with the one-warp-per-step Jacobian refresh
    (the phase scratch is shared: the weight image is rebuilt per call), and the kernel on global scratch (nothing
    resident, CUDA-core line search) -- against the unmodified-reference fixture."""
    for k in ("AMPC_ILQR_NO_MMA", "AMPC_ILQR_NO_JAC_MMA", "AMPC_ILQR_NO_SMEM"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    z = np.load(os.path.join(GOLDEN, "ilqr_cartpole_H10.npz"))
    mlp, cost, umin, umax, _, dt = load_cartpole()
    ctl = _ctl(mlp, cost, umin, umax, dt, int(z["H"]))
    conv, states, ctrls, Ks, ks = ctl.compute_ilqr(z["p0_x0"])
    info = ctl.last_info
    assert conv == bool(z["p0_converged"]) and info["n_iter"] == int(z["p0_n_iter"])
    assert info["alpha_idx"] == [int(a) for a in z["p0_alpha_idx"]]
    np.testing.assert_allclose(states, z["p0_states"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(ctrls, z["p0_ctrls"], rtol=0, atol=1e-7)
    ctl.close()

"""GPU parity: float64 MLP inference / Jacobian kernels (through the C ABI) against
fixtures recorded from the UNMODIFIED reference ``autompc.sysid.mlp.MLP``
(``pred`` / ``pred_batch`` mlp.py:219-236, ``pred_diff`` / ``pred_diff_batch`` :238-305).
Tolerance: float64 on both sides, different summation order -> atol 1e-12 (states ~1), Jacobians 1e-11."""
import os

import numpy as np
import pytest

from oracle.mppi_oracle import MLPParams, mlp_pred_batch, mlp_pred_diff_batch
from tests.helpers import GOLDEN, synthetic_mlp

pytestmark = pytest.mark.gpu


def _model(p):
    from autompc_b200 import B200MLP
    from autompc_b200.plugin import System
    from tests.gpu_helpers import weights_of
    system = System(["x%d" % i for i in range(p.nx)], ["u%d" % i for i in range(p.nu)])
    return B200MLP(system, weights_of(p))


def test_mlp_kernels_match_reference_fixture():
    z = np.load(os.path.join(GOLDEN, "mlp_cases.npz"))
    for c in range(int(z["n_cases"])):
        pre = "c%d_" % c
        p = MLPParams.from_npz(z, pre)
        m = _model(p)
        X, U = z[pre + "X"], z[pre + "U"]
        np.testing.assert_allclose(m.pred_batch(X, U), z[pre + "pred_batch"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(m.pred(X[0], U[0]), z[pre + "pred0"], rtol=0, atol=1e-12)
        xn, jx, ju = m.pred_diff_batch(X, U)
        np.testing.assert_allclose(xn, z[pre + "diff_xn"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(jx, z[pre + "diff_jx"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(ju, z[pre + "diff_ju"], rtol=0, atol=1e-11)
        xn1, jx1, ju1 = m.pred_diff(X[1], U[1])
        np.testing.assert_allclose(xn1, z[pre + "diff1_xn"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(jx1, z[pre + "diff1_jx"], rtol=0, atol=1e-11)
        np.testing.assert_allclose(ju1, z[pre + "diff1_ju"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("act", ["relu", "tanh", "sigmoid", "selu"])
def test_mlp_kernels_match_oracle_large(act):
    """C3-sized network, ragged batch sizes (incl. 1 and a non-multiple of the CTA width)."""
    p = synthetic_mlp(17, 6, [256, 256, 256], act=act, seed=4)
    m = _model(p)
    rng = np.random.default_rng(0)
    for batch in (1, 50, 333):
        X, U = rng.normal(size=(batch, 17)), rng.normal(size=(batch, 6))
        np.testing.assert_allclose(m.pred_batch(X, U), mlp_pred_batch(p, X, U), rtol=0, atol=1e-12)
        xn, jx, ju = m.pred_diff_batch(X, U)
        rxn, rjx, rju = mlp_pred_diff_batch(p, X, U)
        np.testing.assert_allclose(xn, rxn, rtol=0, atol=1e-12)
        np.testing.assert_allclose(jx, rjx, rtol=0, atol=1e-11)
        np.testing.assert_allclose(ju, rju, rtol=0, atol=1e-11)


@pytest.mark.parametrize("dims", [(17, 6, [256, 256, 256], "relu"), (4, 1, [64, 64], "tanh"), (5, 2, [33], "sigmoid")])
def test_sample_blocked_kernels_equal_one_sample_kernels(dims, monkeypatch):
    """pred_batch / k-step rollouts: the eight-samples-per-CTA kernels (batches beyond one wave) and the
    one-sample-per-CTA kernels do the same arithmetic per sample -> identical bits, and both meet the oracle.
    Batch sizes: one sample, a ragged last CTA, the dispatch threshold's two sides."""
    nx, nu, hidden, act = dims
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=11)
    m = _model(p)
    rng = np.random.default_rng(3)
    for batch in (1, 7, 8, 9, 1184, 1185, 2051):
        X, U = rng.normal(size=(batch, nx)), 0.3 * rng.normal(size=(3, batch, nu))
        got = {}
        for forced in ("0", "1", None):
            if forced is None:
                monkeypatch.delenv("AMPC_MLP_BLOCKED", raising=False)
            else:
                monkeypatch.setenv("AMPC_MLP_BLOCKED", forced)
            got[forced] = (m.pred_batch(X, U[0]), m.rollout_batch(X, U))
        for forced in ("1", None):
            np.testing.assert_array_equal(got[forced][0], got["0"][0])
            np.testing.assert_array_equal(got[forced][1], got["0"][1])
        np.testing.assert_allclose(got["1"][0], mlp_pred_batch(p, X, U[0]), rtol=0, atol=1e-12)
        chained = X
        for k in range(3):
            chained = mlp_pred_batch(p, chained, U[k])
        np.testing.assert_allclose(got["1"][1], chained, rtol=0, atol=1e-11)


@pytest.mark.parametrize("dims", [(5, 2, [33, 17], "tanh"), (3, 1, [7], "relu"), (9, 4, [64, 31, 50, 12], "selu"),
                                  (32, 20, [256, 128], "sigmoid")])
def test_jacobian_kernels_odd_widths_and_ragged_column_groups(dims):
    """Forward-mode Jacobians with odd layer widths (the even / odd-k partial sums' tail) and input widths that are
    not a multiple of the four-column groups a thread owns (nin = 7, 4, 13, 52)."""
    nx, nu, hidden, act = dims
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=21)
    m = _model(p)
    rng = np.random.default_rng(5)
    for batch in (1, 6):
        X, U = rng.normal(size=(batch, nx)), rng.normal(size=(batch, nu))
        xn, jx, ju = m.pred_diff_batch(X, U)
        rxn, rjx, rju = mlp_pred_diff_batch(p, X, U)
        np.testing.assert_allclose(xn, rxn, rtol=0, atol=1e-12)
        np.testing.assert_allclose(jx, rjx, rtol=0, atol=1e-11)
        np.testing.assert_allclose(ju, rju, rtol=0, atol=1e-11)


def test_mlp_parameter_round_trip_and_errors():
    p = synthetic_mlp(4, 1, [64, 64], seed=8)
    m = _model(p)
    params = m.get_parameters()                       # same keys as mlp.py:308-313
    assert set(params) >= {"net_state", "xu_means", "xu_std", "dy_means", "dy_std"}
    m2 = _model(synthetic_mlp(4, 1, [64, 64], seed=9))
    m2.set_parameters(params)
    X, U = np.ones((3, 4)), np.ones((3, 1))
    np.testing.assert_array_equal(m.pred_batch(X, U), m2.pred_batch(X, U))
    with pytest.raises(ValueError):
        m.pred_batch(np.ones((3, 5)), U)
    with pytest.raises(NotImplementedError):
        m.train([])


def test_kstep_rollout_and_model_rmse_match_reference_fixture():
    """get_model_rmse (autompc/evaluation/model_metrics.py:12-43) on the device.  Checker: values the
    unmodified reference computed on the same trajectories; and the fused k-step rollout equals chained pred_batch."""
    from autompc_b200 import get_model_rmse
    from oracle.mppi_oracle import model_rmse
    from tests.helpers import load_cartpole
    z = np.load(os.path.join(GOLDEN, "model_rmse_cartpole.npz"))
    p = load_cartpole()[0]
    m = _model(p)
    obs, ctrls = list(z["obs"]), list(z["ctrls"])
    for h in z["horizons"]:
        r = get_model_rmse(m, list(zip(obs, ctrls)), horizon=int(h))
        assert abs(r - float(z["rmse_h%d" % h])) < 1e-11
        assert abs(r - model_rmse(p, obs, ctrls, int(h))) < 1e-11
    lens = z["ragged_lens"]
    ragged = [(obs[i][:n], ctrls[i][:n]) for i, n in enumerate(lens)]
    assert abs(get_model_rmse(m, ragged, horizon=5) - float(z["rmse_ragged_h5"])) < 1e-11
    # one launch == `horizon` chained launches, bit for bit (same per-step arithmetic)
    X, U = obs[0][:40], np.stack([ctrls[0][k:k + 40] for k in range(7)])
    chained = X
    for k in range(7):
        chained = m.pred_batch(chained, U[k])
    np.testing.assert_array_equal(m.rollout_batch(X, U), chained)
    with pytest.raises(ValueError):
        get_model_rmse(m, [(obs[0][:3], ctrls[0][:3])], horizon=5)      # no trajectory longer than the horizon
    with pytest.raises(ValueError):
        m.rollout_batch(X, U[:, :5])

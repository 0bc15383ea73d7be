"""GPU parity for SURVEY.md 8(f) row 4: linear-model ``pred_batch`` (ARX, Koopman) and the direct-transcription
callbacks ``get_constraint`` / ``get_jacobian``, against fixtures recorded from the UNMODIFIED reference
(oracle/make_golden_r2.py) and against the oracle restatement at other sizes."""
import os

import numpy as np
import pytest

from oracle.mppi_oracle import linear_pred_batch, nmpc_constraint, nmpc_jacobian
from tests.helpers import GOLDEN, load_cartpole, synthetic_mlp

pytestmark = pytest.mark.gpu


def _system(nx, nu):
    from autompc_b200.plugin import System
    s = System(["x%d" % i for i in range(nx)], ["u%d" % i for i in range(nu)])
    s.dt = 0.05
    return s


class _Base:
    """Stands in for the trained reference model a B200Linear delegates the host bookkeeping to."""

    def __init__(self, system, A, B):
        self.system, self.A, self.B = system, A, B

    def to_linear(self):
        return self.A, self.B

    def traj_to_state(self, traj):
        return np.zeros(self.A.shape[0])

    def update_state(self, state, new_ctrl, new_obs):
        s = self.A @ state + self.B @ new_ctrl
        s[:self.system.obs_dim] = new_obs
        return s


@pytest.mark.parametrize("kind", ["arx", "koop"])
def test_linear_pred_batch_matches_unmodified_reference(kind):
    """ARX.pred_batch (arx.py:151-154) with the matrices the reference's own ARX.train produced; Koopman.pred_batch
    (koopman.py:170-173) on a lifted state."""
    from autompc_b200 import B200Linear
    z = np.load(os.path.join(GOLDEN, "linear_models.npz"))
    A, B, X, U = z[kind + "_A"], z[kind + "_B"], z[kind + "_X"], z[kind + "_U"]
    system = _system(4, 1)
    m = B200Linear.from_model(_Base(system, A, B))
    assert m.state_dim == A.shape[0]
    np.testing.assert_allclose(m.pred_batch(X, U), z[kind + "_pred_batch"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(m.pred(X[0], U[0]), z[kind + "_pred0"], rtol=0, atol=1e-12)
    xn, Ja, Jb = m.pred_diff(X[1], U[1])
    assert np.array_equal(Ja, A) and np.array_equal(Jb, B)
    xb, Jab, Jbb = m.pred_diff_batch(X, U)
    assert Jab.shape == (X.shape[0],) + A.shape and np.array_equal(Jab[3], A) and np.array_equal(Jbb[2], B)
    if kind == "arx":                                   # host bookkeeping goes through the base model (arx.py:98-103)
        st = z["arx_traj_to_state"]
        np.testing.assert_allclose(m.update_state(st, z["arx_traj_ctrls"][6], z["arx_traj_obs"][6] * 0 + 1.0)[:4], 1.0)
    m.close()


@pytest.mark.parametrize("ns,nu,batch", [(1, 1, 1), (17, 6, 1000), (130, 3, 77), (64, 8, 4097)])
def test_linear_pred_batch_matches_oracle_sizes(ns, nu, batch):
    """Ragged batches (not a multiple of the 8-sample tile), ns above the CTA width, the empty batch."""
    from autompc_b200 import B200Linear
    rng = np.random.default_rng(ns)
    A, B = rng.normal(size=(ns, ns)) / np.sqrt(ns), rng.normal(size=(ns, nu))
    X, U = rng.normal(size=(batch, ns)), rng.normal(size=(batch, nu))
    m = B200Linear(_system(ns, nu), A, B)
    np.testing.assert_allclose(m.pred_batch(X, U), linear_pred_batch(A, B, X, U), rtol=1e-13, atol=1e-12)
    assert m.pred_batch(X[:0], U[:0]).shape == (0, ns)
    with pytest.raises(ValueError):
        m.pred_batch(X[:, :-1] if ns > 1 else np.zeros((2, 3)), U[:2] if ns == 1 else U)
    m.close()


def test_nmpc_callbacks_match_unmodified_reference():
    """NonLinearMPCProblem.get_constraint / get_jacobian (nmpc.py:102-110, :148-187) with the cartpole MLP."""
    from autompc_b200 import B200MLP, NonLinearMPCProblem
    from autompc_b200.plugin import QuadCost, Task
    from tests.gpu_helpers import weights_of
    z = np.load(os.path.join(GOLDEN, "nmpc_cartpole_H8.npz"))
    mlp, cost, umin, umax, _, _ = load_cartpole()
    system = _system(4, 1)
    task = Task(system)
    task.set_ctrl_bounds(np.asarray(umin, float), np.asarray(umax, float))
    task.set_cost(QuadCost(system, cost.Q, cost.R, cost.F, goal=cost.goal))
    prob = NonLinearMPCProblem(system, B200MLP(system, weights_of(mlp)), task, int(z["H"]))
    assert prob.dimx == int(z["dimx"]) and prob.dimc == int(z["dimc"]) and prob.nnz == z["jac"].size
    np.testing.assert_allclose(prob.get_constraint(z["x"]), z["constraint"], rtol=0, atol=1e-12)
    row, col = prob.get_jacobian(z["x"], True)
    assert np.array_equal(row, z["row"]) and np.array_equal(col, z["col"])        # sparsity pattern: exact
    np.testing.assert_allclose(prob.get_jacobian(z["x"], False), z["jac"], rtol=0, atol=1e-11)
    lb, ub = prob.get_variable_bounds()
    assert lb.shape == (prob.dimx,) and np.all(lb[-8:] == -20.0) and np.all(ub[-8:] == 20.0)
    assert np.isfinite(prob.get_cost(z["x"])) and prob.get_gradient(z["x"]).shape == (prob.dimx,)


@pytest.mark.parametrize("nx,nu,hidden,act,H", [(17, 6, [256, 256, 256], "relu", 50), (3, 2, [24, 16], "tanh", 1),
                                                (6, 3, [100], "selu", 13)])
def test_nmpc_callbacks_match_oracle_sizes(nx, nu, hidden, act, H):
    from autompc_b200 import B200MLP, NonLinearMPCProblem
    from autompc_b200.plugin import QuadCost, Task
    from tests.gpu_helpers import weights_of
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=4)
    system = _system(nx, nu)
    task = Task(system)
    task.set_ctrl_bounds(-np.ones(nu), np.ones(nu))
    task.set_cost(QuadCost(system, np.eye(nx), np.eye(nu), np.eye(nx)))
    prob = NonLinearMPCProblem(system, B200MLP(system, weights_of(p)), task, H)
    x = np.random.default_rng(H).normal(size=prob.dimx)
    np.testing.assert_allclose(prob.get_constraint(x), nmpc_constraint(p, H, x), rtol=0, atol=1e-12)
    rows, cols, vals = nmpc_jacobian(p, H, x)
    r, c = prob.get_jacobian(x, True)
    assert np.array_equal(r, rows) and np.array_equal(c, cols)
    np.testing.assert_allclose(prob.get_jacobian(x, False), vals, rtol=0, atol=1e-11)


def test_nmpc_linear_model_jacobian_is_constant():
    from autompc_b200 import B200Linear, NonLinearMPCProblem
    from autompc_b200.plugin import QuadCost, Task
    rng = np.random.default_rng(2)
    A, B = rng.normal(size=(5, 5)), rng.normal(size=(5, 2))
    system = _system(5, 2)
    task = Task(system)
    task.set_ctrl_bounds(-np.ones(2), np.ones(2))
    task.set_cost(QuadCost(system, np.eye(5), np.eye(2), np.eye(5)))
    prob = NonLinearMPCProblem(system, B200Linear(system, A, B), task, 4)
    x = rng.normal(size=prob.dimx)
    st, ct = x[:25].reshape(5, 5), x[25:].reshape(4, 2)
    np.testing.assert_allclose(prob.get_constraint(x), (-st[1:] + linear_pred_batch(A, B, st[:4], ct)).ravel(), atol=1e-12)
    j = prob.get_jacobian(x, False).reshape(4, -1)
    assert np.array_equal(j[2, :25], A.ravel()) and np.array_equal(j[1, 25:35], B.ravel()) and np.all(j[:, 35:] == -1)

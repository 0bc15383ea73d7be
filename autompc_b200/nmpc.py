"""Direct-transcription problem callbacks on the B200 engine -- SURVEY.md 8(f) row 4.

``NonLinearMPCProblem`` mirrors ``autompc.control.nmpc.NonLinearMPCProblem`` (``autompc/control/nmpc.py:36-187``): the
object an NLP solver (IPOPT through ``IpoptWrapper``, nmpc.py:190-211) calls back into.  The two callbacks that touch
the dynamics model run on the device:

* ``get_constraint(x)``   (nmpc.py:102-110)  -- ``pred_batch`` at batch H, fused with ``-state[i+1] + ...``;
* ``get_jacobian(x, False)`` (nmpc.py:170-187) -- ``pred_diff_batch`` at batch H, written straight into the
  reference's sparse value order; ``get_jacobian(x, True)`` (the (row, col) pattern, nmpc.py:148-169) is index
  arithmetic on the host.

The solver itself is out of scope (cyipopt is not installable here), and so is
``DirectTranscriptionController``; ``get_cost`` / ``get_gradient`` (nmpc.py:72-100) are host NumPy on the task's
quadratic cost, as in the reference.  Models: ``B200MLP`` (or anything ``MLPWeights.from_model`` reads) and
``B200Linear``.
"""
import numpy as np

from . import _abi
from .linear import B200Linear
from .mlp import B200MLP, MLPWeights


class NonLinearMPCProblem:
    def __init__(self, system, model, task, horizon):
        self.system, self.task, self.horizon = system, task, int(horizon)
        if isinstance(model, (B200MLP, B200Linear)):
            self.model = model
        elif hasattr(model, "to_linear") and not hasattr(model, "net"):
            self.model = B200Linear.from_model(model)
        else:
            self.model = B200MLP(system, MLPWeights.from_model(model))
        dc, ds = system.ctrl_dim, self.model.state_dim
        self.ctrl_dim, self.obs_dim = dc, ds
        self.dimx = ds * (self.horizon + 1) + dc * self.horizon          # x0..xN, u0..u_{N-1}   (nmpc.py:48-50)
        self.dimc = self.horizon * ds
        self._row, self._col = self.get_jacobian(None, True)
        self._cost = None

    @property
    def nnz(self):
        return self._row.size

    def _split(self, x):
        x = _abi.f64(x, (self.dimx,))
        len1 = (self.horizon + 1) * self.obs_dim
        return x, x[:len1].reshape(self.horizon + 1, self.obs_dim), x[len1:].reshape(self.horizon, self.ctrl_dim)

    # --- dynamics callbacks (device) -------------------------------------------------------
    def get_constraint(self, x):
        """c[i] = -state[i+1] + pred(state[i], ctrl[i])  (nmpc.py:102-110)."""
        x, st, ct = self._split(x)
        H = self.horizon
        if isinstance(self.model, B200MLP):
            self.model._need()
            c = np.empty(self.dimc)
            _abi.check(_abi.lib().ampc_mlp_nmpc_constraint(self.model._h, H, _abi.dptr(x), _abi.dptr(c)))
            return c
        return (-st[1:] + self.model.pred_batch(st[:H], ct)).reshape(-1)

    def get_jacobian(self, x, return_rowcol):
        dims, dimu, H = self.obs_dim, self.ctrl_dim, self.horizon
        if return_rowcol:                                                 # nmpc.py:148-169
            r_s, c_s = np.divmod(np.arange(dims * dims), dims)
            r_u, c_u = np.divmod(np.arange(dims * dimu), dimu)
            base_u = dims * (H + 1)
            row, col = [], []
            for i in range(H):
                cr = i * dims
                row += [cr + r_s, cr + r_u, cr + np.arange(dims)]
                col += [i * dims + c_s, base_u + i * dimu + c_u, (i + 1) * dims + np.arange(dims)]
            return np.concatenate(row).astype(np.float64), np.concatenate(col).astype(np.float64)
        x, st, ct = self._split(x)
        per = dims * dims + dims * dimu + dims
        jac = np.empty(H * per)
        if isinstance(self.model, B200MLP):                               # nmpc.py:170-187
            self.model._need()
            _abi.check(_abi.lib().ampc_mlp_nmpc_jacobian(self.model._h, H, _abi.dptr(x), _abi.dptr(jac)))
            return jac
        A, B = self.model.to_linear()
        jac.reshape(H, per)[:] = np.concatenate([A.ravel(), B.ravel(), -np.ones(dims)])
        return jac

    # --- bounds (host) ---------------------------------------------------------------------
    def get_constr_bounds(self):                                          # nmpc.py:112-115
        return np.zeros(self.dimc), np.zeros(self.dimc)

    def get_variable_bounds(self):                                        # nmpc.py:117-131
        ds, dc, H = self.obs_dim, self.ctrl_dim, self.horizon
        statebd = np.zeros((ds, 2))
        statebd[:, 0], statebd[:, 1] = -np.inf, np.inf
        if hasattr(self.task, "get_obs_bounds"):
            statebd[:self.system.obs_dim, :] = self.task.get_obs_bounds()
        ctrlbd = self.task.get_ctrl_bounds()
        xlb, xub = np.zeros(self.dimx), np.zeros(self.dimx)
        xlb[:(H + 1) * ds].reshape(-1, ds)[:] = statebd[:, 0]
        xub[:(H + 1) * ds].reshape(-1, ds)[:] = statebd[:, 1]
        xlb[(H + 1) * ds:].reshape(-1, dc)[:] = ctrlbd[:, 0]
        xub[(H + 1) * ds:].reshape(-1, dc)[:] = ctrlbd[:, 1]
        return xlb, xub

    # --- objective (host NumPy on the quadratic cost, nmpc.py:72-100) ------------------------
    def _quad(self):
        if self._cost is None:
            c = self.task.get_cost()
            Q, R, F = c.get_cost_matrices()
            self._cost = (np.asarray(Q, float), np.asarray(R, float), np.asarray(F, float),
                          np.asarray(c.get_goal(), float))
        return self._cost

    def get_cost(self, x):
        _, st, ct = self._split(x)
        Q, R, F, g = self._quad()
        n, dt = self.system.obs_dim, self.system.dt
        d = st[:, :n] - g
        tc = d[-1] @ F @ d[-1]
        tc += dt * np.einsum("ki,ij,kj->", d, Q, d)
        tc += dt * np.einsum("ki,ij,kj->", ct, R, ct)
        return float(tc)

    def get_gradient(self, x):
        _, st, ct = self._split(x)
        Q, R, F, g = self._quad()
        n, dt, H = self.system.obs_dim, self.system.dt, self.horizon
        grad = np.zeros(self.dimx)
        gs = grad[:(H + 1) * self.obs_dim].reshape(H + 1, self.obs_dim)
        gu = grad[(H + 1) * self.obs_dim:].reshape(H, self.ctrl_dim)
        gs[-1, :n] = (F + F.T) @ st[-1, :n]                               # terminal: no goal (cost.py:194-199)
        gs[:, :n] += dt * (st[:, :n] - g) @ (Q + Q.T).T
        gu[:] = dt * ct @ (R + R.T).T
        return grad

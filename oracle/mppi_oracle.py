"""float64 NumPy restatement of the reference MPPI solve -- TEST INFRASTRUCTURE.

See ``oracle/__init__.py`` for the rules (never imported by the product path).

What is restated, with the reference lines each function follows
(paths relative to ``/root/reference``):

* ``mlp_pred_batch``   <- ``autompc/sysid/mlp.py:229-236`` (+ ``:20-30`` z-score
  transforms, ``:55-59`` ``ForwardNet.forward``).  float64 like ``mlp.py:165``.
* ``mlp_pred_diff_batch`` <- ``autompc/sysid/mlp.py:281-305`` in closed form
  (chain rule through the layer stack instead of autograd).
* ``QuadCostParams`` methods <- ``autompc/costs/cost.py:66-83``, ``:118-134``,
  ``:166-183`` for a ``QuadCost`` (``autompc/costs/quad_cost.py:7-51``).
* ``SumQuadCostParams`` <- ``autompc/costs/sum_cost.py:9-81``;  ``model_rmse`` <-
  ``autompc/evaluation/model_metrics.py:12-43``.
* ``ThresholdCostParams`` / ``BoxThresholdCostParams`` <- ``autompc/costs/thresh_cost.py:8-83``;
  ``traj_cost`` <- ``Cost.__call__`` (``autompc/costs/cost.py:27-41``).
* ``linear_pred_batch`` <- ``autompc/sysid/arx.py:151-154`` == ``autompc/sysid/koopman.py:170-173``;
  ``nmpc_constraint`` / ``nmpc_jacobian`` <- ``autompc/control/nmpc.py:102-110``, ``:148-187``.
* ``MPPIOracle`` <- ``autompc/control/mppi.py:66-181``: ctor draw ``:97-99``,
  ``do_rollouts`` ``:120-152``, ``update`` ``:110-118``, ``run`` ``:154-168``.

Reference quirks that are reproduced on purpose:
 (a) the action sequence is shifted BEFORE the rollouts on every call
     (``mppi.py:122-123``);
 (b) the "terminal cost" is evaluated on ``path[-1]`` = the LAST SAMPLE's final
     state and added as one scalar to all costs (``mppi.py:79-82``, ``:148``);
 (c) clipped noise is written back into ``eps`` and the clipped ``eps`` is what
     the update weights (``mppi.py:136-139``, ``:117``);
 (d) ``sigma`` is a variance (``mppi.py:18``), controls are normalised by
     ``umax`` (``mppi.py:102``), no ``dt`` factor in the stage cost.

The single deliberate deviation: noise and the constructor's ``act_sequence``
are drawn with trailing dimension ``ctrl_dim`` instead of the hard-coded ``1``
(``mppi.py:22``).  For ``ctrl_dim == 1`` this is identical to the unmodified
reference (same draws, same order); for ``ctrl_dim > 1`` the unmodified
reference raises at ``mppi.py:139``, and this is the documented restatement.
"""
import numpy as np

ACTS = ("relu", "tanh", "sigmoid", "selu")
_SELU_ALPHA = 1.6732632423543772848170429916717
_SELU_SCALE = 1.0507009873554804934193349852946


def _act(name, y):
    if name == "relu":
        return np.maximum(y, 0.0)
    if name == "tanh":
        return np.tanh(y)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-y))
    if name == "selu":
        return _SELU_SCALE * np.where(y > 0, y, _SELU_ALPHA * np.expm1(np.minimum(y, 0.0)))
    raise NotImplementedError(name)


def _act_grad(name, y):
    if name == "relu":
        return (y > 0).astype(np.float64)
    if name == "tanh":
        return 1.0 - np.tanh(y) ** 2
    if name == "sigmoid":
        s = 1.0 / (1.0 + np.exp(-y))
        return s * (1.0 - s)
    if name == "selu":
        return _SELU_SCALE * np.where(y > 0, 1.0, _SELU_ALPHA * np.exp(np.minimum(y, 0.0)))
    raise NotImplementedError(name)


class MLPParams:
    """Plain float64 copy of what ``MLP.get_parameters()`` holds
    (``autompc/sysid/mlp.py:308-313``): ``weights[i]`` is ``(out, in)`` like
    ``torch.nn.Linear``; the last entry is ``output_layer``."""

    def __init__(self, weights, biases, act, xu_mean, xu_std, dy_mean, dy_std, nx, nu):
        self.weights = [np.asarray(w, dtype=np.float64) for w in weights]
        self.biases = [np.asarray(b, dtype=np.float64) for b in biases]
        self.act = act
        self.xu_mean = np.asarray(xu_mean, dtype=np.float64)
        self.xu_std = np.asarray(xu_std, dtype=np.float64)
        self.dy_mean = np.asarray(dy_mean, dtype=np.float64)
        self.dy_std = np.asarray(dy_std, dtype=np.float64)
        self.nx, self.nu = int(nx), int(nu)
        assert self.weights[0].shape[1] == nx + nu and self.weights[-1].shape[0] == nx

    @classmethod
    def from_reference_mlp(cls, mlp):
        """Reads an (unmodified-reference) ``autompc.sysid.mlp.MLP`` instance."""
        sd = mlp.net.state_dict()
        n_hidden = sum(1 for k in sd if k.startswith("layers.") and k.endswith(".weight"))
        ws = [sd["layers.layer%d.weight" % i].cpu().numpy() for i in range(n_hidden)]
        bs = [sd["layers.layer%d.bias" % i].cpu().numpy() for i in range(n_hidden)]
        ws.append(sd["output_layer.weight"].cpu().numpy())
        bs.append(sd["output_layer.bias"].cpu().numpy())
        act = type(mlp.net.nonlin).__name__.lower()
        return cls(ws, bs, act, mlp.xu_means, mlp.xu_std, mlp.dy_means, mlp.dy_std,
                   mlp.system.obs_dim, mlp.system.ctrl_dim)

    @classmethod
    def from_npz(cls, z, prefix=""):
        n = int(z[prefix + "n_layers"])
        return cls([z[prefix + "W%d" % i] for i in range(n)], [z[prefix + "b%d" % i] for i in range(n)],
                   str(z[prefix + "act"]), z[prefix + "xu_mean"], z[prefix + "xu_std"],
                   z[prefix + "dy_mean"], z[prefix + "dy_std"], int(z[prefix + "nx"]), int(z[prefix + "nu"]))

    def to_npz_dict(self, prefix=""):
        d = {prefix + "n_layers": len(self.weights), prefix + "act": self.act,
             prefix + "xu_mean": self.xu_mean, prefix + "xu_std": self.xu_std,
             prefix + "dy_mean": self.dy_mean, prefix + "dy_std": self.dy_std,
             prefix + "nx": self.nx, prefix + "nu": self.nu}
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            d[prefix + "W%d" % i] = w
            d[prefix + "b%d" % i] = b
        return d


def mlp_forward(p, XU, want_pre=False):
    """z-score, layer stack, un-z-score: returns dy (N, nx).  mlp.py:230-235."""
    h = (XU - p.xu_mean) / p.xu_std                       # transform_input  mlp.py:20-24
    pres = []
    for W, b in zip(p.weights[:-1], p.biases[:-1]):       # ForwardNet.forward mlp.py:55-58
        y = h @ W.T + b
        pres.append(y)
        h = _act(p.act, y)
    y = h @ p.weights[-1].T + p.biases[-1]                # output_layer     mlp.py:59
    dy = y * p.dy_std + p.dy_mean                         # transform_output mlp.py:26-30
    return (dy, pres) if want_pre else dy


def mlp_pred_batch(p, state, ctrl):
    """``MLP.pred_batch`` (mlp.py:229-236): state (N,nx), ctrl (N,nu) -> (N,nx)."""
    return state + mlp_forward(p, np.concatenate([state, ctrl], axis=1))


def mlp_pred(p, state, ctrl):
    """``MLP.pred`` (mlp.py:219-227)."""
    return mlp_pred_batch(p, state[None, :], ctrl[None, :])[0]


def mlp_pred_diff_batch(p, state, ctrl):
    """``MLP.pred_diff_batch`` (mlp.py:281-305) in closed form:
    J = diag(dy_std) W_out D_L W_L ... D_1 W_1 diag(1/xu_std); state_jac = J[:, :nx] + I."""
    XU = np.concatenate([state, ctrl], axis=1)
    dy, pres = mlp_forward(p, XU, want_pre=True)
    m = XU.shape[0]
    J = np.broadcast_to(p.weights[0] / p.xu_std[None, :], (m,) + p.weights[0].shape).copy()
    for i, pre in enumerate(pres):
        J = _act_grad(p.act, pre)[:, :, None] * J
        J = np.einsum("oh,mhi->moi", p.weights[i + 1], J)
    J = J * p.dy_std[None, :, None]
    sj = J[:, :, :p.nx] + np.eye(p.nx)[None]
    uj = J[:, :, p.nx:]
    return state + dy, sj, uj


class QuadCostParams:
    """The data of a reference ``QuadCost`` (quad_cost.py:7-51)."""

    def __init__(self, Q, R, F=None, goal=None):
        self.Q = np.array(Q, dtype=np.float64)
        self.R = np.array(R, dtype=np.float64)
        nx = self.Q.shape[0]
        self.F = np.zeros((nx, nx)) if F is None else np.array(F, dtype=np.float64)
        self.goal = np.zeros(nx) if goal is None else np.array(goal, dtype=np.float64)

    # cost.py:66-83
    def eval_obs_cost(self, obs):
        d = obs - self.goal
        return d.T @ self.Q @ d

    # cost.py:118-134
    def eval_ctrl_cost(self, ctrl):
        return ctrl.T @ self.R @ ctrl

    # cost.py:166-183
    def eval_term_obs_cost(self, obs):
        d = obs - self.goal
        return d.T @ self.F @ d

    def obs_cost_batch(self, X):
        D = X - self.goal
        return np.einsum("ki,ij,kj->k", D, self.Q, D)

    def ctrl_cost_batch(self, U):
        return np.einsum("ki,ij,kj->k", U, self.R, U)


class SumQuadCostParams:
    """A reference ``SumCost`` (autompc/costs/sum_cost.py:9-81) whose terms are ``QuadCost`` objects, possibly with
    different goals: every ``eval_*`` is the sum of the terms' values (``_sum_results``, sum_cost.py:52-57)."""

    def __init__(self, terms):
        self.terms = list(terms)

    def eval_obs_cost(self, obs):                       # sum_cost.py:59-60
        return sum(t.eval_obs_cost(obs) for t in self.terms)

    def eval_ctrl_cost(self, ctrl):                     # sum_cost.py:68-69
        return sum(t.eval_ctrl_cost(ctrl) for t in self.terms)

    def eval_term_obs_cost(self, obs):                  # sum_cost.py:77-78
        return sum(t.eval_term_obs_cost(obs) for t in self.terms)

    def obs_cost_batch(self, X):
        return sum(t.obs_cost_batch(X) for t in self.terms)

    def ctrl_cost_batch(self, U):
        return sum(t.ctrl_cost_batch(U) for t in self.terms)


class ThresholdCostParams:
    """``ThresholdCost`` (autompc/costs/thresh_cost.py:8-38): 1 per step where
    ``||obs[a:b] - goal[a:b]||_inf > threshold``; no control or terminal part."""

    def __init__(self, goal, obs_range, threshold):
        self.goal = np.array(goal, dtype=np.float64)
        self.obs_range = (int(obs_range[0]), int(obs_range[1]))
        self.threshold = float(threshold)

    def eval_obs_cost(self, obs):                        # thresh_cost.py:27-32
        a, b = self.obs_range
        return 1.0 if np.linalg.norm(obs[a:b] - self.goal[a:b], np.inf) > self.threshold else 0.0

    def eval_ctrl_cost(self, ctrl):                      # thresh_cost.py:33-34
        return 0.0

    def eval_term_obs_cost(self, obs):                   # thresh_cost.py:36-37
        return 0.0

    def obs_cost_batch(self, X):
        a, b = self.obs_range
        return (np.abs(X[:, a:b] - self.goal[a:b]).max(axis=1) > self.threshold).astype(np.float64)

    def ctrl_cost_batch(self, U):
        return np.zeros(U.shape[0])


class BoxThresholdCostParams:
    """``BoxThresholdCost`` (autompc/costs/thresh_cost.py:40-83): 1 per step where the observation is outside
    ``limits`` (obs_dim, 2); no control or terminal part."""

    def __init__(self, limits):
        self.limits = np.array(limits, dtype=np.float64)

    def eval_obs_cost(self, obs):                        # thresh_cost.py:73-77
        return 1.0 if ((obs < self.limits[:, 0]).any() or (obs > self.limits[:, 1]).any()) else 0.0

    def eval_ctrl_cost(self, ctrl):
        return 0.0

    def eval_term_obs_cost(self, obs):
        return 0.0

    def obs_cost_batch(self, X):
        return ((X < self.limits[:, 0]).any(axis=1) | (X > self.limits[:, 1]).any(axis=1)).astype(np.float64)

    def ctrl_cost_batch(self, U):
        return np.zeros(U.shape[0])


def traj_cost(cost, obs, ctrls):
    """``Cost.__call__`` (autompc/costs/cost.py:27-41): obs (T+1,nx), ctrls (T+1,nu) as a reference Trajectory holds
    them (the last control row is whatever the trajectory holds -- zeros after ``simulate``)."""
    c = 0.0
    for i in range(obs.shape[0]):
        c += cost.eval_obs_cost(obs[i])
        c += cost.eval_ctrl_cost(ctrls[i])
    return c + cost.eval_term_obs_cost(obs[-1])


def linear_pred_batch(A, B, states, ctrls):
    """ARX / Koopman ``pred_batch`` (autompc/sysid/arx.py:151-154, koopman.py:170-173)."""
    return (A @ states.T + B @ ctrls.T).T


def nmpc_constraint(p, H, x):
    """``NonLinearMPCProblem.get_constraint`` (autompc/control/nmpc.py:102-110) for an MLP model: decision vector
    x = [states (H+1, nx) | ctrls (H, nu)], c[i] = -state[i+1] + pred(state[i], ctrl[i])."""
    nx, nu = p.nx, p.nu
    st = x[:(H + 1) * nx].reshape(H + 1, nx)
    ct = x[(H + 1) * nx:].reshape(H, nu)
    return (-st[1:] + mlp_pred_batch(p, st[:H], ct)).reshape(-1)


def nmpc_jacobian(p, H, x):
    """``NonLinearMPCProblem.get_jacobian`` (autompc/control/nmpc.py:148-187): (rows, cols, values) in the
    reference's order: per step the dense state Jacobian, the dense control Jacobian, then -1 on x_{i+1}."""
    nx, nu = p.nx, p.nu
    st = x[:(H + 1) * nx].reshape(H + 1, nx)
    ct = x[(H + 1) * nx:].reshape(H, nu)
    _, Jx, Ju = mlp_pred_diff_batch(p, st[:H], ct)
    rows, cols, vals = [], [], []
    base_u = nx * (H + 1)
    for i in range(H):
        r, c = np.meshgrid(np.arange(nx), np.arange(nx), indexing="ij")
        rows.append(i * nx + r.ravel()); cols.append(i * nx + c.ravel()); vals.append(Jx[i].ravel())
        r, c = np.meshgrid(np.arange(nx), np.arange(nu), indexing="ij")
        rows.append(i * nx + r.ravel()); cols.append(base_u + i * nu + c.ravel()); vals.append(Ju[i].ravel())
        rows.append(i * nx + np.arange(nx)); cols.append((i + 1) * nx + np.arange(nx)); vals.append(-np.ones(nx))
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def model_rmse(p, obs_list, ctrl_list, horizon=1):
    """``get_model_rmse`` (autompc/evaluation/model_metrics.py:12-43) for an MLP (no ``traj_to_states``):
    every window start ``obs[:-horizon]`` is rolled ``horizon`` times through ``pred_batch`` with the recorded
    controls ``ctrls[k:-(horizon-k)]`` and compared with ``obs[horizon:]``;
    rmse = sqrt(mean(all squared errors) * obs_dim)."""
    sq = []
    for obs, ctrls in zip(obs_list, ctrl_list):
        state = obs[:-horizon, :]                                        # model_metrics.py:33
        for k in range(horizon):                                         # :34-35
            state = mlp_pred_batch(p, state, ctrls[k:len(ctrls) - (horizon - k), :])
        sq.append((state - obs[horizon:]) ** 2)                          # :38-40
    sq = np.concatenate(sq)
    return float(np.sqrt(np.mean(sq, axis=None) * p.nx))                 # :42


class MPPIOracle:
    """Restatement of ``autompc.control.mppi.MPPI`` (mppi.py:66-181).

    ``faithful_loop=True`` evaluates the stage cost with the reference's Python
    loop over samples (mppi.py:73-78) -- same numbers, reference-like speed;
    the default is the vectorised form used for parity at larger K.
    """

    def __init__(self, mlp, cost, umin, umax, horizon=20, num_path=1000, sigma=1.0, lmda=1.0,
                 faithful_loop=False, draw_init=True):
        self.mlp, self.cost = mlp, cost
        self.nx, self.nu = mlp.nx, mlp.nu
        self.H, self.num_path = int(horizon), int(num_path)
        self.sigma, self.lmda = sigma, lmda
        self.scale = np.sqrt(sigma)                                # mppi.py:18
        self.umin = np.asarray(umin, dtype=np.float64).reshape(self.nu)
        self.umax = np.asarray(umax, dtype=np.float64).reshape(self.nu)
        self.ctrl_scale = self.umax                                # mppi.py:102
        self.faithful_loop = faithful_loop
        if draw_init:                                              # mppi.py:99
            self.act_sequence = np.random.normal(scale=self.scale, size=(self.H, self.nu))
        else:
            self.act_sequence = np.zeros((self.H, self.nu))
        self.cur_step = 0
        self.last_costs = None
        self.last_eps = None

    def sample_eps(self):
        """mppi.py:126 (draw (K,H,nu) in C order, then transpose to (H,K,nu))."""
        return np.random.normal(scale=self.scale,
                                size=(self.num_path, self.H, self.nu)).transpose((1, 0, 2)).copy()

    def _stage_cost(self, path, u):
        if self.faithful_loop:                                     # mppi.py:73-78
            costs = np.zeros(path.shape[0])
            for i in range(path.shape[0]):
                costs[i] += self.cost.eval_obs_cost(path[i, :self.nx])
                costs[i] += self.cost.eval_ctrl_cost(u[i, :])
            return costs
        return self.cost.obs_cost_batch(path[:, :self.nx]) + self.cost.ctrl_cost_batch(u)

    def do_rollouts(self, cur_state, eps=None):
        """mppi.py:120-152.  ``eps`` (H,K,nu) may be supplied (external-noise
        parity mode of the CUDA engine); otherwise drawn like the reference."""
        self.act_sequence[:-1] = self.act_sequence[1:]             # mppi.py:122
        self.act_sequence[-1] = self.act_sequence[-2]              # mppi.py:123
        eps = self.sample_eps() if eps is None else np.array(eps, dtype=np.float64)
        path = np.zeros((self.num_path, self.nx))
        path[:] = cur_state
        costs = np.zeros(self.num_path)
        action_cost = np.zeros_like(costs)
        lo, hi = self.umin / self.ctrl_scale, self.umax / self.ctrl_scale
        for i in range(self.H):
            actions = eps[i] + self.act_sequence[i]
            actions = np.minimum(hi, np.maximum(lo, actions))      # mppi.py:137-138
            eps[i] = actions - self.act_sequence[i]                # mppi.py:139
            costs += self._stage_cost(path, actions * self.ctrl_scale)          # mppi.py:142
            action_cost += self.lmda / self.sigma * np.einsum("ij,ij->i", actions, eps[i])  # :143
            path = mlp_pred_batch(self.mlp, path, actions * self.ctrl_scale)    # mppi.py:144
        self.term_const = self.cost.eval_term_obs_cost(path[-1, :self.nx])      # mppi.py:79-82
        costs += self.term_const                                   # mppi.py:148
        costs += action_cost                                       # mppi.py:150
        self.last_path = path
        return costs, eps

    def update(self, costs, eps):
        """mppi.py:110-118."""
        S = np.exp(-1 / self.lmda * (costs - np.amin(costs)))
        weight = S / np.sum(S)
        self.act_sequence += np.sum(eps * weight[None, :, None], axis=1)
        self.last_weight = weight

    def solve(self, x0, eps=None):
        costs, eps = self.do_rollouts(x0, eps)
        self.update(costs, eps)
        self.cur_step += 1
        self.last_costs, self.last_eps = costs, eps
        return self.act_sequence[0] * self.ctrl_scale

    def run(self, constate, new_obs, eps=None):
        """mppi.py:154-168 (``MLP.update_state`` returns ``new_obs.copy()``, mlp.py:170-171)."""
        x0 = np.array(new_obs, dtype=np.float64)
        u = self.solve(x0, eps)
        return u.copy(), np.concatenate([x0, u])

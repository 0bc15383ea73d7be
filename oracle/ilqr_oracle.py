"""float64 NumPy restatement of the reference IterativeLQR solve -- TEST INFRASTRUCTURE.

Follows ``/root/reference/autompc/control/ilqr.py``:
``compute_ilqr_default`` ``:100-265`` (init rollout ``:141-149``, backward
Riccati ``:159-187``, batched 10-alpha line search ``:197-225``, Jacobian refresh
``:226-234``, stopping rule ``:235-261``) and ``run`` ``:267-295``; cost
derivatives from ``autompc/costs/cost.py:85-116``, ``:136-164``, ``:185-213``
(note the terminal ``_diff/_hess`` ignore the goal, ``cost.py:194-211`` --
reproduced).  Dynamics derivatives come from ``mlp_pred_diff_batch`` (closed
form of ``autompc/sysid/mlp.py:238-305``).

Besides the reference's return values the oracle records the integer trace the
CUDA engine must match exactly: accepted line-search index per iteration, the
iteration count and the ``converged`` flag.
"""
import numpy as np

from .mppi_oracle import mlp_pred_batch, mlp_pred_diff_batch


def ilqr_solve(mlp, cost, dt, x0, H, ubounds, uguess=None, u_threshold=1e-3, max_iter=50,
               ls_max_iter=10, ls_discount=0.2, ls_cost_threshold=0.3):
    """Returns dict(converged, states, ctrls, Ks, ks, alpha_idx[list], n_iter, obj_trace)."""
    nx, nu = mlp.nx, mlp.nu
    Q, R, F, goal = cost.Q, cost.R, cost.F, cost.goal
    Qs, Rs, Fs = Q + Q.T, R + R.T, F + F.T

    def eval_obj(xs, us):                                           # ilqr.py:124-129
        obj = 0
        for i in range(H):
            obj += dt * (cost.eval_obs_cost(xs[i, :nx]) + cost.eval_ctrl_cost(us[i]))
        obj += cost.eval_term_obs_cost(xs[-1, :nx])
        return obj

    states = np.zeros((H + 1, nx))
    ctrls = np.zeros((H, nu))
    ls_states = np.zeros((ls_max_iter, H + 1, nx))
    ls_ctrls = np.zeros((ls_max_iter, H, nu))
    Ks = np.zeros((H, nu, nx))
    ks = np.zeros((H, nu))
    Jacs = np.zeros((H, nx, nx + nu))
    states[0] = x0
    if uguess is not None:
        ctrls[:] = uguess
    for i in range(H):                                              # ilqr.py:144-147
        xn, jx, ju = mlp_pred_diff_batch(mlp, states[i][None], ctrls[i][None])
        states[i + 1] = xn[0]
        Jacs[i, :, :nx] = jx[0]
        Jacs[i, :, nx:] = ju[0]
    obj = eval_obj(states, ctrls)
    Ct = np.zeros((nx + nu, nx + nu))
    ct = np.zeros(nx + nu)
    converged = False
    alpha_trace, obj_trace = [], [obj]
    n_iter = 0
    ls_fail = False
    for itr in range(max_iter):
        n_iter = itr + 1
        Vn = Fs.copy()                                              # ilqr.py:159-163, cost.py:208-211
        vn = Fs @ states[H, :nx]
        lin_cost_reduce = quad_cost_reduce = 0
        for t in range(H, 0, -1):                                   # ilqr.py:165-187
            Ct[:nx, :nx] = Qs * dt
            Ct[nx:, nx:] = Rs * dt
            ct[:nx] = (Qs @ (states[t - 1, :nx] - goal)) * dt
            ct[nx:] = (Rs @ ctrls[t - 1]) * dt
            J = Jacs[t - 1]
            Qt = Ct + J.T @ Vn @ J
            qt = ct + J.T @ vn
            Ks[t - 1] = -np.linalg.solve(Qt[nx:, nx:], Qt[nx:, :nx])
            ks[t - 1] = -np.linalg.solve(Qt[nx:, nx:], qt[nx:])
            lin_cost_reduce += qt[nx:].dot(ks[t - 1])
            quad_cost_reduce += ks[t - 1] @ Qt[nx:, nx:] @ ks[t - 1]
            Vn = (Qt[:nx, :nx] + Qt[:nx, nx:] @ Ks[t - 1] + Ks[t - 1].T @ Qt[nx:, :nx]
                  + Ks[t - 1].T @ Qt[nx:, nx:] @ Ks[t - 1])
            vn = qt[:nx] + Qt[:nx, nx:] @ ks[t - 1] + Ks[t - 1].T @ (qt[nx:] + Qt[nx:, nx:] @ ks[t - 1])
        ls_success = False
        best_alpha = None
        best_alpha_idx = None
        best_obj = np.inf
        ks_norm = np.linalg.norm(ks)
        alphas = np.array([ls_discount ** i for i in range(ls_max_iter)])
        ls_states[:, 0, :] = x0
        for i in range(H):                                          # ilqr.py:197-205
            for j, alpha in enumerate(alphas):
                ls_ctrls[j, i, :] = alpha * ks[i] + ctrls[i] + Ks[i] @ (ls_states[j, i, :] - states[i, :])
                if ubounds is not None:
                    ls_ctrls[j, i, :] = np.clip(ls_ctrls[j, i, :], ubounds[0], ubounds[1])
            ls_states[:, i + 1, :] = mlp_pred_batch(mlp, ls_states[:, i, :], ls_ctrls[:, i, :])
        new_obj = None
        used_idx = None
        for lsitr, ls_alpha in enumerate(alphas):                   # ilqr.py:208-225
            used_idx = lsitr
            new_states = ls_states[lsitr]
            new_ctrls = ls_ctrls[lsitr]
            new_obj = eval_obj(new_states, new_ctrls)
            expect = ls_alpha * lin_cost_reduce + ls_alpha ** 2 * quad_cost_reduce / 2
            if (obj - new_obj) / (-expect) > ls_cost_threshold:
                best_obj, best_alpha, best_alpha_idx = new_obj, ls_alpha, lsitr
                break
            if new_obj < best_obj:
                best_obj, best_alpha, best_alpha_idx = new_obj, ls_alpha, lsitr
            if ks_norm < u_threshold:
                break
        if best_obj < obj or ks_norm < u_threshold:                 # ilqr.py:228-234
            ls_success = True
            used_idx = best_alpha_idx
            new_ctrls = ls_ctrls[best_alpha_idx]
            new_states = ls_states[best_alpha_idx]
            _, jxs, jus = mlp_pred_diff_batch(mlp, new_states[:-1], new_ctrls)
            Jacs[:, :, :nx] = jxs
            Jacs[:, :, nx:] = jus
            new_obj = eval_obj(new_states, new_ctrls)
        if (not ls_success and new_obj > obj + 1e-3) or best_alpha is None:   # ilqr.py:235-238
            ls_fail = True
            break
        alpha_trace.append(int(used_idx))   # index of the line-search rollout actually adopted
        du_norm = np.linalg.norm(new_ctrls - ctrls)                 # ilqr.py:246
        if du_norm < u_threshold:
            converged = True
        states = np.copy(new_states)
        ctrls = np.copy(new_ctrls)
        obj = new_obj
        obj_trace.append(obj)
        if converged:
            break
    return dict(converged=converged, states=states, ctrls=ctrls, Ks=Ks.copy(), ks=ks.copy(),
                alpha_idx=alpha_trace, n_iter=n_iter, obj_trace=obj_trace, ls_fail=ls_fail)


class ILQROracle:
    """``IterativeLQR.run`` (ilqr.py:267-295) with the default ``reuse_feedback=-1``
    (-> 0: re-solve from ``uguess = 0`` on every step, ``:281-288``)."""

    def __init__(self, mlp, cost, dt, umin, umax, horizon):
        self.mlp, self.cost, self.dt, self.H = mlp, cost, dt, int(horizon)
        self.ubounds = (np.asarray(umin, dtype=np.float64), np.asarray(umax, dtype=np.float64))
        self.last = None

    def run(self, constate, new_obs):
        state = np.array(new_obs, dtype=np.float64)
        r = ilqr_solve(self.mlp, self.cost, self.dt, state, self.H, self.ubounds)
        self.last = r
        u = r["ctrls"][0] + r["Ks"][0] @ (state - r["states"][0])
        return u, np.concatenate([state, u])

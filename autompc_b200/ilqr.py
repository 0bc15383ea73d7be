"""IterativeLQR on the B200 engine -- drop-in for ``autompc.control.ilqr.IterativeLQR``.

Construction and ``run`` follow ``autompc/control/ilqr.py:43-98`` / ``:267-295``
(``reuse_feedback`` semantics included); ``compute_ilqr`` is one CUDA launch of
the float64 solve kernel (``csrc/ilqr.cu``) instead of the Python loops of
``compute_ilqr_default`` (``ilqr.py:100-265``).  Only ``mode=None`` exists in
the reference (``barrier`` / ``auglag`` name methods that are not defined).
"""
import ctypes as C

import numpy as np

from . import _abi
from .mlp import MLPWeights
from .mppi import cost_spec_of
from .plugin import Controller, ControllerFactory


class IterativeLQR(Controller):
    def __init__(self, system, task, model, horizon, reuse_feedback=-1, ubounds=None, mode=None, verbose=False,
                 device=0, max_iter=50, ls_max_iter=10, ls_discount=0.2, ls_cost_threshold=0.3, u_threshold=1e-3):
        super().__init__(system, task, model)
        self.horizon = int(horizon)
        self.dt = system.dt
        if self.dt is None:
            raise ValueError("IterativeLQR needs system.dt (ilqr.py:122)")
        if mode is not None:
            raise Exception("mode has to be None/barrier/auglag; only None is implemented (as in the reference)")
        if reuse_feedback is None or reuse_feedback <= 0:                   # ilqr.py:56-61
            self.reuse_feedback = 0
        elif reuse_feedback > horizon:
            self.reuse_feedback = horizon
        else:
            self.reuse_feedback = reuse_feedback
        if ubounds is None and task.are_ctrl_bounded():                     # ilqr.py:63-67
            bounds = task.get_ctrl_bounds()
            self.ubounds = (bounds[:, 0].copy(), bounds[:, 1].copy())
        else:
            self.ubounds = ubounds
        self.verbose = verbose
        self.device = int(device)
        self.max_iter = int(max_iter)
        self.weights = MLPWeights.from_model(model)
        nx, nu = self.weights.nx, self.weights.nu
        self._mlp_holder = _abi.MlpDescHolder(self.weights)
        lo = self.ubounds[0] if self.ubounds is not None else np.full(nu, -np.inf)
        hi = self.ubounds[1] if self.ubounds is not None else np.full(nu, np.inf)
        # QuadCost, or a SumCost of quadratics folded into one (their gradients / Hessians add up to the folded
        # quadratic's; the terminal ones ignore the goals altogether, cost.py:194-211).  Threshold costs are not
        # differentiable (thresh_cost.py:22-25): iLQR refuses them like the reference's eval_*_diff would.
        spec = cost_spec_of(task.get_cost(), np.stack([lo, hi], axis=1), nx, nu)
        if spec.n_box or not spec.quad:
            raise ValueError("IterativeLQR needs a twice-differentiable cost (QuadCost or a SumCost of them)")
        self._cost_holder = spec.holder
        cfg = _abi.IlqrCfg(self.horizon, nx, nu, float(self.dt), 1 if self.ubounds is not None else 0,
                           self.max_iter, int(ls_max_iter), float(ls_discount), float(ls_cost_threshold),
                           float(u_threshold), self.device)
        h = C.c_void_p()
        _abi.check(_abi.lib().ampc_ilqr_create(C.byref(h), C.byref(cfg), C.byref(self._mlp_holder.desc),
                                               C.byref(self._cost_holder.desc)))
        self._h = h
        self.reset()

    def reset(self):                                                        # ilqr.py:78-82
        self._need_recompute = True
        self._step_count = 0
        self._states = None
        self._guess = None
        self.last_info = None

    @property
    def state_dim(self):
        return self.model.state_dim + self.system.ctrl_dim

    def traj_to_state(self, traj):
        # the reference returns only the model state (ilqr.py:97-98) although run() strips a control
        # suffix (ilqr.py:278-279); harmless for MLP (update_state ignores it).  We follow MPPI's
        # convention so that constate[:-nu] is well defined.
        return np.concatenate([self.model.traj_to_state(traj), traj[-1].ctrl])

    def compute_ilqr(self, state, uguess=None, silent=True):
        """Returns (converged, states, ctrls, Ks, ks) like ``compute_ilqr_default`` (ilqr.py:265)."""
        H, nx, nu = self.horizon, self.weights.nx, self.weights.nu
        x0 = _abi.f64(state, (nx,))
        ug = None if uguess is None else _abi.f64(uguess, (H, nu))
        states, ctrls = np.empty((H + 1, nx)), np.empty((H, nu))
        Ks, ks = np.empty((H, nu, nx)), np.empty((H, nu))
        info = np.zeros(3, dtype=np.int32)
        alpha = np.full(self.max_iter, -1, dtype=np.int32)
        i32 = C.POINTER(C.c_int32)
        _abi.check(_abi.lib().ampc_ilqr_solve_host(self._h, _abi.dptr(x0), None if ug is None else _abi.dptr(ug),
                                                   _abi.dptr(states), _abi.dptr(ctrls), _abi.dptr(Ks), _abi.dptr(ks),
                                                   info.ctypes.data_as(i32), alpha.ctypes.data_as(i32)))
        self.last_info = dict(converged=bool(info[0]), n_iter=int(info[1]), ls_fail=bool(info[2]),
                              alpha_idx=[int(a) for a in alpha if a >= 0])
        return bool(info[0]), states, ctrls, Ks, ks

    def launch_device(self, stream=0):
        """Re-runs the last solve's problem asynchronously on `stream` with no host copies (kernel-only timing)."""
        _abi.check(_abi.lib().ampc_ilqr_launch(self._h, stream))

    def debug_profile(self):
        """SM cycles the last solve spent per phase (debug tap of the kernel)."""
        out = (C.c_uint64 * 8)()
        _abi.check(_abi.lib().ampc_ilqr_debug_profile(self._h, out))
        names = ("setup_init_rollout", "backward", "line_search", "objective_accept", "jacobian", "copy_out", "total",
                 "iterations")
        return dict(zip(names, [int(v) for v in out]))

    def run(self, constate, new_obs, silent=True):                          # ilqr.py:267-295
        nu = self.system.ctrl_dim
        constate = np.asarray(constate)
        state = self.model.update_state(constate[:-nu], constate[-nu:], np.asarray(new_obs, dtype=np.float64))
        if self._need_recompute:
            converged, states, ctrls, Ks, ks = self.compute_ilqr(state, None, silent=silent)
            self._states, self._ctrls, self._gain, self._ks = states, ctrls, Ks, ks
            self._need_recompute = False
            self._step_count = 0
        if self._step_count == self.reuse_feedback:
            self._need_recompute = True
        x0, u0, k0 = self._states[self._step_count], self._ctrls[self._step_count], self._gain[self._step_count]
        u = u0 + k0 @ (state - x0)
        self._step_count += 1
        return u, np.concatenate([state, u])

    def close(self):
        if getattr(self, "_h", None):
            _abi.lib().ampc_ilqr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IterativeLQRFactory(ControllerFactory):
    """``horizon`` in [5, 25], default 20 (ilqr.py:38-43)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.Controller = IterativeLQR
        self.name = "IterativeLQR"

    def get_configuration_space(self):
        from ConfigSpace import ConfigurationSpace
        from ConfigSpace.hyperparameters import UniformIntegerHyperparameter
        cs = ConfigurationSpace()
        cs.add_hyperparameter(UniformIntegerHyperparameter(name="horizon", lower=5, upper=25, default_value=20))
        return cs

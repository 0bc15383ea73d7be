"""Linear dynamics models on the B200 engine -- SURVEY.md 8(f) row 4.

``B200Linear`` is a ``Model`` (same interface as ``autompc/sysid/model.py:55-244``) for the reference's linear model
family: ``ARX`` (``autompc/sysid/arx.py:42-173``) and ``Koopman`` (``autompc/sysid/koopman.py:79-189``) both predict
with ``statesnew = (A @ states.T + B @ ctrls.T).T`` (``arx.py:151-154``, ``koopman.py:170-173``) on a state that is a
history stack / a lifted observation.  ``pred`` / ``pred_batch`` run the float64 CUDA kernel in ``csrc/linear.cu``;
what maps observations to that state (``traj_to_state``, ``update_state``: feature stacking, basis functions) is host
bookkeeping and is delegated to the trained reference model the object was built from.  Training stays in the
reference.  No CPU fallback: without the library / a GPU the prediction calls raise.
"""
import ctypes as C

import numpy as np

from . import _abi
from .plugin import Model


class B200Linear(Model):
    def __init__(self, system, A, B, base=None, device=0):
        Model.__init__(self, system)
        self.A = _abi.f64(A)
        self.B = _abi.f64(B)
        ns = self.A.shape[0]
        if self.A.shape != (ns, ns) or self.B.shape != (ns, system.ctrl_dim):
            raise ValueError("expected A (ns,ns) and B (ns,%d); got %s and %s"
                             % (system.ctrl_dim, self.A.shape, self.B.shape))
        if base is None and ns != system.obs_dim:
            raise ValueError("state_dim %d != obs_dim %d: pass the trained reference model as `base` "
                             "(it maps observations to the model state)" % (ns, system.obs_dim))
        self.base, self.device, self._h = base, int(device), None

    @classmethod
    def from_model(cls, model, device=0):
        """Accepts a trained reference ``ARX`` / ``Koopman`` (anything with ``to_linear()``)."""
        if isinstance(model, B200Linear):
            return model
        if not hasattr(model, "to_linear"):
            raise ValueError("B200Linear needs a linear model (ARX, Koopman); got %s" % type(model).__name__)
        A, B = model.to_linear()
        return cls(model.system, A, B, base=model, device=device)

    def _need(self):
        if self._h is None:
            h = C.c_void_p()
            _abi.check(_abi.lib().ampc_linear_create(C.byref(h), self.A.shape[0], self.B.shape[1], _abi.dptr(self.A),
                                                     _abi.dptr(self.B), self.device))
            self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _abi.lib().ampc_linear_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- Model interface ------------------------------------------------------------------
    @property
    def state_dim(self):
        return self.A.shape[0]

    def traj_to_state(self, traj):
        return self.base.traj_to_state(traj) if self.base is not None else traj[-1].obs.copy()

    def update_state(self, state, new_ctrl, new_obs):
        if self.base is not None:
            return self.base.update_state(state, new_ctrl, new_obs)
        return np.array(new_obs, dtype=np.float64)

    def train(self, trajs, silent=False):
        raise NotImplementedError("training stays in the reference (ARX.train / Koopman.train)")

    def pred_batch(self, states, ctrls):                 # arx.py:151-154, koopman.py:170-173
        self._need()
        X, U = _abi.f64(states), _abi.f64(ctrls)
        ns, nu = self.A.shape[0], self.B.shape[1]
        if X.ndim != 2 or U.ndim != 2 or X.shape[0] != U.shape[0] or X.shape[1] != ns or U.shape[1] != nu:
            raise ValueError("pred_batch expects states (N,%d) and ctrls (N,%d)" % (ns, nu))
        out = np.empty_like(X)
        _abi.check(_abi.lib().ampc_linear_pred_batch(self._h, X.shape[0], _abi.dptr(X), _abi.dptr(U), _abi.dptr(out)))
        return out

    def pred(self, state, ctrl):                         # arx.py:146-149, koopman.py:165-167
        return self.pred_batch(np.asarray(state)[None, :], np.asarray(ctrl)[None, :])[0]

    def pred_diff(self, state, ctrl):                    # arx.py:156-159, koopman.py:175-178
        return self.pred(state, ctrl), np.copy(self.A), np.copy(self.B)

    def pred_diff_batch(self, states, ctrls):            # model.py:155-184 (loop of pred_diff): constant Jacobians
        xn = self.pred_batch(states, ctrls)
        n = xn.shape[0]
        return (xn, np.broadcast_to(self.A, (n,) + self.A.shape).copy(),
                np.broadcast_to(self.B, (n,) + self.B.shape).copy())

    def to_linear(self):                                 # arx.py:161-162, koopman.py:180-181
        return np.copy(self.A), np.copy(self.B)

    def get_parameters(self):                            # koopman.py:183-185
        return {"A": np.copy(self.A), "B": np.copy(self.B)}

    def set_parameters(self, params):                    # koopman.py:187-189
        self.close()
        self.A, self.B = _abi.f64(params["A"]), _abi.f64(params["B"])

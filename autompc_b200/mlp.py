"""MLP dynamics model on the B200 engine.

``MLPWeights`` is the plain data a reference ``autompc.sysid.mlp.MLP`` holds
after training (``get_parameters()``, ``autompc/sysid/mlp.py:308-313``).
``B200MLP`` is a ``Model`` (same interface as ``autompc/sysid/model.py:55-244``)
whose ``pred`` / ``pred_batch`` / ``pred_diff`` / ``pred_diff_batch`` run the
float64 CUDA kernels in ``csrc/mlp_ops.cu``.  Training is out of scope: weights
come from the reference's own ``MLP.train`` (or any ``get_parameters()`` dict).
"""
import ctypes as C

import numpy as np

from . import _abi
from .plugin import Model

_ACT_BY_CLASS = {"relu": "relu", "selu": "selu", "tanh": "tanh", "sigmoid": "sigmoid"}


class MLPWeights:
    """Host float64 copy of the network: ``W[i]`` is (out,in) like torch.nn.Linear."""

    def __init__(self, W, b, act, xu_mean, xu_std, dy_mean, dy_std, nx, nu):
        self.W = [np.ascontiguousarray(np.asarray(w, dtype=np.float64)) for w in W]
        self.b = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in b]
        if act not in _abi.ACT_CODES:
            raise NotImplementedError("Currently supported nonlinearity: relu, tanh, sigmoid, selu")
        self.act = act
        self.nx, self.nu = int(nx), int(nu)
        self.xu_mean = np.asarray(xu_mean, dtype=np.float64).reshape(nx + nu)
        self.xu_std = np.asarray(xu_std, dtype=np.float64).reshape(nx + nu)
        self.dy_mean = np.asarray(dy_mean, dtype=np.float64).reshape(nx)
        self.dy_std = np.asarray(dy_std, dtype=np.float64).reshape(nx)
        self.dims = [self.W[0].shape[1]] + [w.shape[0] for w in self.W]
        if self.dims[0] != nx + nu or self.dims[-1] != nx:
            raise ValueError("MLP dims %s do not match nx+nu=%d -> nx=%d" % (self.dims, nx + nu, nx))
        for i, (w, x) in enumerate(zip(self.W, self.b)):
            if w.shape != (self.dims[i + 1], self.dims[i]) or x.shape != (self.dims[i + 1],):
                raise ValueError("layer %d has inconsistent shapes" % i)

    @classmethod
    def from_state_dict(cls, net_state, act, xu_means, xu_std, dy_means, dy_std, nx, nu):
        """``net_state`` uses ForwardNet's names (mlp.py:39-43): layers.layer{i}.weight/bias, output_layer.*"""
        def arr(t):
            return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        n_hidden = sum(1 for k in net_state if k.startswith("layers.") and k.endswith(".weight"))
        W = [arr(net_state["layers.layer%d.weight" % i]) for i in range(n_hidden)]
        b = [arr(net_state["layers.layer%d.bias" % i]) for i in range(n_hidden)]
        W.append(arr(net_state["output_layer.weight"]))
        b.append(arr(net_state["output_layer.bias"]))
        return cls(W, b, act, xu_means, xu_std, dy_means, dy_std, nx, nu)

    @classmethod
    def from_model(cls, model):
        """Accepts a ``B200MLP``, an ``MLPWeights`` or a trained reference ``autompc.sysid.mlp.MLP``."""
        if isinstance(model, MLPWeights):
            return model
        if isinstance(model, B200MLP):
            return model.weights
        net = getattr(model, "net", None)
        if net is None or not hasattr(model, "xu_means"):
            raise ValueError("the B200 engine supports MLP dynamics only (autompc.sysid.mlp.MLP); got %s"
                             % type(model).__name__)
        act = type(net.nonlin).__name__.lower()
        return cls.from_state_dict(net.state_dict(), _ACT_BY_CLASS.get(act, act), model.xu_means, model.xu_std,
                                   model.dy_means, model.dy_std, model.system.obs_dim, model.system.ctrl_dim)

    @classmethod
    def from_npz(cls, z, prefix=""):
        n = int(z[prefix + "n_layers"])
        return cls([z[prefix + "W%d" % i] for i in range(n)], [z[prefix + "b%d" % i] for i in range(n)],
                   str(z[prefix + "act"]), z[prefix + "xu_mean"], z[prefix + "xu_std"], z[prefix + "dy_mean"],
                   z[prefix + "dy_std"], int(z[prefix + "nx"]), int(z[prefix + "nu"]))


class B200MLP(Model):
    """Drop-in for ``autompc.sysid.mlp.MLP`` at inference time."""

    def __init__(self, system, weights=None, device=0):
        Model.__init__(self, system)
        self.device = device
        self._h = None
        self.weights = None
        if weights is not None:
            self._load(MLPWeights.from_model(weights))

    def _load(self, w):
        if w.nx != self.system.obs_dim or w.nu != self.system.ctrl_dim:
            raise ValueError("weights are for nx=%d nu=%d, system has %d/%d"
                             % (w.nx, w.nu, self.system.obs_dim, self.system.ctrl_dim))
        self._free()
        self.weights = w      # the device handle is created on first use (pred*)

    def _free(self):
        if getattr(self, "_h", None):
            _abi.lib().ampc_mlp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    # --- Model interface (sysid/model.py) -------------------------------------------------
    def traj_to_state(self, traj):                       # mlp.py:167-168
        return traj[-1].obs.copy()

    def update_state(self, state, new_ctrl, new_obs):    # mlp.py:170-171
        return new_obs.copy()

    @property
    def state_dim(self):                                 # mlp.py:173-175
        return self.system.obs_dim

    def train(self, trajs, silent=False, seed=100):
        raise NotImplementedError("training stays in the reference (autompc.sysid.mlp.MLP.train); "
                                  "load its get_parameters() with set_parameters()")

    def _need(self):
        if self.weights is None:
            raise RuntimeError("B200MLP has no parameters; call set_parameters() first")
        if self._h is None:
            w = self.weights
            holder = _abi.MlpDescHolder(w)
            h = C.c_void_p()
            _abi.check(_abi.lib().ampc_mlp_create(C.byref(h), C.byref(holder.desc), w.nx, w.nu, self.device))
            self._h = h

    def pred_batch(self, state, ctrl):                   # mlp.py:229-236
        self._need()
        X = _abi.f64(state)
        U = _abi.f64(ctrl)
        if X.ndim != 2 or U.ndim != 2 or X.shape[0] != U.shape[0] or X.shape[1] != self.weights.nx \
                or U.shape[1] != self.weights.nu:
            raise ValueError("pred_batch expects state (N,%d) and ctrl (N,%d)" % (self.weights.nx, self.weights.nu))
        out = np.empty_like(X)
        _abi.check(_abi.lib().ampc_mlp_pred_batch(self._h, X.shape[0], _abi.dptr(X), _abi.dptr(U), _abi.dptr(out)))
        return out

    def rollout_batch(self, state, ctrls):
        """``pred_batch`` applied ``len(ctrls)`` times in ONE launch: state (N,nx), ctrls (horizon,N,nu) -> (N,nx).
        The inner loop of ``get_model_rmse`` (autompc/evaluation/model_metrics.py:33-35)."""
        self._need()
        X = _abi.f64(state)
        U = _abi.f64(ctrls)
        if X.ndim != 2 or U.ndim != 3 or U.shape[0] < 1 or U.shape[1] != X.shape[0] \
                or X.shape[1] != self.weights.nx or U.shape[2] != self.weights.nu:
            raise ValueError("rollout_batch expects state (N,%d) and ctrls (horizon,N,%d)"
                             % (self.weights.nx, self.weights.nu))
        out = np.empty_like(X)
        _abi.check(_abi.lib().ampc_mlp_rollout_batch(self._h, X.shape[0], U.shape[0], _abi.dptr(X), _abi.dptr(U),
                                                     _abi.dptr(out)))
        return out

    def pred(self, state, ctrl):                         # mlp.py:219-227
        return self.pred_batch(np.asarray(state)[None, :], np.asarray(ctrl)[None, :])[0]

    def pred_diff_batch(self, state, ctrl):              # mlp.py:281-305
        self._need()
        X = _abi.f64(state)
        U = _abi.f64(ctrl)
        if X.ndim != 2 or U.ndim != 2 or X.shape[0] != U.shape[0] or X.shape[1] != self.weights.nx \
                or U.shape[1] != self.weights.nu:
            raise ValueError("pred_diff_batch expects state (N,%d) and ctrl (N,%d)"
                             % (self.weights.nx, self.weights.nu))
        m, nx, nu = X.shape[0], self.weights.nx, self.weights.nu
        xn, jx, ju = np.empty((m, nx)), np.empty((m, nx, nx)), np.empty((m, nx, nu))
        _abi.check(_abi.lib().ampc_mlp_pred_diff_batch(self._h, m, _abi.dptr(X), _abi.dptr(U), _abi.dptr(xn),
                                                       _abi.dptr(jx), _abi.dptr(ju)))
        return xn, jx, ju

    def pred_diff(self, state, ctrl):                    # mlp.py:238-279
        xn, jx, ju = self.pred_diff_batch(np.asarray(state)[None, :], np.asarray(ctrl)[None, :])
        return xn[0], jx[0], ju[0]

    def get_parameters(self):                            # mlp.py:308-313 (same keys)
        w = self.weights
        net_state = {}
        for i in range(len(w.W) - 1):
            net_state["layers.layer%d.weight" % i] = w.W[i].copy()
            net_state["layers.layer%d.bias" % i] = w.b[i].copy()
        net_state["output_layer.weight"] = w.W[-1].copy()
        net_state["output_layer.bias"] = w.b[-1].copy()
        return {"net_state": net_state, "xu_means": w.xu_mean.copy(), "xu_std": w.xu_std.copy(),
                "dy_means": w.dy_mean.copy(), "dy_std": w.dy_std.copy(), "nonlintype": w.act}

    def set_parameters(self, params):                    # mlp.py:315-321
        act = params.get("nonlintype", self.weights.act if self.weights is not None else "relu")
        self._load(MLPWeights.from_state_dict(params["net_state"], act, params["xu_means"], params["xu_std"],
                                              params["dy_means"], params["dy_std"], self.system.obs_dim,
                                              self.system.ctrl_dim))

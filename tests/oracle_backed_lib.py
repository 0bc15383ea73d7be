"""A test double of ``libampc_b200.so`` whose MPPI entry points are answered by the float64 ORACLE -- TEST INFRASTRUCTURE.

The build container has the reference but no GPU; the GPU box has a GPU but no reference.  To run the engine's HOST side
(``autompc_b200.MPPI`` / ``MPPIFactory``: the part that must satisfy the reference's plugin surface) under the
reference's own ``Pipeline`` and ``simulate()`` in the build container, ``tests/test_dropin_cpu.py`` swaps
``autompc_b200._abi._lib`` for this object.  It reads the very ctypes structures the shim passes to the C ABI
(``ampc_mppi_cfg``, ``ampc_mlp_desc``, ``ampc_quad_cost``), so the marshalling is exercised too.  Product code never
imports this module; on a GPU the same host code drives the real library (tests/test_mppi_gpu.py).
"""
import ctypes as C

import numpy as np

from oracle.mppi_oracle import (BoxThresholdCostParams, MLPParams, MPPIOracle, QuadCostParams, SumQuadCostParams)


def _arr(ptr, shape):
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape).copy()


def _obj(x):
    return x._obj if hasattr(x, "_obj") else x


class OracleBackedLib:
    """Implements the subset of include/ampc_b200.h that ``MPPI.run`` with ``noise='numpy'`` touches."""

    def __init__(self):
        self.handles = {}
        self.calls = []
        self._err = b""

    # -- plumbing
    def ampc_last_error(self):
        return self._err

    def ampc_launch_count(self):
        return 0

    def ampc_mppi_create(self, out, cfg, mlp, cost):
        cfg, mlp, cost, out = _obj(cfg), _obj(mlp), _obj(cost), _obj(out)
        nx, nu, n = cfg.nx, cfg.nu, mlp.n_layers
        dims = [mlp.dims[i] for i in range(n + 1)]
        W = [_arr(mlp.W[i], (dims[i + 1], dims[i])) for i in range(n)]
        b = [_arr(mlp.b[i], (dims[i + 1],)) for i in range(n)]
        act = {0: "relu", 1: "tanh", 2: "sigmoid", 3: "selu"}[mlp.act]
        p = MLPParams(W, b, act, _arr(mlp.xu_mean, (nx + nu,)), _arr(mlp.xu_std, (nx + nu,)), _arr(mlp.dy_mean, (nx,)),
                      _arr(mlp.dy_std, (nx,)), nx, nu)
        q = QuadCostParams(_arr(cost.Q, (nx, nx)), _arr(cost.R, (nu, nu)), _arr(cost.F, (nx, nx)), _arr(cost.goal, (nx,)))
        o = MPPIOracle(p, q, _arr(cost.umin, (nu,)), _arr(cost.umax, (nu,)), horizon=cfg.H, num_path=cfg.K,
                       sigma=cfg.sigma, lmda=cfg.lmda, draw_init=False)
        hid = len(self.handles) + 1
        self.handles[hid] = dict(o=o, quad=q, cfg=(cfg.K, cfg.H, nx, nu, cfg.precision))
        out.value = hid
        self.calls.append("create")
        return 0

    def _h(self, h):
        return self.handles[h.value if hasattr(h, "value") else int(h)]

    def ampc_mppi_destroy(self, h):
        self.calls.append("destroy")
        return 0

    def ampc_mppi_set_act_seq(self, h, ptr):
        d = self._h(h)
        K, H, nx, nu, _ = d["cfg"]
        d["o"].act_sequence = _arr(ptr, (H, nu)).astype(np.float32).astype(np.float64)   # the device keeps float32
        return 0

    def ampc_mppi_get_act_seq(self, h, ptr):
        d = self._h(h)
        K, H, nx, nu, _ = d["cfg"]
        np.ctypeslib.as_array(ptr, shape=(H * nu,))[:] = d["o"].act_sequence.ravel()
        return 0

    def ampc_mppi_set_box_costs(self, h, n, lo, hi, w):
        d = self._h(h)
        K, H, nx, nu, _ = d["cfg"]
        lo, hi = _arr(lo, (n, nx)), _arr(hi, (n, nx))
        terms = [d["quad"]] + [BoxThresholdCostParams(np.stack([lo[b], hi[b]], axis=1)) for b in range(n)]
        d["o"].cost = SumQuadCostParams(terms)
        return 0

    def ampc_mppi_solve_host(self, h, x0, eps, seed, counter, u):
        d = self._h(h)
        K, H, nx, nu, _ = d["cfg"]
        if not eps:
            raise AssertionError("the oracle-backed double needs noise='numpy' (host noise in the reference's order)")
        e = _arr(eps, (H, K, nu))
        uo = d["o"].solve(_arr(x0, (nx,)), eps=e)
        np.ctypeslib.as_array(u, shape=(nu,))[:] = uo
        self.calls.append("solve_host")
        return 0

    def ampc_mppi_get_costs(self, h, costs, term):
        d = self._h(h)
        o = d["o"]
        np.ctypeslib.as_array(costs, shape=(d["cfg"][0],))[:] = o.last_costs - o.term_const
        if term:
            _obj(term).value = float(o.term_const)
        return 0

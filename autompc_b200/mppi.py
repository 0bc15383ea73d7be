"""MPPI controller on the B200 engine -- drop-in for ``autompc.control.mppi.MPPI``.

Same construction contract (``Controller(system, task, model, **hyperparams)``,
``autompc/control/controller.py:30-33``), same hyper-parameter names
(``horizon``, ``sigma``, ``lmda``, ``num_path``, ``seed``, ``niter``;
``autompc/control/mppi.py:87-94``), same ``run`` / ``reset`` / ``traj_to_state``
/ ``state_dim`` behaviour (``mppi.py:107-108``, ``:154-176``).  One ``run`` is
one CUDA launch (shift + K rollouts + softmax update) through the C ABI.

Engine-only keyword arguments (all optional):

``noise``      ``"philox"`` (default; in-kernel Philox4x32-10 keyed by
               ``(seed, cur_step, sample, step)``) or ``"numpy"`` (parity mode:
               the noise is drawn on the host from the global NumPy stream in
               exactly the reference's order, ``mppi.py:126``, and uploaded).
``precision``  ``"fp32"`` (CUDA-core FMA) | ``"fp16"`` | ``"bf16"`` (tcgen05 tensor-core kernel
               with IEEE-half / bfloat16 operands, fp32 accumulation, state, cost and
               softmax) | ``"auto"`` (default: fp16 when the shape and the weight range
               allow it, else bf16, else fp32).  The reference computes in float64
               (``autompc/sysid/mlp.py:165``); measured deviation of the updated,
               normalised action sequence from it at the C3 configuration: fp32 1e-5,
               fp16 2.3e-4 (11-bit significands = the operand precision of tf32),
               bf16 3.4e-3; worst parity case 1.2e-3 / 1.1e-2 (tests state 2e-3 / 2e-2).
               Ask for ``"fp32"`` when parity matters more than speed.
``terminal``   ``"reference"`` (default: last sample's terminal cost added to
               all samples, ``mppi.py:79-82``) or ``"per_sample"``.
``device``     CUDA ordinal.
``group``      a ``torch.distributed`` process group: samples are sharded over
               its ranks.
``exchange``   how the ranks' softmax partial records meet: ``"nvlink"`` (default;
               fused into the rollout kernel's tail: peer-memory stores + flags over
               NVLink, one launch per solve, mailboxes shared through CUDA IPC) or
               ``"nccl"`` (rollout kernel, one all-gather, merge kernel).  Falls back to
               ``"nccl"`` on every rank if any rank cannot map its peers.
"""
import ctypes as C

import numpy as np

from . import _abi
from .mlp import MLPWeights
from .plugin import Controller, ControllerFactory


def shard_of(num_path, world, rank):
    """Contiguous shard [k_offset, k_offset + K_local) of the samples owned by `rank`:
    the first num_path % world ranks get one extra sample.  Returns (K_local, k_offset)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank %d/%d" % (world, rank))
    base, rem = divmod(int(num_path), world)
    return base + (1 if rank < rem else 0), rank * base + min(rank, rem)


def fold_quadratic_terms(terms):
    """Folds sum_i (x - g_i)^T M_i (x - g_i) into (x - g)^T M (x - g) + const  (M = sum M_i).

    The linear terms match iff (M + M^T) g = sum_i (M_i + M_i^T) g_i, which always has a solution when the symmetric
    parts are positive semi-definite (range of a sum of PSD matrices = sum of the ranges); the minimum-norm one is
    taken.  Returns (M, g, const).  Raises ValueError if the system is inconsistent (indefinite terms)."""
    M = sum(m for m, _ in terms)
    rhs = sum((m + m.T) @ g for m, g in terms)
    S = M + M.T
    g = np.linalg.pinv(S) @ rhs
    if not np.allclose(S @ g, rhs, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(rhs).max()))):
        raise ValueError("this sum of quadratic costs has no single-quadratic form (indefinite terms)")
    const = float(sum(g_i @ m @ g_i for m, g_i in terms) - g @ M @ g)
    return M, g, const


def box_of_threshold_cost(term, nx):
    """(lo, hi) of the box whose complement a reference threshold cost charges 1 for, or None if `term` is not one.
    ``ThresholdCost`` (autompc/costs/thresh_cost.py:8-32): ||x - goal||_inf > threshold over obs_range, i.e. outside
    goal -+ threshold on those dimensions;  ``BoxThresholdCost`` (:40-77): outside ``limits``."""
    if hasattr(term, "_limits"):
        lim = np.asarray(term._limits, dtype=np.float64).reshape(nx, 2)
        return lim[:, 0].copy(), lim[:, 1].copy()
    if hasattr(term, "_threshold") and hasattr(term, "_obs_range"):
        goal = np.asarray(term._goal, dtype=np.float64).reshape(nx)
        thr = float(np.asarray(term._threshold))
        lo, hi = np.full(nx, -np.inf), np.full(nx, np.inf)
        a, b = int(term._obs_range[0]), int(term._obs_range[1])
        lo[a:b], hi[a:b] = goal[a:b] - thr, goal[a:b] + thr
        return lo, hi
    return None


class CostSpec:
    """A task cost as the engine sees it: ONE quadratic (Q, R, F, goal [, terminal goal]) plus constants common to
    all samples, plus threshold (box) terms.  ``quad`` is False when the cost has no quadratic term at all."""

    def __init__(self, holder, stage_const, term_const, box_lo, box_hi, box_w, quad):
        self.holder, self.stage_const, self.term_const = holder, stage_const, term_const
        self.box_lo, self.box_hi, self.box_w, self.quad = box_lo, box_hi, box_w, quad

    @property
    def n_box(self):
        return len(self.box_w)


def cost_spec_of(cost, bounds, nx, nu):
    """Reads a reference cost object:

    * a ``QuadCost`` (autompc/costs/quad_cost.py:7-51) is taken as is;
    * ``ThresholdCost`` / ``BoxThresholdCost`` (autompc/costs/thresh_cost.py) become box terms evaluated per step by the
      kernels (``ampc_mppi_set_box_costs``);
    * a ``SumCost`` (autompc/costs/sum_cost.py:9-81) of such terms -- e.g. ``QuadCostFactory + GaussRegFactory``, whose
      goals differ so that the reference evaluates the sum term by term -- has its quadratic terms folded
      (``fold_quadratic_terms``): the engine's kernels see one quadratic, the constants ride along on the host
      (they are common to all samples, i.e. they cancel in the MPPI weights, mppi.py:115-116).
    Anything else raises ValueError (no CPU fallback for arbitrary Python costs)."""
    terms = list(cost.costs) if hasattr(cost, "costs") else [cost]
    parts, lo, hi, w = [], [], [], []
    for t in terms:
        box = box_of_threshold_cost(t, nx)
        if box is not None:
            lo.append(box[0]); hi.append(box[1]); w.append(1.0)
            continue
        try:
            parts.append((t.get_cost_matrices(), np.asarray(t.get_goal(), dtype=np.float64)))
        except Exception as e:
            raise ValueError("the B200 engine supports QuadCost, ThresholdCost, BoxThresholdCost and SumCosts of "
                             "them: %s (%s)" % (type(t).__name__, e))
    bounds = np.asarray(bounds, dtype=np.float64)
    box = (np.array(lo).reshape(len(w), nx), np.array(hi).reshape(len(w), nx), np.array(w))
    if not parts:
        z = np.zeros
        return CostSpec(_abi.QuadCostHolder(z((nx, nx)), z((nu, nu)), z((nx, nx)), z(nx), bounds[:, 0], bounds[:, 1], nx, nu),
                        0.0, 0.0, *box, quad=False)
    if len(parts) == 1:
        (Q, R, F), goal = parts[0]
        return CostSpec(_abi.QuadCostHolder(Q, R, F, goal, bounds[:, 0], bounds[:, 1], nx, nu), 0.0, 0.0, *box, quad=True)
    Q, g, c_stage = fold_quadratic_terms([(np.asarray(m[0], dtype=np.float64), gl) for m, gl in parts])
    F, gF, c_term = fold_quadratic_terms([(np.asarray(m[2], dtype=np.float64), gl) for m, gl in parts])
    R = sum(np.asarray(m[1], dtype=np.float64) for m, _ in parts)
    return CostSpec(_abi.QuadCostHolder(Q, R, F, g, bounds[:, 0], bounds[:, 1], nx, nu, goal_term=gF), c_stage, c_term,
                    *box, quad=True)


class MPPI(Controller):
    def __init__(self, system, task, model, **kwargs):
        super().__init__(system, task, model)
        self.kwargs = kwargs
        self.dim_state, self.dim_ctrl = model.state_dim, system.ctrl_dim
        self.seed = int(kwargs.get("seed", 0))
        self.H = int(kwargs.get("horizon", 20))
        self.num_path = int(kwargs.get("num_path", 1000))
        self.num_iter = kwargs.get("niter", 1)
        self.sigma = float(kwargs.get("sigma", 1))
        self.lmda = float(kwargs.get("lmda", 1.0))
        self.noise = kwargs.get("noise", "philox")
        if self.noise not in ("philox", "numpy"):
            raise ValueError("noise must be 'philox' or 'numpy'")
        precision = kwargs.get("precision", "auto")
        terminal = kwargs.get("terminal", "reference")
        if terminal not in ("reference", "per_sample"):
            raise ValueError("terminal must be 'reference' or 'per_sample'")
        self.device = int(kwargs.get("device", 0))
        self.group = kwargs.get("group", None)
        self.exchange = kwargs.get("exchange", "nvlink")
        if self.exchange not in ("nvlink", "nccl"):
            raise ValueError("exchange must be 'nvlink' or 'nccl'")
        self.weights = MLPWeights.from_model(model)
        nx, nu = self.weights.nx, self.weights.nu
        if nx != system.obs_dim or nu != system.ctrl_dim:
            raise ValueError("model dims do not match the system")
        self.umin = np.asarray(task.get_ctrl_bounds(), dtype=np.float64)[:, 0]
        self.umax = np.asarray(task.get_ctrl_bounds(), dtype=np.float64)[:, 1]
        self.ctrl_scale = self.umax                                        # mppi.py:102
        # --- sharding over ranks
        self.world, self.rank = 1, 0
        if self.group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.K_local, self.k_offset = shard_of(self.num_path, self.world, self.rank)
        if self.K_local < 1:
            raise ValueError("num_path=%d cannot be sharded over %d ranks" % (self.num_path, self.world))
        # --- engine handle
        self._mlp_holder = _abi.MlpDescHolder(self.weights)
        self._host_io = None
        self._cost = cost_spec_of(task.get_cost(), task.get_ctrl_bounds(), nx, nu)
        self._cost_holder, self._stage_const, self._term_const = (self._cost.holder, self._cost.stage_const,
                                                                  self._cost.term_const)
        lib = _abi.lib()
        self._h = None
        order = {"auto": ["fp16", "bf16", "fp32"], "fp32": ["fp32"], "bf16": ["bf16"], "fp16": ["fp16"]}.get(precision)
        if order is None:
            raise ValueError("precision must be 'auto', 'fp32', 'fp16' or 'bf16'")
        err = None
        for prec in order:
            cfg = _abi.MppiCfg(self.K_local, self.H, nx, nu, self.sigma, self.lmda,
                               0 if terminal == "reference" else 1, _abi.PREC_CODES[prec], self.k_offset,
                               self.num_path, self.device)
            h = C.c_void_p()
            rc = lib.ampc_mppi_create(C.byref(h), C.byref(cfg), C.byref(self._mlp_holder.desc),
                                      C.byref(self._cost_holder.desc))
            if rc == _abi.AMPC_OK:
                self._h, self.precision = h, prec
                break
            err = rc
            if not (precision == "auto" and rc == _abi.AMPC_ERR_UNSUPPORTED):
                _abi.check(rc)
        if self._h is None:
            _abi.check(err)
        if self._cost.n_box:                                               # thresh_cost.py terms -> per-step predicates
            _abi.check(lib.ampc_mppi_set_box_costs(self._h, self._cost.n_box, _abi.dptr(self._cost.box_lo),
                                                   _abi.dptr(self._cost.box_hi), _abi.dptr(self._cost.box_w)))
        self._dev_bufs = None
        if self.world > 1 and self.exchange == "nvlink":
            self._connect_peers()
        self._init_act_sequence()

    def _connect_peers(self):
        """Export this rank's mailbox over CUDA IPC, gather the handles, map the peers' mailboxes."""
        import torch
        import torch.distributed as dist
        lib = _abi.lib()
        ok = 1
        try:
            mine = C.create_string_buffer(64)
            _abi.check(lib.ampc_mppi_mailbox_ipc(self._h, self.world, mine))
            handles = [None] * self.world
            dist.all_gather_object(handles, mine.raw, group=self.group)
            blob = C.create_string_buffer(b"".join(handles), 64 * self.world)
            _abi.check(lib.ampc_mppi_connect_peers_ipc(self._h, self.world, self.rank, blob))
        except Exception as e:           # all ranks must take the same path: agree below
            ok, self._peer_error = 0, e
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.exchange = "nccl"

    def _init_act_sequence(self):
        # mppi.py:97-99 -- N(0, sqrt(sigma)) draw from the global NumPy stream; trailing dim is
        # ctrl_dim (the reference hard-codes 1, mppi.py:22, and only runs for ctrl_dim == 1).
        self.cur_step = 0
        self.niter = 1                                                     # mppi.py:105
        act = np.random.normal(scale=np.sqrt(self.sigma), size=(self.H, self.dim_ctrl))
        # sharded controller: every rank must roll out around the SAME nominal sequence; rank 0's draw is the one
        self._set_act(self._from_rank0(act))

    def _from_rank0(self, a):
        """Broadcast of a host float64 array from rank 0 of the group (identity when unsharded).  The ranks'
        process-global NumPy streams are not assumed to be in step."""
        if self.world == 1:
            return a
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        on_gpu = dist.get_backend(self.group) == "nccl"
        if on_gpu:
            t = t.to(torch.device("cuda", self.device))
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0), group=self.group)
        return t.cpu().numpy()

    # ------------------------------------------------------------------ engine access ---
    def _set_act(self, act):
        a = _abi.f64(act, (self.H, self.dim_ctrl))
        _abi.check(_abi.lib().ampc_mppi_set_act_seq(self._h, _abi.dptr(a)))

    @property
    def act_sequence(self):
        a = np.empty((self.H, self.dim_ctrl))
        _abi.check(_abi.lib().ampc_mppi_get_act_seq(self._h, _abi.dptr(a)))
        return a

    @act_sequence.setter
    def act_sequence(self, value):
        self._set_act(value)

    def last_costs(self):
        """(costs (K_local,), terminal scalar) of the latest solve -- parity tap."""
        c = np.empty(self.K_local)
        t = C.c_double(0.0)
        _abi.check(_abi.lib().ampc_mppi_get_costs(self._h, _abi.dptr(c), C.byref(t)))
        # constants of a folded SumCost (zero for a plain QuadCost): H stage constants per sample, one terminal constant
        return c + self.H * self._stage_const, t.value + self._term_const

    def philox_noise(self, counter=None):
        """(H, K_local, nu) unclipped noise the Philox path uses for solve `counter`."""
        counter = self.cur_step if counter is None else counter
        e = np.empty((self.H, self.K_local, self.dim_ctrl), dtype=np.float32)
        _abi.check(_abi.lib().ampc_mppi_get_noise(self._h, self.seed, counter,
                                                  e.ctypes.data_as(C.POINTER(C.c_float))))
        return e

    def sample_numpy_noise(self):
        """mppi.py:126: K*H*nu draws in C order over (K,H,nu), transposed to (H,K,nu)."""
        eps = np.random.normal(scale=np.sqrt(self.sigma), size=(self.num_path, self.H, self.dim_ctrl))
        eps = self._from_rank0(eps).transpose((1, 0, 2))     # sharded: rank 0's stream is the one (see _from_rank0)
        return np.ascontiguousarray(eps[:, self.k_offset:self.k_offset + self.K_local, :])

    # ------------------------------------------------------------------ solve ---
    def solve(self, x0, eps=None):
        """One MPPI iteration from observation x0 (mppi.py:158-161); returns the scaled first action."""
        x0 = _abi.f64(x0, (self.dim_state,))
        if eps is None and self.noise == "numpy":
            eps = self.sample_numpy_noise()
        if self.world > 1:
            u = self._solve_sharded(x0, eps)
        elif eps is None:
            # hot path of run(): preallocated buffers and cached ctypes pointers (``a.ctypes.data_as`` builds a new
            # interface object per call: several microseconds next to a 240 us solve)
            io = self._host_io
            if io is None:
                xb, ub = np.empty(self.dim_state), np.empty(self.dim_ctrl)
                io = self._host_io = (xb, ub, _abi.dptr(xb), _abi.dptr(ub), _abi.lib().ampc_mppi_solve_host)
            io[0][:] = x0
            rc = io[4](self._h, io[2], None, self.seed, self.cur_step, io[3])
            if rc:
                _abi.check(rc)
            u = io[1].copy()                    # callee returns fresh arrays (controller.py:61-121)
        else:
            u = np.empty(self.dim_ctrl)
            eps = _abi.f64(eps, (self.H, self.K_local, self.dim_ctrl))
            _abi.check(_abi.lib().ampc_mppi_solve_host(self._h, _abi.dptr(x0), _abi.dptr(eps), self.seed, self.cur_step,
                                                       _abi.dptr(u)))
        self.cur_step += 1
        return u

    def _shard_bufs(self):
        import torch
        if self._dev_bufs is None:
            dev = torch.device("cuda", self.device)
            nrec = _abi.lib().ampc_mppi_record_floats(self._h)
            self._dev_bufs = dict(
                x0=torch.empty(self.dim_state, dtype=torch.float32, device=dev),
                u=torch.empty(self.dim_ctrl, dtype=torch.float32, device=dev),
                rec=torch.empty(nrec, dtype=torch.float32, device=dev),
                recs=torch.empty(self.world * nrec, dtype=torch.float32, device=dev),
                x0_pin=torch.empty(self.dim_state, dtype=torch.float32).pin_memory(),
                u_pin=torch.empty(self.dim_ctrl, dtype=torch.float32).pin_memory())
        return self._dev_bufs

    def solve_device_sharded(self, x0_dev, u_dev, eps_dev=None):
        """Sharded solve on device tensors, asynchronous on torch's current stream.  exchange="nvlink": one
        launch (rollouts + peer-memory exchange + merge).  exchange="nccl": local rollouts -> one all-gather
        of the (min, sum w, sum w*eps) record -> merge kernel on every rank."""
        import torch
        import torch.distributed as dist
        lib = _abi.lib()
        b = self._shard_bufs()
        stream = torch.cuda.current_stream().cuda_stream
        if self.exchange == "nvlink":
            _abi.check(lib.ampc_mppi_solve_fused(self._h, x0_dev.data_ptr(),
                                                 None if eps_dev is None else eps_dev.data_ptr(), self.seed,
                                                 self.cur_step, u_dev.data_ptr(), stream))
            self.cur_step += 1
            return
        _abi.check(lib.ampc_mppi_rollout_partial(self._h, x0_dev.data_ptr(),
                                                 None if eps_dev is None else eps_dev.data_ptr(), self.seed,
                                                 self.cur_step, b["rec"].data_ptr(), stream))
        dist.all_gather_into_tensor(b["recs"], b["rec"], group=self.group)
        _abi.check(lib.ampc_mppi_merge(self._h, b["recs"].data_ptr(), self.world, u_dev.data_ptr(), stream))
        self.cur_step += 1

    def _solve_sharded(self, x0, eps):
        if self.exchange == "nvlink":
            # one launch, no torch on the path: observation in the kernel parameters, merged control in mapped pinned
            # host memory (same as the unsharded hot path)
            io = self._host_io
            if io is None:
                xb, ub = np.empty(self.dim_state), np.empty(self.dim_ctrl)
                io = self._host_io = (xb, ub, _abi.dptr(xb), _abi.dptr(ub), _abi.lib().ampc_mppi_solve_fused_host)
            io[0][:] = x0
            e = None
            if eps is not None:
                eps = _abi.f64(eps, (self.H, self.K_local, self.dim_ctrl))
                e = _abi.dptr(eps)
            rc = io[4](self._h, io[2], e, self.seed, self.cur_step, io[3])
            if rc:
                _abi.check(rc)
            return io[1].copy()
        import torch
        dev = torch.device("cuda", self.device)
        b = self._shard_bufs()
        with torch.cuda.device(dev):
            b["x0_pin"].copy_(torch.from_numpy(x0).to(torch.float32))
            b["x0"].copy_(b["x0_pin"], non_blocking=True)
            e_dev = None
            if eps is not None:
                e_dev = torch.from_numpy(np.ascontiguousarray(eps, dtype=np.float32)).to(dev)
            self.solve_device_sharded(b["x0"], b["u"], e_dev)
            self.cur_step -= 1            # solve() advances the counter
            b["u_pin"].copy_(b["u"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return b["u_pin"].numpy().astype(np.float64)

    def solve_device(self, x0_dev, u_dev, eps_dev=None, stream=0):
        """Asynchronous solve on device float32 tensors/pointers (no host round trip).
        ``x0_dev`` / ``u_dev`` / ``eps_dev`` expose ``data_ptr()`` (torch tensors)."""
        if self.world > 1:
            return self.solve_device_sharded(x0_dev, u_dev, eps_dev)
        _abi.check(_abi.lib().ampc_mppi_solve(self._h, x0_dev.data_ptr(),
                                              None if eps_dev is None else eps_dev.data_ptr(), self.seed,
                                              self.cur_step, u_dev.data_ptr(), stream))
        self.cur_step += 1

    # ------------------------------------------------------------------ Controller API ---
    def run(self, constate, new_obs):                                      # mppi.py:154-168
        nu = self.system.ctrl_dim
        constate = np.asarray(constate)
        x0 = self.model.update_state(constate[:-nu], constate[-nu:], np.asarray(new_obs, dtype=np.float64))
        for _ in range(self.niter):
            ret_action = self.solve(x0)
        statenew = np.concatenate([x0, ret_action])
        return ret_action, statenew

    def reset(self):                                                       # mppi.py:107-108
        self._init_act_sequence()

    def traj_to_state(self, traj):                                         # mppi.py:170-172
        return np.concatenate([self.model.traj_to_state(traj), traj[-1].ctrl])

    @property
    def state_dim(self):                                                   # mppi.py:174-176
        return self.model.state_dim + self.system.ctrl_dim

    @staticmethod
    def is_compatible(system, task, model):                                # mppi.py:178-181
        return True

    def close(self):
        if getattr(self, "_h", None):
            _abi.lib().ampc_mppi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MPPIFactory(ControllerFactory):
    """Same hyper-parameters and ranges as ``autompc.control.mppi.MPPIFactory`` (mppi.py:26-64)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.Controller = MPPI
        self.name = "MPPI"

    def get_configuration_space(self):
        import ConfigSpace as CS
        import ConfigSpace.hyperparameters as CSH
        cs = CS.ConfigurationSpace()
        cs.add_hyperparameter(CSH.UniformIntegerHyperparameter(name="horizon", lower=5, upper=30, default_value=20))
        cs.add_hyperparameter(CSH.UniformFloatHyperparameter(name="sigma", lower=1e-4, upper=2.0, default_value=1.0))
        cs.add_hyperparameter(CSH.UniformFloatHyperparameter(name="lmda", lower=0.1, upper=2.0, default_value=1.0))
        cs.add_hyperparameter(CSH.UniformIntegerHyperparameter(name="num_path", lower=100, upper=1000,
                                                               default_value=200))
        return cs

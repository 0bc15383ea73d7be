"""Device-resident closed loop -- drop-in for ``autompc.utils.simulation.simulate`` on the engine.

``simulate(controller, init_obs, sim_model=..., max_steps=T)`` has the reference's signature
(``autompc/utils/simulation.py:11``) and loop semantics (``:45-63``):  T x [ ``u, constate =
controller.run(constate, obs)`` ; ``obs = sim_model.pred(obs, u)`` ], but the T solves and plant steps are
enqueued on the controller's CUDA stream without any host round trip (one H2D of the initial observation, one
D2H of the trajectory).  ``evaluate_candidates`` runs many such closed loops at once -- the inner loop of the
tuner's ``eval_cfg`` (``autompc/tuning/pipeline_tuner.py:213-239``): candidates are independent, each handle has
its own stream, and over a process group they are dealt round-robin to the ranks (no data-path collective).

Restrictions (ValueError otherwise): the controller is an ``autompc_b200.MPPI`` using in-kernel noise, the
simulation model is an MLP (``B200MLP`` or anything ``MLPWeights.from_model`` accepts), ``term_cond`` and
``dynamics`` are None -- arbitrary Python callbacks need the reference's host loop.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _abi
from .mlp import B200MLP, MLPWeights
from .mppi import MPPI

SimResult = namedtuple("SimResult", ["obs", "ctrls", "cost"])


def _sim_handle(controller, sim_model):
    if not isinstance(sim_model, B200MLP):
        sim_model = B200MLP(controller.system, MLPWeights.from_model(sim_model), device=controller.device)
    sim_model._need()
    return sim_model


def _check(controller, term_cond, dynamics, sim_model):
    if not isinstance(controller, MPPI):
        raise ValueError("the device-resident closed loop drives autompc_b200.MPPI controllers")
    if dynamics is not None or term_cond is not None:
        raise ValueError("dynamics / term_cond callbacks run on the host: use autompc.utils.simulation.simulate")
    if sim_model is None:
        raise ValueError("Must specify dynamics function or simulation model")
    if controller.noise != "philox" or controller.world != 1:
        raise ValueError("the closed loop needs noise='philox' and an unsharded controller")


def _set_eval_cost(controller, cost):
    """The trajectory cost the closed loop accumulates: the controller's own by default; ``cost`` (any cost object
    ``cost_spec_of`` reads: QuadCost / ThresholdCost / BoxThresholdCost / SumCost of them) when the candidate is
    scored with the TASK's cost rather than the one it optimises (tuning/pipeline_tuner.py:230-231)."""
    if cost is None:
        if getattr(controller, "_eval_cost", None) is not None:      # back to the controller's own cost
            spec = controller._cost
            _abi.check(_abi.lib().ampc_mppi_set_eval_cost(
                controller._h, C.byref(spec.holder.desc), spec.n_box, _abi.dptr(spec.box_lo), _abi.dptr(spec.box_hi),
                _abi.dptr(spec.box_w)))
            controller._eval_cost = None
        return
    from .mppi import cost_spec_of
    spec = cost_spec_of(cost, controller.task.get_ctrl_bounds(), controller.dim_state, controller.dim_ctrl)
    _abi.check(_abi.lib().ampc_mppi_set_eval_cost(
        controller._h, C.byref(spec.holder.desc) if spec.quad else None, spec.n_box, _abi.dptr(spec.box_lo),
        _abi.dptr(spec.box_hi), _abi.dptr(spec.box_w)))
    controller._eval_cost = spec


def _start(controller, sim, init_obs, T):
    x0 = _abi.f64(init_obs, (controller.dim_state,))
    _abi.check(_abi.lib().ampc_mppi_closed_loop_start(controller._h, sim._h, _abi.dptr(x0), T, controller.seed,
                                                      controller.cur_step))
    controller.cur_step += T


def _finish(controller, T):
    nx, nu = controller.dim_state, controller.dim_ctrl
    obs, ctrls, cost = np.empty((T + 1, nx)), np.zeros((T + 1, nu)), C.c_double(0.0)
    _abi.check(_abi.lib().ampc_mppi_closed_loop_finish(controller._h, T, _abi.dptr(obs), _abi.dptr(ctrls[:T]),
                                                       C.byref(cost)))
    # constants of a folded SumCost: Cost.__call__ (cost.py:27-41) adds the stage cost of all T+1 states + one terminal
    spec = getattr(controller, "_eval_cost", None) or controller._cost
    total = cost.value + (T + 1) * spec.stage_const + spec.term_const
    return SimResult(obs, ctrls, total)            # like the reference trajectory: last control row is zero


def simulate(controller, init_obs, term_cond=None, dynamics=None, sim_model=None, max_steps=10000, silent=True,
             cost=None):
    """Returns ``SimResult(obs (T+1,nx), ctrls (T+1,nu), cost)``; ``cost`` is ``cost(traj)`` as ``Cost.__call__``
    evaluates it (``autompc/costs/cost.py:27-41``), accumulated on the device in float64, for the ``cost`` argument
    (engine-only; default: the controller's own task cost)."""
    _check(controller, term_cond, dynamics, sim_model)
    sim = _sim_handle(controller, sim_model)
    _set_eval_cost(controller, cost)
    _start(controller, sim, init_obs, int(max_steps))
    return _finish(controller, int(max_steps))


def evaluate_candidates(controllers, init_obs, max_steps, sim_model, group=None, cost=None):
    """Closed-loop evaluation of independent candidate controllers (tuning/pipeline_tuner.py:213-239).
    All closed loops of this rank are in flight together (one stream per controller).  With ``group`` the
    candidates are dealt round-robin over the ranks and the costs are gathered (list of floats, every rank).
    ``cost``: the cost object every trajectory is scored with (the tuner uses the task's, pipeline_tuner.py:230-231);
    default: each controller's own."""
    rank, world = 0, 1
    if group is not None:
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = [i for i in range(len(controllers)) if i % world == rank]
    T = int(max_steps)
    sims = {}
    for i, c in enumerate(controllers):
        if i % world != rank:
            # the reference evaluates candidates one after the other and every reset() draws a new nominal sequence from
            # the global NumPy stream (mppi.py:99, :107-108): consume the draws of the candidates other ranks own, so that
            # ranks seeded alike give every candidate the sequence the sequential loop would have given it
            np.random.normal(scale=np.sqrt(c.sigma), size=(c.H, c.dim_ctrl))
            continue
        _check(c, None, None, sim_model)
        if c.device not in sims:
            sims[c.device] = _sim_handle(c, sim_model)
        c.reset()                                   # pipeline_tuner.py:222
        _set_eval_cost(c, cost)
        _start(c, sims[c.device], init_obs, T)
    results = {i: _finish(controllers[i], T) for i in mine}
    costs = [results[i].cost if i in results else 0.0 for i in range(len(controllers))]
    if group is not None:
        import torch
        import torch.distributed as dist
        if dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", controllers[mine[0]].device) if mine else torch.device("cuda")
        else:
            dev = torch.device("cpu")
        t = torch.tensor(costs, dtype=torch.float64, device=dev)
        dist.all_reduce(t, group=group)             # result gather of len(controllers) scalars, not on the data path
        costs = t.cpu().tolist()
    return costs, results

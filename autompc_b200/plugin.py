"""The reference plugin surface this engine plugs into.

When the real ``autompc`` package is importable its own ABCs are used, so the
controllers here pass ``Pipeline``'s isinstance sorting
(``autompc/pipeline.py:51-81``).  Otherwise (e.g. on a box without the
reference) minimal mirrors with the same names, argument meaning and error
behaviour are defined so the same host code runs:

* ``Controller`` / ``ControllerFactory``  <- ``autompc/control/controller.py:6-121``
* ``Model``                               <- ``autompc/sysid/model.py:55-244``
* ``System``                              <- ``autompc/system.py:3-79``
* ``Task`` (cost + control bounds only)   <- ``autompc/tasks/task.py:5-267``
* ``QuadCost``                            <- ``autompc/costs/quad_cost.py:7-51``, ``cost.py:43-64``
* ``SumCost`` (``+`` of costs)            <- ``autompc/costs/sum_cost.py:9-137``, ``cost.py:213-220``
* ``ThresholdCost`` / ``BoxThresholdCost`` <- ``autompc/costs/thresh_cost.py:8-83``
"""
from abc import ABC, abstractmethod

import numpy as np

try:  # the real thing, when present
    from autompc.control.controller import Controller, ControllerFactory  # type: ignore
    from autompc.sysid.model import Model  # type: ignore
    HAVE_AUTOMPC = True
except Exception:  # pragma: no cover - exercised on boxes without the reference
    HAVE_AUTOMPC = False

    class ControllerFactory(ABC):
        def __init__(self, system, **kwargs):
            self.system = system
            self.kwargs = kwargs

        def __call__(self, cfg, task, model):
            controller_kwargs = cfg.get_dictionary()
            controller_kwargs.update(self.kwargs)
            return self.Controller(self.system, task, model, **controller_kwargs)

        def get_configuration_space(self):
            raise NotImplementedError

    class Controller(ABC):
        def __init__(self, system, task, model):
            self.system = system
            self.model = model
            self.task = task

        @abstractmethod
        def traj_to_state(self, traj):
            raise NotImplementedError

        @abstractmethod
        def run(self, state, new_obs):
            raise NotImplementedError

        def reset(self):
            pass

        @property
        @abstractmethod
        def state_dim(self):
            raise NotImplementedError

    class Model(ABC):
        def __init__(self, system):
            self.system = system

        @abstractmethod
        def traj_to_state(self, traj):
            raise NotImplementedError

        @abstractmethod
        def update_state(self, state, new_ctrl, new_obs):
            raise NotImplementedError

        @abstractmethod
        def pred(self, state, ctrl):
            raise NotImplementedError

        def pred_batch(self, states, ctrls):
            out = np.empty_like(states)
            for i in range(states.shape[0]):
                out[i, :] = self.pred(states[i, :], ctrls[i, :])
            return out

        @abstractmethod
        def pred_diff(self, state, ctrl):
            raise NotImplementedError

        @property
        @abstractmethod
        def state_dim(self):
            raise NotImplementedError

        @property
        def is_linear(self):
            return not getattr(self, "to_linear") is None

        @property
        def is_diff(self):
            return not getattr(self, "pred_diff") is None


try:
    from autompc.costs.thresh_cost import ThresholdCost, BoxThresholdCost  # type: ignore
except Exception:  # pragma: no cover

    class _ThreshBase:
        def eval_ctrl_cost(self, ctrl):                  # thresh_cost.py:33-34, :78-79
            return 0.0

        def eval_term_obs_cost(self, obs):               # thresh_cost.py:36-37, :81-82
            return 0.0

        def get_cost_matrices(self):                     # cost.py:43-51
            raise ValueError("Cost is not quadratic.")

        def __add__(self, other):                        # cost.py:213-220
            more = other.costs if isinstance(other, SumCost) else [other]
            return SumCost(self.system, [self, *more])

        def __call__(self, traj):                        # cost.py:27-41
            c = 0.0
            for i in range(len(traj)):
                c += self.eval_obs_cost(traj[i].obs) + self.eval_ctrl_cost(traj[i].ctrl)
            return c + self.eval_term_obs_cost(traj[-1].obs)

    class ThresholdCost(_ThreshBase):
        """1 per step where ||x - goal||_inf > threshold over obs_range (thresh_cost.py:8-32)."""

        def __init__(self, system, goal, obs_range, threshold):
            self.system = system
            self._goal, self._threshold, self._obs_range = np.copy(goal), np.copy(threshold), obs_range[:]
            self.is_quad, self.has_goal = False, True

        def get_goal(self):
            return np.copy(self._goal)

        def eval_obs_cost(self, obs):
            a, b = self._obs_range[0], self._obs_range[1]
            return 1.0 if np.linalg.norm(obs[a:b] - self._goal[a:b], np.inf) > self._threshold else 0.0

    class BoxThresholdCost(_ThreshBase):
        """1 per step where the observation is outside of limits (obs_dim, 2) (thresh_cost.py:40-77)."""

        def __init__(self, system, limits, goal=None):
            self.system = system
            self._limits = np.copy(limits)
            self.is_quad, self.has_goal = False, goal is not None
            if goal is not None:
                self._goal = np.copy(goal)

        def get_goal(self):
            if not self.has_goal:
                raise ValueError("Cost does not have goal")
            return np.copy(self._goal)

        def eval_obs_cost(self, obs):
            return 1.0 if ((obs < self._limits[:, 0]).any() or (obs > self._limits[:, 1]).any()) else 0.0


try:
    from autompc.system import System  # type: ignore
    from autompc.tasks.task import Task  # type: ignore
    from autompc.costs.quad_cost import QuadCost  # type: ignore
    from autompc.costs.sum_cost import SumCost  # type: ignore
except Exception:  # pragma: no cover

    class SumCost:
        """Sum of cost terms, built with ``+`` (sum_cost.py:9-29, :125-137)."""

        def __init__(self, system, costs):
            self.system = system
            self._costs = list(costs)

        @property
        def costs(self):
            return self._costs[:]

        def get_cost_matrices(self):                     # sum_cost.py:30-44: only for a common goal
            goals = [c.get_goal() for c in self._costs]
            if any(not np.array_equal(goals[0], g) for g in goals[1:]):
                raise NotImplementedError
            mats = [c.get_cost_matrices() for c in self._costs]
            return tuple(sum(m[i] for m in mats) for i in range(3))

        def get_goal(self):
            return self._costs[0].get_goal()

        def __add__(self, other):
            more = other.costs if isinstance(other, SumCost) else [other]
            return SumCost(self.system, [*self._costs, *more])

    class System:
        def __init__(self, observations, controls, dt=None):
            if (len(set(observations)) != len(observations) or len(set(controls)) != len(controls)
                    or set(controls) & set(observations)):
                raise ValueError("Observation and control labels must be unique")
            self._observations, self._controls, self.dt = list(observations), list(controls), dt

        observations = property(lambda self: self._observations[:])
        controls = property(lambda self: self._controls[:])
        obs_dim = property(lambda self: len(self._observations))
        ctrl_dim = property(lambda self: len(self._controls))

    class QuadCost:
        def __init__(self, system, Q, R, F=None, goal=None):
            if Q.shape != (system.obs_dim, system.obs_dim):
                raise ValueError("Q is the wrong shape")
            if R.shape != (system.ctrl_dim, system.ctrl_dim):
                raise ValueError("R is the wrong shape")
            if F is None:
                F = np.zeros((system.obs_dim, system.obs_dim))
            elif F.shape != (system.obs_dim, system.obs_dim):
                raise ValueError("F is the wrong shape")
            self.system = system
            self._Q, self._R, self._F = np.copy(Q), np.copy(R), np.copy(F)
            self._goal = np.zeros(system.obs_dim) if goal is None else np.copy(goal)
            self.is_quad = self.has_goal = True

        def get_cost_matrices(self):
            return np.copy(self._Q), np.copy(self._R), np.copy(self._F)

        def get_goal(self):
            return np.copy(self._goal)

        def __add__(self, other):                        # cost.py:213-220
            more = other.costs if isinstance(other, SumCost) else [other]
            return SumCost(self.system, [self, *more])

    class Task:
        def __init__(self, system):
            self.system = system
            self._ctrl_bounds = np.zeros((system.ctrl_dim, 2))
            self._ctrl_bounds[:, 0], self._ctrl_bounds[:, 1] = -np.inf, np.inf
            self._obs_bounds = np.zeros((system.obs_dim, 2))            # tasks/task.py:21-24
            self._obs_bounds[:, 0], self._obs_bounds[:, 1] = -np.inf, np.inf
            self._init_obs = None
            self._num_steps = None

        def set_cost(self, cost):
            self.cost = cost

        def get_cost(self):
            return self.cost

        def set_obs_bound(self, obs_label, lower, upper):       # tasks/task.py:149-165
            i = self.system.observations.index(obs_label)
            self._obs_bounds[i, :] = [lower, upper]

        def set_obs_bounds(self, lowers, uppers):               # tasks/task.py:167-180
            self._obs_bounds[:, 0], self._obs_bounds[:, 1] = lowers, uppers

        def get_obs_bounds(self):                               # tasks/task.py:245-255
            return self._obs_bounds.copy()

        def are_obs_bounded(self):                              # tasks/task.py:215-228
            return bool(np.any(self._obs_bounds[:, 0] != -np.inf) or np.any(self._obs_bounds[:, 1] != np.inf))

        def set_ctrl_bound(self, ctrl_label, lower, upper):
            i = self.system.controls.index(ctrl_label)
            self._ctrl_bounds[i, :] = [lower, upper]

        def set_ctrl_bounds(self, lowers, uppers):
            self._ctrl_bounds[:, 0], self._ctrl_bounds[:, 1] = lowers, uppers

        def get_ctrl_bounds(self):
            return self._ctrl_bounds.copy()

        def are_ctrl_bounded(self):
            return bool(np.any(self._ctrl_bounds[:, 0] != -np.inf) or np.any(self._ctrl_bounds[:, 1] != np.inf))

        def set_init_obs(self, init_obs):
            self._init_obs = np.array(init_obs)

        def get_init_obs(self):
            return None if self._init_obs is None else self._init_obs.copy()

        def set_num_steps(self, n):
            self._num_steps = n

        def get_num_steps(self):
            return self._num_steps

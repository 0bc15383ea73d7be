"""A small FUNCTIONAL stand-in for the ``ConfigSpace`` package -- TEST INFRASTRUCTURE.

``ConfigSpace`` is not installed here.  ``ref_loader`` registers a permissive dummy that is enough to
import the reference's hot-path modules; running the reference's ``Pipeline`` (``autompc/pipeline.py:90-168``) and the
factories' ``get_configuration_space()`` needs objects that actually hold hyper-parameters.  This module implements
the subset the reference uses on that path: hyper-parameter classes with ``name / lower / upper / default_value /
choices``, ``ConfigurationSpace`` (add / get hyper-parameters, default and sampled ``Configuration``), ``Configuration``
(``get_dictionary``, item access, ``in``), and empty condition / forbidden machinery
(``autompc/utils/cs_utils.py:53-151`` walks them).  Call ``install()`` BEFORE ``ref_loader.load()``.
"""
import sys
import types
from collections import OrderedDict

import numpy as np


class Hyperparameter:
    def __init__(self, name, default_value=None):
        self.name, self.default_value = name, default_value


class NumericalHyperparameter(Hyperparameter):
    def __init__(self, name, lower, upper, default_value=None, log=False, q=None):
        super().__init__(name, default_value)
        self.lower, self.upper, self.log = lower, upper, log
        if default_value is None:
            self.default_value = self._mid()
        if not (lower <= self.default_value <= upper):
            raise ValueError("default %r of %s outside [%r, %r]" % (self.default_value, name, lower, upper))

    def _mid(self):
        return float(np.sqrt(self.lower * self.upper)) if self.log else 0.5 * (self.lower + self.upper)


class FloatHyperparameter(NumericalHyperparameter):
    pass


class UniformFloatHyperparameter(FloatHyperparameter):
    def sample(self, rng):
        if self.log:
            return float(np.exp(rng.uniform(np.log(self.lower), np.log(self.upper))))
        return float(rng.uniform(self.lower, self.upper))


class UniformIntegerHyperparameter(NumericalHyperparameter):
    def _mid(self):
        return int(round(super()._mid()))

    def sample(self, rng):
        return int(rng.integers(self.lower, self.upper + 1))


class CategoricalHyperparameter(Hyperparameter):
    def __init__(self, name, choices, default_value=None):
        super().__init__(name, choices[0] if default_value is None else default_value)
        self.choices = list(choices)

    def sample(self, rng):
        return self.choices[int(rng.integers(len(self.choices)))]


class Constant(Hyperparameter):
    def __init__(self, name, value):
        super().__init__(name, value)
        self.value = value

    def sample(self, rng):
        return self.value


class Configuration:
    def __init__(self, configuration_space, values=None):
        self.configuration_space = configuration_space
        self._values = dict(values or {})

    def get_dictionary(self):
        return dict(self._values)

    def __getitem__(self, k):
        return self._values[k]

    def __setitem__(self, k, v):
        if k not in self.configuration_space._hyperparameters:
            raise KeyError("hyperparameter %r is not in the configuration space" % k)
        self._values[k] = v

    def __contains__(self, k):
        return k in self._values

    def keys(self):
        return self._values.keys()


class ConfigurationSpace:
    def __init__(self, name=None, seed=None):
        self.name = name
        self._hyperparameters = OrderedDict()
        self._conditions = []
        self.forbidden_clauses = []
        self._rng = np.random.default_rng(seed)

    def add_hyperparameter(self, hp):
        if hp.name in self._hyperparameters:
            raise ValueError("hyperparameter %r already in the space" % hp.name)
        self._hyperparameters[hp.name] = hp
        return hp

    def add_hyperparameters(self, hps):
        for hp in hps:
            self.add_hyperparameter(hp)

    def get_hyperparameters(self):
        return list(self._hyperparameters.values())

    def get_hyperparameter_names(self):
        return list(self._hyperparameters.keys())

    def get_hyperparameter(self, name):
        return self._hyperparameters[name]

    def get_conditions(self):
        return list(self._conditions)

    def add_condition(self, c):
        self._conditions.append(c)

    def add_conditions(self, cs):
        self._conditions.extend(cs)

    def add_forbidden_clause(self, c):
        self.forbidden_clauses.append(c)

    def add_forbidden_clauses(self, cs):
        self.forbidden_clauses.extend(cs)

    def get_parents_of(self, hp):
        return []

    def get_default_configuration(self):
        return Configuration(self, {n: hp.default_value for n, hp in self._hyperparameters.items()})

    def sample_configuration(self, size=1):
        out = [Configuration(self, {n: hp.sample(self._rng) for n, hp in self._hyperparameters.items()})
               for _ in range(size)]
        return out[0] if size == 1 else out

    def seed(self, seed):
        self._rng = np.random.default_rng(seed)


class ConditionComponent:
    def get_descendant_literal_conditions(self):
        return [self]


class AbstractCondition(ConditionComponent):
    def __init__(self, child, parent, value=None):
        self.child, self.parent, self.value = child, parent, value


class AbstractConjunction(ConditionComponent):
    pass


class EqualsCondition(AbstractCondition):
    pass


class InCondition(AbstractCondition):
    def __init__(self, child, parent, values):
        super().__init__(child, parent, values)
        self.values = values


class AbstractForbiddenComponent:
    pass


class AbstractForbiddenClause(AbstractForbiddenComponent):
    pass


class AbstractForbiddenConjunction(AbstractForbiddenComponent):
    pass


def install():
    """Registers the stand-in as ``ConfigSpace`` (+ ``.hyperparameters`` / ``.conditions`` / ``.forbidden``)."""
    cs = types.ModuleType("ConfigSpace")
    hp = types.ModuleType("ConfigSpace.hyperparameters")
    cond = types.ModuleType("ConfigSpace.conditions")
    forb = types.ModuleType("ConfigSpace.forbidden")
    cs.ConfigurationSpace, cs.Configuration = ConfigurationSpace, Configuration
    for c in (Hyperparameter, NumericalHyperparameter, FloatHyperparameter, UniformFloatHyperparameter,
              UniformIntegerHyperparameter, CategoricalHyperparameter, Constant):
        setattr(hp, c.__name__, c)
        setattr(cs, c.__name__, c)
    for c in (ConditionComponent, AbstractCondition, AbstractConjunction, EqualsCondition, InCondition):
        setattr(cond, c.__name__, c)
        setattr(cs, c.__name__, c)
    for c in (AbstractForbiddenComponent, AbstractForbiddenClause, AbstractForbiddenConjunction):
        setattr(forb, c.__name__, c)
    cs.hyperparameters, cs.conditions, cs.forbidden = hp, cond, forb
    cs._ampc_functional_shim = True
    sys.modules.update({"ConfigSpace": cs, "ConfigSpace.hyperparameters": hp, "ConfigSpace.conditions": cond,
                        "ConfigSpace.forbidden": forb})
    return cs

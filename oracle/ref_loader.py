"""Import the UNMODIFIED reference hot-path modules from ``/root/reference``.

Test infrastructure (see ``oracle/__init__.py``).  Only usable in the build
container: ``/root/reference`` does not exist on the GPU box, so nothing marked
``gpu``, ``smoke()`` or ``bench.py`` calls this.  It exists to (a) validate the
NumPy restatement in ``oracle/mppi_oracle.py`` / ``oracle/ilqr_oracle.py`` and
(b) generate the golden fixtures (``oracle/make_golden.py``).

``import autompc`` needs ConfigSpace/smac/pysindy/gpytorch, none of which are
installed; the hot-path modules themselves only need ``ConfigSpace`` symbols at
class-definition time.  We register a permissive ConfigSpace stand-in and empty
package shells whose ``__path__`` points into the reference tree, so the
sub-package ``__init__`` files (which pull the absent deps) are bypassed while
every hot-path source file is executed as-is.
"""
import contextlib
import importlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("AMPC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "autompc", "control"))


class _Any:
    """Swallows any ConfigSpace construction at import / ctor time."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


_loaded = None


def load():
    """Returns a namespace with the reference classes (System, Task, QuadCost,
    MLP, MPPI, IterativeLQR, Controller, Model, Trajectory helpers)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "autompc" in sys.modules and not getattr(sys.modules["autompc"], "_ampc_oracle_shell", False):
        raise RuntimeError("a real 'autompc' package is already imported")
    cs = types.ModuleType("ConfigSpace")
    hp = types.ModuleType("ConfigSpace.hyperparameters")
    cond = types.ModuleType("ConfigSpace.conditions")
    cs.ConfigurationSpace = cs.Configuration = _Any
    for n in ("UniformIntegerHyperparameter", "UniformFloatHyperparameter",
              "CategoricalHyperparameter", "Constant"):
        setattr(hp, n, _Any)
        setattr(cs, n, _Any)
    for n in ("InCondition", "EqualsCondition"):
        setattr(cond, n, _Any)
    cs.hyperparameters, cs.conditions = hp, cond
    for name, mod in (("ConfigSpace", cs), ("ConfigSpace.hyperparameters", hp),
                      ("ConfigSpace.conditions", cond)):
        sys.modules.setdefault(name, mod)
    root = os.path.join(REF_ROOT, "autompc")
    pkg = types.ModuleType("autompc")
    pkg.__path__ = [root]
    pkg._ampc_oracle_shell = True
    sys.modules["autompc"] = pkg
    for sub in ("control", "sysid", "costs", "tasks", "utils"):
        m = types.ModuleType("autompc." + sub)
        m.__path__ = [os.path.join(root, sub)]
        sys.modules["autompc." + sub] = m
        setattr(pkg, sub, m)
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):
        ns.System = importlib.import_module("autompc.system").System
        traj = importlib.import_module("autompc.trajectory")
        ns.Trajectory, ns.zeros, ns.extend = traj.Trajectory, traj.zeros, traj.extend
        pkg.zeros, pkg.extend, pkg.System = traj.zeros, traj.extend, ns.System
        ns.Task = importlib.import_module("autompc.tasks.task").Task
        ns.Cost = importlib.import_module("autompc.costs.cost").Cost
        ns.QuadCost = importlib.import_module("autompc.costs.quad_cost").QuadCost
        ns.SumCost = importlib.import_module("autompc.costs.sum_cost").SumCost
        mdl = importlib.import_module("autompc.sysid.model")
        ns.Model, ns.ModelFactory = mdl.Model, mdl.ModelFactory
        ns.MLP = importlib.import_module("autompc.sysid.mlp").MLP
        ctl = importlib.import_module("autompc.control.controller")
        ns.Controller, ns.ControllerFactory = ctl.Controller, ctl.ControllerFactory
        ns.MPPI = importlib.import_module("autompc.control.mppi").MPPI
        ns.IterativeLQR = importlib.import_module("autompc.control.ilqr").IterativeLQR
    _loaded = ns
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints on every constructor / reset (mppi.py:89-96)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield

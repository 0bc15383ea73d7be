"""CPU (build container only): the engine's host side under the UNMODIFIED reference's own ``Pipeline`` and
``simulate()`` -- SURVEY.md 8(a) row a13 / 8(b) "drops into pipeline.py unchanged".

What runs, in a subprocess (so that ``autompc_b200.plugin`` binds to the reference's ABCs):

* ``autompc.pipeline.Pipeline(system, <reference MLP model>, <reference QuadCostFactory>, autompc_b200.MPPIFactory)``
  -- the ``isinstance`` sorting (pipeline.py:51-81), ``get_configuration_space()`` with the ``_ctrlr:`` / ``_cost:``
  prefixes (pipeline.py:90-105), and ``__call__(cfg, task, trajs)`` (pipeline.py:107-168), which builds the engine's
  controller through the reference's ``ControllerFactory.__call__`` (controller.py:30-33);
* ``autompc.utils.simulation.simulate(controller, init_obs, sim_model=model, max_steps=T)`` (simulation.py:11-64)
  driving ``controller.traj_to_state`` / ``run``;
* ``get_configuration_space()`` of ``MPPIFactory`` and ``IterativeLQRFactory`` against the reference factories'
  (names, types, ranges, defaults), and the iLQR factory's sorting into the reference ``Pipeline``;
* the same closed loop with the reference's OWN ``autompc.control.mppi.MPPI`` built by the same pipeline recipe, from
  the same NumPy seed: the two trajectories must agree (the engine draws the noise in the reference's order,
  mppi.py:99, :126).

There is no GPU in the build container and no reference on the GPU box, so here the C library behind
``autompc_b200._abi`` is replaced by ``tests/oracle_backed_lib.py`` (the float64 oracle answering the same C-ABI
calls with the same ctypes structures); everything above the ABI is the shipped code.  On the GPU the same classes
drive the CUDA library (tests/test_mppi_gpu.py::test_closed_loop_through_plugin_surface and the fixture tests).
"""
import os
import subprocess
import sys

import pytest

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, os, importlib, copy
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import configspace_shim
configspace_shim.install()
from oracle import ref_loader
ns = ref_loader.load()
with ref_loader.quiet():
    importlib.import_module("autompc.costs.cost_factory")
    costs_pkg = sys.modules["autompc.costs"]
    costs_pkg.QuadCost = ns.QuadCost                      # quad_cost_factory.py:5 does `from . import QuadCost`
    QuadCostFactory = importlib.import_module("autompc.costs.quad_cost_factory").QuadCostFactory
    Pipeline = importlib.import_module("autompc.pipeline").Pipeline
    simulate = importlib.import_module("autompc.utils.simulation").simulate
    RefMPPIFactory = importlib.import_module("autompc.control.mppi").MPPIFactory

import autompc_b200
from autompc_b200 import _abi, plugin
assert plugin.HAVE_AUTOMPC, "plugin.py must bind to the reference ABCs when autompc is importable"
from tests.oracle_backed_lib import OracleBackedLib
fake = OracleBackedLib()
_abi._lib = fake                                          # the CUDA library's stand-in (no GPU here)

from oracle.make_golden import make_cartpole
from oracle.make_golden_f import reference_mlp_from_npz
z = np.load(os.path.join(%(root)r, "tests", "golden", "cartpole_mlp.npz"))
system, task = make_cartpole(ns)
model = reference_mlp_from_npz(ns, system, z)

def build(factory):
    pipe = Pipeline(system, model, QuadCostFactory(system), factory)
    cs = pipe.get_configuration_space()
    names = cs.get_hyperparameter_names()
    cfg = cs.get_default_configuration()
    cfg["_ctrlr:horizon"] = 9
    cfg["_ctrlr:num_path"] = 150
    cfg["_ctrlr:sigma"] = 0.7
    cfg["_ctrlr:lmda"] = 0.6
    cfg["_cost:theta_Q"] = 12.0
    cfg["_cost:omega_F"] = 3.0
    np.random.seed(42)
    with ref_loader.quiet():
        controller, new_task, m = pipe(cfg, task, [])
    return pipe, names, controller, new_task, m

# ---- the engine's factory inside the reference pipeline
pipe, names, ctl, new_task, m = build(autompc_b200.MPPIFactory(system, noise="numpy", precision="fp32"))
assert pipe.controller_factory is not None and pipe.model is model and pipe.cost_factory is not None
for n in ("_ctrlr:horizon", "_ctrlr:sigma", "_ctrlr:lmda", "_ctrlr:num_path", "_cost:theta_Q", "_cost:u_R", "_cost:dx_F"):
    assert n in names, (n, names)
assert isinstance(ctl, autompc_b200.MPPI) and isinstance(ctl, ns.Controller)
assert ctl.H == 9 and ctl.num_path == 150 and ctl.sigma == 0.7 and ctl.lmda == 0.6 and m is model
Q, R, F = new_task.get_cost().get_cost_matrices()
assert Q[0, 0] == 12.0 and F[1, 1] == 3.0 and ctl.task is new_task
assert ctl.state_dim == 5
with ref_loader.quiet():
    traj = simulate(ctl, task.get_init_obs(), sim_model=model, max_steps=6, silent=True)
assert fake.calls.count("solve_host") == 6 and len(traj) == 7

# ---- the reference's own MPPI through the same recipe, same seed
pipe_r, names_r, ctl_r, _, _ = build(RefMPPIFactory(system))
assert sorted(names_r) == sorted(names)                   # same hyper-parameter names and prefixes
with ref_loader.quiet():
    traj_r = simulate(ctl_r, task.get_init_obs(), sim_model=model, max_steps=6, silent=True)
np.testing.assert_allclose(traj.obs, traj_r.obs, rtol=0, atol=1e-5)     # act_sequence rides through float32 on the device
np.testing.assert_allclose(traj.ctrls, traj_r.ctrls, rtol=0, atol=1e-4)
# reset() redraws like the reference (mppi.py:107-108)
np.random.seed(7); ctl.reset(); a = ctl.act_sequence.copy()
np.random.seed(7)
with ref_loader.quiet():
    ctl_r.reset()
np.testing.assert_allclose(a, ctl_r.act_sequence, rtol=0, atol=1e-7)

# ---- both factories declare the reference's hyper-parameters: same names, types, ranges and defaults
# (mppi.py:26-64, ilqr.py:36-41), and the iLQR factory sorts into the reference pipeline as a controller factory
RefILQRFactory = importlib.import_module("autompc.control.ilqr").IterativeLQRFactory
def hp_table(cs):
    return sorted((h.name, type(h).__name__, h.lower, h.upper, h.default_value) for h in cs.get_hyperparameters())
assert hp_table(autompc_b200.MPPIFactory(system).get_configuration_space()) == \
    hp_table(RefMPPIFactory(system).get_configuration_space())
assert hp_table(autompc_b200.IterativeLQRFactory(system).get_configuration_space()) == \
    hp_table(RefILQRFactory(system).get_configuration_space())
pipe_i = Pipeline(system, model, QuadCostFactory(system), autompc_b200.IterativeLQRFactory(system))
pipe_ir = Pipeline(system, model, QuadCostFactory(system), RefILQRFactory(system))
assert isinstance(pipe_i.controller_factory, autompc_b200.IterativeLQRFactory)
assert sorted(pipe_i.get_configuration_space().get_hyperparameter_names()) == \
    sorted(pipe_ir.get_configuration_space().get_hyperparameter_names())
print("dropin ok", float(np.abs(traj.obs - traj_r.obs).max()))
'''


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_engine_controller_under_reference_pipeline_and_simulate():
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "dropin ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


LINEAR_SCRIPT = r'''
import sys, os
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import ref_loader
from oracle.make_golden import gen_trajs, make_cartpole
from oracle.make_golden_r2 import load_extra
ns = load_extra(ref_loader.load())
import autompc_b200
from autompc_b200 import plugin
assert plugin.HAVE_AUTOMPC
system, task = make_cartpole(ns)
trajs = gen_trajs(ns, system, 8, 60, seed=9)
with ref_loader.quiet():
    arx = ns.ARX(system, history=3)
    arx.train(trajs)
    # (the reference only sets `product_terms` when it arrives as a string, koopman.py:98-99)
    koop = ns.Koopman(system, method="lstsq", poly_basis=True, poly_degree=3, product_terms="false")
    koop.train(trajs)
z = np.load(os.path.join(%(root)r, "tests", "golden", "linear_models.npz"))
assert np.array_equal(arx.A, z["arx_A"]) and np.array_equal(arx.B, z["arx_B"])   # the fixture's model, trained again
for ref in (arx, koop):
    m = autompc_b200.B200Linear.from_model(ref)
    assert isinstance(m, ns.Model) and m.base is ref and m.state_dim == ref.state_dim
    A, B = ref.to_linear()
    assert np.array_equal(m.A, A) and np.array_equal(m.B, B) and m.A.flags.c_contiguous
    for cut in (1, 2, 7, 30):                                  # shorter than the history, equal, longer
        np.testing.assert_array_equal(m.traj_to_state(trajs[1][:cut]), ref.traj_to_state(trajs[1][:cut]))
    st = ref.traj_to_state(trajs[0][:7])
    np.testing.assert_array_equal(m.update_state(st, trajs[0][6].ctrl, trajs[0][7].obs),
                                  ref.update_state(st, trajs[0][6].ctrl, trajs[0][7].obs))
    pa, pb = m.get_parameters()["A"], m.get_parameters()["B"]
    assert np.array_equal(pa, A) and np.array_equal(pb, B)
    # the direct-transcription problem on the model's (stacked / lifted) state: sizes and bounds as the reference's
    task.set_obs_bound("theta", -3.0, 3.5)
    prob = autompc_b200.NonLinearMPCProblem(system, m, task, 5)
    with ref_loader.quiet():
        np.random.seed(0)
        prob_r = ns.NonLinearMPCProblem(system, ref, task, 5)
    assert (prob.dimx, prob.dimc, prob.nnz) == (prob_r.dimx, prob_r.dimc, prob_r.nnz)
    for a, b in zip(prob.get_variable_bounds(), prob_r.get_variable_bounds()):
        assert np.array_equal(a, b)
    r, c = prob.get_jacobian(None, True)
    rr, cr = prob_r.get_jacobian(prob_r._x, True)
    assert np.array_equal(r, rr) and np.array_equal(c, cr)
    x = np.random.default_rng(1).normal(size=prob.dimx)
    np.testing.assert_allclose(prob.get_cost(x), prob_r.get_cost(x), rtol=1e-12)
    np.testing.assert_allclose(prob.get_gradient(x), prob_r.get_gradient(x), rtol=1e-12, atol=1e-13)
    # constant Jacobian values of a linear model, in the reference's sparse order (nmpc.py:170-187)
    np.testing.assert_array_equal(prob.get_jacobian(x, False), prob_r.get_jacobian(x, False))
print("linear ok")
'''


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_linear_models_and_transcription_host_side_under_reference_objects():
    """``B200Linear.from_model(<reference ARX / Koopman>)``: matrices, ``traj_to_state`` / ``update_state`` (history
    stacking, basis lifting: arx.py:64-103, koopman.py:100-118) and the direct-transcription problem's sizes, bounds,
    pattern, cost, gradient and (constant) Jacobian values built on it, against the live unmodified reference objects.
    Host code only: ``pred_batch`` on the device is tests/test_linear_nmpc_gpu.py's."""
    out = subprocess.run([sys.executable, "-c", LINEAR_SCRIPT % {"root": ROOT}], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "linear ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]

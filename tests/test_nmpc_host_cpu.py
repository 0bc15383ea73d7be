"""Host side of the direct-transcription problem (no GPU): ``NonLinearMPCProblem.get_cost`` / ``get_gradient``
(autompc/control/nmpc.py:72-100), ``get_constr_bounds`` / ``get_variable_bounds`` (:112-131) and the sparsity pattern
(:148-169) of the engine's class against fixtures recorded from the UNMODIFIED reference
(``oracle/make_golden_nmpc_host.py``, ``oracle/make_golden_r2.py``).  The two callbacks that run the model on the device
are the GPU suite's (tests/test_linear_nmpc_gpu.py); nothing here touches the library.
Tolerance: float64 on both sides, einsum vs the reference's per-knot Python loop -> rtol 1e-12."""
import os

import numpy as np

from tests.helpers import GOLDEN, load_cartpole


def _problem(z, with_obs_bounds=True):
    from autompc_b200 import B200MLP, NonLinearMPCProblem
    from autompc_b200.mlp import MLPWeights
    from autompc_b200.plugin import QuadCost, System, Task
    p, _, umin, umax, _, dt = load_cartpole()
    system = System(["theta", "omega", "x", "dx"], ["u"])
    system.dt = float(z["dt"]) if "dt" in z.files else dt
    task = Task(system)
    task.set_ctrl_bounds(np.asarray(umin, float), np.asarray(umax, float))
    if with_obs_bounds and "obs_bounds" in z.files:
        task.set_obs_bounds(z["obs_bounds"][:, 0], z["obs_bounds"][:, 1])
    if "Q" in z.files:
        task.set_cost(QuadCost(system, z["Q"], z["R"], z["F"], goal=z["goal"]))
    else:
        task.set_cost(QuadCost(system, np.eye(4), np.eye(1), np.eye(4)))
    w = MLPWeights(p.weights, p.biases, p.act, p.xu_mean, p.xu_std, p.dy_mean, p.dy_std, p.nx, p.nu)
    return NonLinearMPCProblem(system, B200MLP(system, w), task, int(z["H"]))


def test_nmpc_cost_gradient_and_bounds_match_unmodified_reference():
    z = np.load(os.path.join(GOLDEN, "nmpc_host_cartpole_H8.npz"))
    prob = _problem(z)
    assert prob.dimx == z["x"].size
    np.testing.assert_allclose(prob.get_cost(z["x"]), float(z["cost"]), rtol=1e-12, atol=0)
    # the reference's terminal gradient ignores the goal (cost.py:194-199); the fixture's goal is not the origin
    np.testing.assert_allclose(prob.get_gradient(z["x"]), z["gradient"], rtol=1e-12, atol=1e-13)
    lb, ub = prob.get_variable_bounds()
    assert np.array_equal(lb, z["xlb"]) and np.array_equal(ub, z["xub"])
    clb, cub = prob.get_constr_bounds()
    assert np.array_equal(clb, z["clb"]) and np.array_equal(cub, z["cub"])


def test_nmpc_gradient_is_the_derivative_of_the_cost_up_to_the_reference_terminal_quirk():
    """Central differences of get_cost: equal to get_gradient everywhere except the terminal state's block, where the
    reference drops the goal (2 F x instead of 2 F (x - goal)) -- the difference is exactly -2 F goal."""
    z = np.load(os.path.join(GOLDEN, "nmpc_host_cartpole_H8.npz"))
    prob = _problem(z)
    x, H = z["x"], int(z["H"])
    g = prob.get_gradient(x)
    num = np.empty_like(g)
    for i in range(x.size):
        e = np.zeros_like(x)
        e[i] = 1e-5
        num[i] = (prob.get_cost(x + e) - prob.get_cost(x - e)) / 2e-5
    quirk = np.zeros_like(g)
    quirk[H * 4:(H + 1) * 4] = (z["F"] + z["F"].T) @ z["goal"]
    np.testing.assert_allclose(g - quirk, num, rtol=1e-6, atol=1e-7)


def test_nmpc_sparsity_pattern_matches_unmodified_reference_without_a_gpu():
    z = np.load(os.path.join(GOLDEN, "nmpc_cartpole_H8.npz"))
    prob = _problem(z)
    row, col = prob.get_jacobian(None, True)
    assert prob.dimx == int(z["dimx"]) and prob.dimc == int(z["dimc"]) and prob.nnz == z["jac"].size
    assert row.dtype == z["row"].dtype and np.array_equal(row, z["row"]) and np.array_equal(col, z["col"])

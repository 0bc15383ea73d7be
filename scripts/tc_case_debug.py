"""Debug helper: bf16 path vs oracle for a few shapes (prints max relative cost error and its sign)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.mppi_oracle import MPPIOracle, QuadCostParams
from tests.helpers import synthetic_mlp
from tests.test_mppi_gpu import _engine

def run(nx, nu, hidden, K, H, dense, prec="bf16"):
    rng = np.random.default_rng(5)
    p = synthetic_mlp(nx, nu, hidden, act="relu", seed=3)
    if dense:
        A, B, C = rng.normal(size=(nx, nx)), rng.normal(size=(nu, nu)), rng.normal(size=(nx, nx))
        cost = QuadCostParams(A @ A.T / nx, 0.01 * (B @ B.T) / nu, C @ C.T / nx, goal=0.1 * rng.normal(size=nx))
    else:
        cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx), goal=0.05 * rng.normal(size=nx))
    umax = rng.uniform(0.5, 2.0, size=nu); umin = -umax * rng.uniform(0.5, 1.0, size=nu)
    np.random.seed(1)
    ctl = _engine(p, cost, umin, umax, horizon=H, num_path=K, noise="numpy", precision=prec)
    np.random.seed(1)
    o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K)
    x0 = rng.normal(size=nx)
    eps = o.sample_eps()
    ctl.act_sequence = o.act_sequence
    ctl.solve(x0, eps=eps); o.solve(x0, eps=eps.copy())
    c, _ = ctl.last_costs(); ref = o.last_costs - o.term_const
    rel = (c - ref) / np.abs(ref)
    print("nx=%d nu=%d hidden=%s K=%d H=%d dense=%d %s: rel err mean %.2e max|.| %.2e" % (nx, nu, hidden, K, H, dense, prec, rel.mean(), np.abs(rel).max()))
    ctl.close()

for args in [(16, 9, [64], 200, 5, True), (16, 9, [64], 200, 5, False), (16, 9, [64], 200, 1 + 1, True), (16, 9, [64, 64], 200, 5, True),
             (15, 9, [64], 200, 5, True), (16, 6, [64], 200, 5, True), (16, 9, [128], 200, 5, True), (16, 9, [64], 200, 5, True, "fp32"),
             (12, 9, [64], 200, 5, True), (24, 9, [64], 200, 5, True)]:
    run(*args)

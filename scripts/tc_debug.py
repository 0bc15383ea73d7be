"""Debug/report script for the tcgen05 MPPI path: prints errors against the float64 oracle for a
ladder of problems (test infrastructure; uses oracle/)."""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.mppi_oracle import MPPIOracle, QuadCostParams  # noqa: E402
from tests.helpers import synthetic_mlp  # noqa: E402
from tests.gpu_helpers import problem_of  # noqa: E402
from autompc_b200 import MPPI  # noqa: E402


def run_case(tag, nx, nu, hidden, act, K, H, cg, prec="bf16", lmda=1.0, nsolve=2):
    os.environ.pop("AMPC_TC_FORCE_CG", None)
    if cg:
        os.environ["AMPC_TC_FORCE_CG"] = str(cg)
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
    cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx))
    umin, umax = -np.ones(nu), np.ones(nu)
    system, task, model = problem_of(p, cost, umin, umax)
    try:
        np.random.seed(1)
        ctl = MPPI(system, task, model, horizon=H, num_path=K, lmda=lmda, noise="numpy", precision=prec)
        np.random.seed(1)
        o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, lmda=lmda)
        x0 = np.random.default_rng(0).normal(size=nx)
        for s in range(nsolve):
            eps = o.sample_eps()
            ctl.act_sequence = o.act_sequence
            t0 = time.perf_counter()
            u = ctl.solve(x0, eps=eps)
            dt = time.perf_counter() - t0
            uo = o.solve(x0, eps=eps.copy())
            costs, term = ctl.last_costs()
            ref = o.last_costs - o.term_const
            rel = np.abs(costs - ref) / np.abs(ref)
            print("%-28s cg=%s solve %d: cost rel err max %.3e med %.3e | argmin %s | act err %.3e | u err %.3e | term %.5g vs %.5g | %.1f ms"
                  % (tag, cg or "auto", s, rel.max(), np.median(rel), int(np.argmin(costs)) == int(np.argmin(ref)),
                     np.abs(ctl.act_sequence - o.act_sequence).max(), np.abs(u - uo).max(), term, o.term_const, dt * 1e3),
                  flush=True)
            if rel.max() > 0.05:
                bad = np.argsort(-rel)[:5]
                print("   worst samples", bad, costs[bad], ref[bad], flush=True)
        ctl.close()
    except Exception:
        print("%-28s cg=%s FAILED" % (tag, cg or "auto"), flush=True)
        traceback.print_exc()
        return False
    return True


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ok = run_case("fp32 ref 4-1 [32]", 4, 1, [32], "relu", 128, 3, 0, prec="fp32")
    if which in ("all", "cg1"):
        ok = run_case("tiny 4-1 [32] K128 H3", 4, 1, [32], "relu", 128, 3, 1) and ok
        ok = run_case("cartpole [64,64] K256 H20", 4, 1, [64, 64], "relu", 256, 20, 1) and ok
        ok = run_case("17-6 [128,128] K300 H10", 17, 6, [128, 128], "relu", 300, 10, 1) and ok
        ok = run_case("17-6 [64,64] tanh", 17, 6, [64, 64], "tanh", 300, 10, 1) and ok
    if which == "shapes":
        for hid in ([64, 64, 64], [128, 128, 128], [256, 256], [256, 128, 256], [256, 256, 256], [256, 256, 256, 256]):
            run_case("17-6 %s K256 H6" % hid, 17, 6, hid, "relu", 256, 6, 0, nsolve=1)
    if ok and which in ("all", "cg2"):
        run_case("tiny 4-1 [32] K128 H3", 4, 1, [32], "relu", 128, 3, 2)
        run_case("cartpole [64,64] K256 H20", 4, 1, [64, 64], "relu", 256, 20, 2)
        run_case("17-6 [128,128] K300 H10", 17, 6, [128, 128], "relu", 300, 10, 2)
        run_case("C3 [256x3] K2048 H50", 17, 6, [256, 256, 256], "relu", 2048, 50, 0)
        run_case("C3 [256x3] K16384 H50", 17, 6, [256, 256, 256], "relu", 16384, 50, 0, nsolve=1)

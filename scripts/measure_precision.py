"""Measures the deviation of every arithmetic mode of the MPPI rollout kernel from the float64 oracle on the
parity-test cases (tests/test_mppi_gpu.py) and on BASELINE config C3 at full size; prints one JSON line per
(case, precision).  The tolerances stated in the tests are these numbers with head-room."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.mppi_oracle import MLPParams, MPPIOracle, QuadCostParams, mlp_pred_batch  # noqa: E402
from tests.helpers import synthetic_mlp  # noqa: E402
from tests.test_mppi_gpu import TC_CASES, _engine  # noqa: E402


def measure(p, cost, umin, umax, K, H, sigma, lmda, x0, precision, n_solves=3):
    np.random.seed(1)
    ctl = _engine(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda, noise="numpy",
                  precision=precision)
    np.random.seed(1)
    o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda)
    out = dict(act=0.0, cost_rel=0.0, cost_rel_max=0.0, argmin_ok=True)
    for _ in range(n_solves):
        eps = o.sample_eps()
        ctl.act_sequence = o.act_sequence
        u = ctl.solve(x0, eps=eps)
        o.solve(x0, eps=eps.copy())
        costs, term = ctl.last_costs()
        ref = o.last_costs - o.term_const
        out["act"] = max(out["act"], float(np.max(np.abs(ctl.act_sequence - o.act_sequence))))
        out["cost_rel"] = max(out["cost_rel"], float(np.max(np.abs(costs - ref)) / np.abs(ref).max()))
        out["argmin_ok"] = out["argmin_ok"] and int(np.argmin(costs)) == int(np.argmin(ref))
        x0 = mlp_pred_batch(p, x0[None], u[None])[0]
    ctl.close()
    return out


def main():
    precisions = sys.argv[1:] or ["fp32", "fp16", "bf16"]
    for ci, case in enumerate(TC_CASES):
        nx, nu, hidden, act, K, H, sigma, lmda, force_cg = case[:9]
        dense = len(case) > 9 and case[9]
        if force_cg:
            os.environ["AMPC_TC_FORCE_CG"] = force_cg
        else:
            os.environ.pop("AMPC_TC_FORCE_CG", None)
        rng = np.random.default_rng(5)
        p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
        if dense:
            A, B, C = rng.normal(size=(nx, nx)), rng.normal(size=(nu, nu)), rng.normal(size=(nx, nx))
            cost = QuadCostParams(A @ A.T / nx, 0.01 * (B @ B.T) / nu, C @ C.T / nx, goal=0.1 * rng.normal(size=nx))
        else:
            cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx), goal=0.05 * rng.normal(size=nx))
        umax = rng.uniform(0.5, 2.0, size=nu)
        umin = -umax * rng.uniform(0.5, 1.0, size=nu)
        x0 = rng.normal(size=nx)
        for prec in precisions:
            try:
                r = measure(p, cost, umin, umax, K, H, sigma, lmda, x0.copy(), prec)
            except ValueError as e:
                r = dict(error=str(e))
            print(json.dumps(dict(case=ci, dims=[nx, nu, hidden, act, K, H], precision=prec, **r)), flush=True)
    os.environ.pop("AMPC_TC_FORCE_CG", None)
    # BASELINE config C3 at its own size
    from autompc_b200.problems import halfcheetah_dim_problem
    system, task, w, x0 = halfcheetah_dim_problem()
    Q, R, F = task.get_cost().get_cost_matrices()
    p = MLPParams(w.W, w.b, w.act, w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, w.nx, w.nu)
    cost = QuadCostParams(Q, R, F, task.get_cost().get_goal())
    b = task.get_ctrl_bounds()
    for prec in precisions:
        if prec == "fp32" and os.environ.get("AMPC_SKIP_FP32_C3"):
            continue
        r = measure(p, cost, b[:, 0], b[:, 1], 16384, 50, 1.0, 1.0, x0.copy(), prec, n_solves=2)
        print(json.dumps(dict(case="C3 K=16384 H=50", precision=prec, **r)), flush=True)


def fixtures(precisions):
    """The unmodified reference's recorded MPPI runs (ctrl_dim == 1, trained cartpole MLP) on every arithmetic mode."""
    from tests.helpers import GOLDEN, load_cartpole
    mlp, cost, umin, umax, _, _ = load_cartpole()
    for name in ["mppi_cartpole_K256_H20", "mppi_cartpole_K100_H5", "mppi_cartpole_K512_H30", "mppi_cartpole_K4096_H30"]:
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        for prec in precisions:
            np.random.seed(int(z["seed"]))
            ctl = _engine(mlp, cost, umin, umax, horizon=int(z["H"]), num_path=int(z["K"]), sigma=float(z["sigma"]),
                          lmda=float(z["lmda"]), noise="numpy", precision=prec)
            out = dict(act=0.0, u=0.0, cost_abs_rel=0.0, cost_diff_rel=0.0, argmin_ok=True, free_run_act=0.0)
            constate = np.zeros(5)
            for s_ in range(int(z["n_steps"])):
                if s_ > 0:
                    out["free_run_act"] = max(out["free_run_act"], float(np.abs(ctl.act_sequence - z["act_%d" % (s_ - 1)]).max()))
                    ctl.act_sequence = z["act_%d" % (s_ - 1)]
                u, constate = ctl.run(constate, z["x0_%d" % s_])
                costs, term = ctl.last_costs()
                ref = z["costs_%d" % s_]
                out["cost_abs_rel"] = max(out["cost_abs_rel"], float(np.max(np.abs(costs + term - ref) / np.abs(ref))))
                out["cost_diff_rel"] = max(out["cost_diff_rel"], float(np.max(np.abs((costs - costs.min()) - (ref - ref.min()))) / np.abs(ref).max()))
                out["argmin_ok"] = out["argmin_ok"] and int(np.argmin(costs)) == int(z["argmin_%d" % s_])
                out["act"] = max(out["act"], float(np.abs(ctl.act_sequence - z["act_%d" % s_]).max()))
                out["u"] = max(out["u"], float(np.abs(u - z["u_%d" % s_]).max()))
            ctl.close()
            print(json.dumps(dict(case=name, precision=prec, **out)), flush=True)


if __name__ == "__main__":
    fixtures(sys.argv[1:] or ["fp32", "fp16", "bf16"])
    main()

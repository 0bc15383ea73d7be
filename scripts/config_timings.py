"""Latency / throughput of the other BASELINE.json configurations on one B200 (the cartpole-sized
problems are latency numbers, neither roofline applies).  Writes gpurun_out/config_timings.json.
   C2: cartpole MPPI K=4096 H=30 MLP[2x64]      -> us per solve (device events) and through MPPI.run
   C4: cartpole IterativeLQR H=50                -> ms per solve (one launch, float64)
   C5: 64 candidate MPPI controllers x 200-step closed loops (device resident, all in flight) -> candidate evals / s
"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autompc_b200 import MPPI, IterativeLQR, B200MLP, evaluate_candidates
from autompc_b200.mlp import MLPWeights
from autompc_b200.problems import cartpole_problem

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cartpole_mlp.npz"))
system, task, w, x0 = cartpole_problem(MLPWeights.from_npz(z))
model = B200MLP(system, w)
out = {}
dev = torch.device("cuda", 0)

# ---- C2
np.random.seed(0)
for prec in ("bf16", "fp32"):
    ctl = MPPI(system, task, model, horizon=30, num_path=4096, precision=prec)
    x0d = torch.tensor(x0, dtype=torch.float32, device=dev); ud = torch.zeros(1, dtype=torch.float32, device=dev)
    for _ in range(20): ctl.solve_device(x0d, ud)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(500): ctl.solve_device(x0d, ud)
    t1.record(); torch.cuda.synchronize()
    constate = np.zeros(5)
    for _ in range(20): ctl.run(constate, x0)
    t = time.perf_counter()
    for _ in range(500): u, constate = ctl.run(constate, x0)
    e2e = (time.perf_counter() - t) / 500
    out["C2_mppi_K4096_H30_%s" % prec] = {"us_per_solve_device": 1e3 * t0.elapsed_time(t1) / 500, "us_per_solve_run": 1e6 * e2e}
    ctl.close()

# ---- C4
il = IterativeLQR(system, task, model, horizon=50)
for _ in range(3): il.compute_ilqr(x0)
t = time.perf_counter()
for _ in range(20): conv, *_ = il.compute_ilqr(x0)
out["C4_ilqr_H50"] = {"ms_per_solve": 1e3 * (time.perf_counter() - t) / 20, "iterations": int(il.last_info["n_iter"]),
                      "converged": bool(conv)}
il.close()

# ---- C5: 64 candidates from the reference's ranges (mppi.py:52-63), 200 closed-loop steps each
rng = np.random.default_rng(100)
cfgs = [dict(horizon=int(rng.integers(5, 31)), num_path=int(rng.integers(100, 1001)), sigma=float(rng.uniform(1e-4, 2.0)),
             lmda=float(rng.uniform(0.1, 2.0)), seed=i) for i in range(64)]
np.random.seed(1)
ctls = [MPPI(system, task, model, **c) for c in cfgs]
evaluate_candidates(ctls[:4], x0, 10, model)                     # warm-up
torch.cuda.synchronize()
t = time.perf_counter()
costs, _ = evaluate_candidates(ctls, x0, 200, model)
dt = time.perf_counter() - t
out["C5_64_candidates_x_200_steps"] = {"seconds": dt, "candidate_evals_per_s": 64 / dt, "mpc_steps_per_s": 64 * 200 / dt,
                                       "finite_costs": int(np.isfinite(costs).sum())}
for c in ctls: c.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/config_timings.json", "w"), indent=1)
print(json.dumps(out, indent=1))

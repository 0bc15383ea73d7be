"""One cartpole iLQR solve (BASELINE config C4) + the kernel's per-phase cycle counts; for ncu / profiling."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autompc_b200 import IterativeLQR, B200MLP
from autompc_b200.mlp import MLPWeights
from autompc_b200.problems import cartpole_problem
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cartpole_mlp.npz"))
system, task, w, x0 = cartpole_problem(MLPWeights.from_npz(z))
il = IterativeLQR(system, task, B200MLP(system, w), horizon=int(os.environ.get("H", "50")))
for _ in range(2):
    conv, *_ = il.compute_ilqr(x0)
p = il.debug_profile()
print(json.dumps(dict(converged=bool(conv), info=il.last_info, cycles=p,
                      share={k: round(v / max(p["total"], 1), 3) for k, v in p.items() if k not in ("total", "iterations")})))

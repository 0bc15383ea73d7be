"""One launch each of the float64 batch kernels at the C3 network size, for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:'blocked_kernel|pred_diff_kernel' -c 3 \
       -f -o gpurun_out/mlp_f64 python scripts/mlp_ncu_driver.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from autompc_b200 import B200MLP
from autompc_b200.problems import halfcheetah_dim_problem

system, task, w, x0 = halfcheetah_dim_problem()
m = B200MLP(system, w)
rng = np.random.default_rng(0)
X = rng.normal(size=(65536, 17)); U = 0.3 * rng.normal(size=(20, 65536, 6))
m.pred_batch(X, U[0])                              # pred_batch_blocked_kernel, batch 65536
m.rollout_batch(X[:8192], U[:, :8192])             # rollout_batch_blocked_kernel, batch 8192 x horizon 20
m.pred_diff_batch(X[:400], U[0, :400])             # pred_diff_kernel, batch 400
print("done")

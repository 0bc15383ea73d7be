import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from autompc_b200 import IterativeLQR, B200MLP
from autompc_b200.mlp import MLPWeights
from autompc_b200.problems import cartpole_problem
z = np.load("/root/repo/tests/golden/cartpole_mlp.npz")
system, task, w, x0 = cartpole_problem(MLPWeights.from_npz(z))
il = IterativeLQR(system, task, B200MLP(system, w), horizon=50)
il.compute_ilqr(x0); il.compute_ilqr(x0)

"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container (needs ``/root/reference``):

    python -m oracle.make_golden

Every fixture is produced by executing the reference's own classes
(``autompc.control.mppi.MPPI``, ``autompc.sysid.mlp.MLP``,
``autompc.control.ilqr.IterativeLQR``, ``autompc.costs.quad_cost.QuadCost``)
through ``oracle/ref_loader.py``; no number in a fixture comes from the oracle
restatement.  The noise is NOT stored: MPPI uses the global legacy NumPy stream
(``mppi.py:23``), which is frozen across NumPy versions, so tests replay it from
the recorded seed (ctor consumes H draws, ``mppi.py:99``; each solve K*H draws
in C order over (K,H,1), ``mppi.py:126``).
"""
import contextlib
import io
import os
import re
import sys

import numpy as np
import torch

from . import ref_loader
from .mppi_oracle import MLPParams

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CART_Q = np.diag([30.0, 3.0, 0.005, 0.1])       # examples/3_Controllers_and_Tasks.ipynb cell 6
CART_R = np.diag([0.02])
CART_F = np.diag([2.0, 3000, 0.15, 0.3])
CART_X0 = np.array([3.1, 0.0, 0.0, 0.0])        # benchmarks/cartpole.py:55


def cartpole_step(y, u, dt=0.05, g=9.8, m=1.0, L=1.0, b=1.0):
    """Synthetic-data source: explicit-Euler cartpole (same physical model as
    the reference benchmark, benchmarks/cartpole.py:17-36; re-derived here
    because that module imports matplotlib)."""
    theta, omega, x, dx = y
    f = np.array([omega, g * np.sin(theta) / L - b * omega / (m * L ** 2) + u[0] * np.cos(theta) / L, dx, u[0]])
    return y + dt * f


def make_cartpole(ns):
    system = ns.System(["theta", "omega", "x", "dx"], ["u"])
    system.dt = 0.05
    task = ns.Task(system)
    task.set_ctrl_bound("u", -20.0, 20.0)
    task.set_cost(ns.QuadCost(system, CART_Q, CART_R, CART_F, goal=np.zeros(4)))
    task.set_init_obs(CART_X0)
    task.set_num_steps(200)
    return system, task


def gen_trajs(ns, system, n_trajs, traj_len, seed):
    rng = np.random.default_rng(seed)
    trajs = []
    for _ in range(n_trajs):
        y = np.array([rng.uniform(lo, hi) for lo, hi in zip([-1.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0])])
        y[0] += np.pi * rng.integers(0, 2)
        traj = ns.zeros(system, traj_len)
        for i in range(traj_len):
            traj[i].obs[:] = y
            u = rng.uniform(-20.0, 20.0, 1)
            traj[i].ctrl[:] = u
            y = cartpole_step(y, u)
        trajs.append(traj)
    return trajs


def train_cartpole_mlp(ns, system):
    trajs = gen_trajs(ns, system, 150, 200, seed=100)
    with ref_loader.quiet():
        mlp = ns.MLP(system, n_hidden_layers=2, hidden_size=64, nonlintype="relu",
                     n_train_iters=25, use_cuda=False)
        mlp.train(trajs, silent=True)
    return mlp


def random_mlp(ns, system, n_hidden, hidden, act, seed):
    """Untrained reference MLP with injected normalisers."""
    with ref_loader.quiet():
        mlp = ns.MLP(system, n_hidden_layers=n_hidden, hidden_size=hidden, nonlintype=act,
                     use_cuda=False, seed=seed)
    rng = np.random.default_rng(seed)
    n_in = system.obs_dim + system.ctrl_dim
    mlp.xu_means = rng.normal(size=n_in)
    mlp.xu_std = rng.uniform(0.5, 2.0, size=n_in)
    mlp.dy_means = 0.01 * rng.normal(size=system.obs_dim)
    mlp.dy_std = rng.uniform(0.01, 0.1, size=system.obs_dim)
    mlp.net.eval()
    for p in mlp.net.parameters():
        p.requires_grad_(False)
    return mlp


def golden_mppi(ns, system, task, mlp, K, H, seed, n_steps, sigma=1.0, lmda=1.0, keep_costs=True, x_init=None):
    """Runs reference ``MPPI`` closed-loop on the cartpole and records per step
    what ``run`` computes (by calling ``do_rollouts`` / ``update`` exactly as
    ``run`` does, mppi.py:158-166)."""
    np.random.seed(seed)
    with ref_loader.quiet():
        ctl = ns.MPPI(system, task, mlp, horizon=H, num_path=K, sigma=sigma, lmda=lmda)
    out = {"K": K, "H": H, "seed": seed, "sigma": sigma, "lmda": lmda, "n_steps": n_steps,
           "act0": ctl.act_sequence.copy()}
    x = CART_X0.copy() if x_init is None else np.array(x_init, dtype=np.float64)
    constate = np.concatenate([x, np.zeros(1)])
    for s in range(n_steps):
        out["x0_%d" % s] = x.copy()
        x0 = ctl.model.update_state(constate[:-1], constate[-1:], x)
        costs, eps = ctl.do_rollouts(x0, ctl.seed + ctl.cur_step)
        ctl.update(costs, eps)
        ctl.cur_step += 1
        u = ctl.act_sequence[0].copy() * ctl.ctrl_scale
        constate = np.concatenate([x0, u])
        if keep_costs:
            out["costs_%d" % s] = costs.copy()
        out["costs_min_%d" % s] = costs.min()
        out["argmin_%d" % s] = int(np.argmin(costs))
        out["eps_clip_sum_%d" % s] = eps.sum(axis=1)            # (H,1) checksum of the clipped noise
        out["act_%d" % s] = ctl.act_sequence.copy()
        out["u_%d" % s] = u.copy()
        x = cartpole_step(x, u)
    return out


def golden_run_api(ns, system, task, mlp, K, H, seed, n_steps):
    """Same through the public ``MPPI.run`` (mppi.py:154-168) -- pins that the
    decomposition above is what ``run`` does."""
    np.random.seed(seed)
    with ref_loader.quiet():
        ctl = ns.MPPI(system, task, mlp, horizon=H, num_path=K, sigma=1.0, lmda=1.0)
    x = CART_X0.copy()
    constate = np.concatenate([x, np.zeros(1)])
    us = []
    for _ in range(n_steps):
        u, constate = ctl.run(constate, x)
        us.append(u.copy())
        x = cartpole_step(x, u)
    return np.array(us)


def golden_mlp(ns):
    out = {}
    sysc = ns.System(["a", "b", "c", "d"], ["u"])
    sysh = ns.System(["o%d" % i for i in range(17)], ["u%d" % i for i in range(6)])
    cases = [("relu", sysc, 2, 64), ("tanh", sysc, 1, 16), ("sigmoid", sysc, 3, 32), ("selu", sysc, 2, 24),
             ("relu", sysh, 3, 64), ("tanh", sysh, 4, 40)]
    out["n_cases"] = len(cases)
    for c, (act, system, nh, hid) in enumerate(cases):
        mlp = random_mlp(ns, system, nh, hid, act, seed=10 + c)
        rng = np.random.default_rng(1000 + c)
        m = 12
        X = rng.normal(size=(m, system.obs_dim)) * 1.5
        U = rng.normal(size=(m, system.ctrl_dim))
        pre = "c%d_" % c
        out.update(MLPParams.from_reference_mlp(mlp).to_npz_dict(pre))
        out[pre + "X"], out[pre + "U"] = X, U
        out[pre + "pred_batch"] = mlp.pred_batch(X, U)
        out[pre + "pred0"] = mlp.pred(X[0], U[0])
        xn, jx, ju = mlp.pred_diff_batch(X, U)
        out[pre + "diff_xn"], out[pre + "diff_jx"], out[pre + "diff_ju"] = xn, jx, ju
        xn1, jx1, ju1 = mlp.pred_diff(X[1], U[1])
        out[pre + "diff1_xn"], out[pre + "diff1_jx"], out[pre + "diff1_ju"] = xn1, jx1, ju1
    return out


def golden_ilqr(ns, system, task, mlp, H, x0s):
    out = {"H": H, "n": len(x0s)}
    for i, x0 in enumerate(x0s):
        with ref_loader.quiet():
            ctl = ns.IterativeLQR(system, task, mlp, horizon=H, verbose=True)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            conv, states, ctrls, Ks, ks = ctl.compute_ilqr_default(x0, np.zeros((H, 1)), silent=False)
        txt = buf.getvalue()
        alphas = [float(a) for a in re.findall(r"alpha is successful at ([0-9.eE+-]+) with", txt)]
        idx = [int(round(np.log(a) / np.log(0.2))) if a > 0 else -1 for a in alphas]
        n_iter = len(re.findall(r"At iteration \d+", txt))
        with ref_loader.quiet():
            ctl2 = ns.IterativeLQR(system, task, mlp, horizon=H)
            u, newstate = ctl2.run(np.concatenate([x0, np.zeros(1)]), x0)
        pre = "p%d_" % i
        out[pre + "x0"] = x0
        out[pre + "converged"] = bool(conv)
        out[pre + "states"], out[pre + "ctrls"], out[pre + "Ks"], out[pre + "ks"] = states, ctrls, Ks, ks
        out[pre + "alpha_idx"] = np.array(idx, dtype=np.int64)
        out[pre + "n_iter"] = n_iter
        out[pre + "ls_fail"] = "Line search fails" in txt
        out[pre + "run_u"] = u
    return out


def main():
    ns = ref_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)
    system, task = make_cartpole(ns)
    mlp = train_cartpole_mlp(ns, system)
    np.savez_compressed(os.path.join(GOLD, "cartpole_mlp.npz"),
                        **MLPParams.from_reference_mlp(mlp).to_npz_dict(),
                        Q=CART_Q, R=CART_R, F=CART_F, goal=np.zeros(4), umin=np.array([-20.0]),
                        umax=np.array([20.0]), x0=CART_X0, dt=0.05)
    # --- MPPI through the unmodified reference (ctrl_dim == 1) ---
    g = golden_mppi(ns, system, task, mlp, K=256, H=20, seed=0, n_steps=4)
    g["run_api_us"] = golden_run_api(ns, system, task, mlp, K=256, H=20, seed=0, n_steps=4)
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_K256_H20.npz"), **g)
    g = golden_mppi(ns, system, task, mlp, K=512, H=30, seed=3, n_steps=6, sigma=0.05, lmda=0.5,
                    x_init=[0.2, 0.4, -0.3, 0.1])
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_K512_H30.npz"), **g)
    g = golden_mppi(ns, system, task, mlp, K=4096, H=30, seed=1, n_steps=2)
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_K4096_H30.npz"), **g)
    g = golden_mppi(ns, system, task, mlp, K=100, H=5, seed=2, n_steps=3, sigma=0.37, lmda=0.4,
                    x_init=[0.3, -0.2, 0.1, 0.0])
    np.savez_compressed(os.path.join(GOLD, "mppi_cartpole_K100_H5.npz"), **g)
    # --- ctrl_dim > 1: record that the unmodified reference raises (mppi.py:139) ---
    sysh = ns.System(["o%d" % i for i in range(17)], ["u%d" % i for i in range(6)])
    sysh.dt = 0.05
    taskh = ns.Task(sysh)
    for n in sysh.controls:
        taskh.set_ctrl_bound(n, -1.0, 1.0)
    taskh.set_cost(ns.QuadCost(sysh, np.eye(17), 0.01 * np.eye(6), 10 * np.eye(17)))
    mh = random_mlp(ns, sysh, 3, 256, "relu", seed=100)
    raised = ""
    try:
        np.random.seed(0)
        with ref_loader.quiet():
            c = ns.MPPI(sysh, taskh, mh, horizon=5, num_path=64)
            c.run(np.zeros(23), np.zeros(17))
    except ValueError as e:
        raised = str(e)
    np.savez_compressed(os.path.join(GOLD, "mppi_nu6_reference_raises.npz"), message=raised)
    # --- MLP inference + Jacobians ---
    np.savez_compressed(os.path.join(GOLD, "mlp_cases.npz"), **golden_mlp(ns))
    # --- iLQR ---
    x0s = [CART_X0, np.array([0.3, -0.2, 0.1, 0.0]), np.array([1.5, 0.5, -0.5, 0.2])]
    np.savez_compressed(os.path.join(GOLD, "ilqr_cartpole_H50.npz"), **golden_ilqr(ns, system, task, mlp, 50, x0s))
    np.savez_compressed(os.path.join(GOLD, "ilqr_cartpole_H10.npz"), **golden_ilqr(ns, system, task, mlp, 10, x0s[:2]))
    print("golden fixtures written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  %-40s %8d B" % (f, os.path.getsize(os.path.join(GOLD, f))))


if __name__ == "__main__":
    sys.exit(main())

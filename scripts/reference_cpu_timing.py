"""BASELINE.md section 3, rows (i) / (ii) for the configurations the UNMODIFIED reference can run (ctrl_dim = 1):
its own ``MPPI.run`` (autompc/control/mppi.py:154-168, ``MLP.pred_batch`` on torch CPU float64, Python loop over the K
samples in ``cost_eqn``) timed in the BUILD CONTAINER (the reference tree does not exist on the GPU box), next to the
float64 NumPy port that ``bench.py`` times as ``cpu_baseline`` on the GPU box -- so that the port can be related to the
real thing on one host.  C1 = cartpole K=256 H=20, C2 = cartpole K=4096 H=30, the trained cartpole MLP 2x64.

C4 = cartpole IterativeLQR H=50 (``compute_ilqr_default``, autompc/control/ilqr.py:100-265) from the task's initial
observation, next to the float64 NumPy port ``bench.py --workload c4`` reports.

C5 = the tuning batch on a bounded sample.

    python scripts/reference_cpu_timing.py > profiles/r02c_reference_cpu.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from oracle import ref_loader
from oracle.make_golden import CART_X0, make_cartpole
from oracle.make_golden_f import reference_mlp_from_npz


def time_reference(ns, system, task, mlp, K, H, threads, n_warm, n_steps):
    torch.set_num_threads(threads)
    np.random.seed(0)
    with ref_loader.quiet():
        ctl = ns.MPPI(system, task, mlp, horizon=H, num_path=K, sigma=1.0, lmda=1.0)
        x = CART_X0.copy()
        constate = np.concatenate([x, np.zeros(1)])
        ts = []
        for i in range(n_warm + n_steps):
            t0 = time.perf_counter()
            u, constate = ctl.run(constate, x)
            ts.append(time.perf_counter() - t0)
    ts = np.array(ts[n_warm:])
    return {"ms_per_step_mean": 1e3 * float(ts.mean()), "ms_per_step_median": 1e3 * float(np.median(ts)),
            "steps_per_s": float(1.0 / ts.mean()), "steps_timed": int(n_steps), "warmup": int(n_warm),
            "torch_threads": int(threads)}


def main():
    ns = ref_loader.load()
    z = np.load(os.path.join(ROOT, "tests", "golden", "cartpole_mlp.npz"))
    system, task = make_cartpole(ns)
    mlp = reference_mlp_from_npz(ns, system, z)
    ncpu = os.cpu_count() or 1
    out = {"host": bench.cpu_info(), "torch": torch.__version__, "numpy": np.__version__,
           "what": "unmodified reference autompc.control.mppi.MPPI.run vs the float64 NumPy port (oracle/mppi_oracle.py), "
                   "same host, trained cartpole MLP[5-64-64-4], QuadCost of examples/3_Controllers_and_Tasks.ipynb cell 6",
           "configs": []}
    for name, K, H, n_warm, n_steps in (("C1", 256, 20, 5, 40), ("C2", 4096, 30, 1, 5)):
        rows = []
        for threads in (1, ncpu):
            r = time_reference(ns, system, task, mlp, K, H, threads, n_warm, n_steps)
            r.update(variant="unmodified reference MPPI.run", kind="reference")
            rows.append(r)
        torch.set_num_threads(ncpu)
        wl = bench.workload("c2")
        wl.update(weights=bench.trained_cartpole_weights(), K=K, H=H)
        for variant, threads, faithful in (("port, vectorised, all threads", None, False), ("port, vectorised, 1 thread", 1, False),
                                           ("port, reference-faithful Python K-loop, 1 thread", 1, True)):
            rate, cores, sample = bench.cpu_port_rate(wl, 6.0, threads=threads, faithful=faithful)
            rows.append({"variant": variant, "kind": "port", "steps_per_s": rate, "ms_per_step_mean": 1e3 / rate,
                         "cores": cores, "sample": sample})
        out["configs"].append({"config": "%s cartpole MPPI K=%d H=%d" % (name, K, H), "rows": rows})
    # --- C4: the unmodified reference's iLQR solve vs the port
    wl = bench.workload("c4")
    rows = []
    for threads in (1, ncpu):
        torch.set_num_threads(threads)
        with ref_loader.quiet():
            ctl = ns.IterativeLQR(system, task, mlp, horizon=wl["H"])
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                conv, states, ctrls, Ks, ks = ctl.compute_ilqr_default(CART_X0.copy(), np.zeros((wl["H"], 1)), silent=True)
                ts.append(time.perf_counter() - t0)
        rows.append({"variant": "unmodified reference compute_ilqr_default", "kind": "reference", "torch_threads": threads,
                     "ms_per_step_mean": 1e3 * float(np.mean(ts[1:])), "steps_per_s": float(1.0 / np.mean(ts[1:])),
                     "steps_timed": 2, "warmup": 1, "converged": bool(conv)})
    torch.set_num_threads(ncpu)
    rate, cores, sample = bench.cpu_ilqr_rate(wl, 6.0)
    rows.append({"variant": "port, 1 thread", "kind": "port", "steps_per_s": rate, "ms_per_step_mean": 1e3 / rate,
                 "cores": cores, "sample": sample})
    out["configs"].append({"config": "C4 cartpole IterativeLQR H=50 (one solve = one step)", "rows": rows})
    # --- C5: the tuner's inner loop (tuning/pipeline_tuner.py:213-239) on a bounded sample: the reference's own
    # simulate() driving its own MPPI for the first candidates, 5 closed-loop steps each; scaled to the whole job by
    # the share of sample-steps done (the same scaling as bench.cpu_c5_rate), next to that port
    import importlib
    with ref_loader.quiet():
        simulate = importlib.import_module("autompc.utils.simulation").simulate
    wl = bench.workload("c5")
    cands, _ = bench.c5_candidates(wl)
    work_total = float(sum(kw["num_path"] * kw["horizon"] for kw, _ in cands)) * wl["T"]
    torch.set_num_threads(1)
    work_done, steps, t0 = 0.0, 0, time.perf_counter()
    for kw, etask in cands:
        Q, R, F = etask.get_cost().get_cost_matrices()
        rtask = ns.Task(system)
        rtask.set_ctrl_bound("u", -20.0, 20.0)
        rtask.set_cost(ns.QuadCost(system, Q, R, F, goal=np.zeros(4)))
        np.random.seed(kw["seed"])
        with ref_loader.quiet():
            ctl = ns.MPPI(system, rtask, mlp, horizon=kw["horizon"], num_path=kw["num_path"], sigma=kw["sigma"],
                          lmda=kw["lmda"])
            simulate(ctl, CART_X0.copy(), sim_model=mlp, max_steps=5, silent=True)
        steps += 5
        work_done += 5.0 * kw["num_path"] * kw["horizon"]
        if time.perf_counter() - t0 > 20.0:
            break
    dt = time.perf_counter() - t0
    rows = [{"variant": "unmodified reference simulate() + MPPI.run", "kind": "reference", "torch_threads": 1,
             "steps_per_s": (wl["n_cand"] * wl["T"]) / (dt * work_total / work_done),
             "ms_per_step_mean": 1e3 * (dt * work_total / work_done) / (wl["n_cand"] * wl["T"]),
             "sample": "%d closed-loop steps (5 per candidate) in %.1f s = %.4g of the job's sample-steps"
                       % (steps, dt, work_done / work_total)}]
    torch.set_num_threads(ncpu)
    rate, cores, sample = bench.cpu_c5_rate(wl, 10.0)
    rows.append({"variant": "port, all threads", "kind": "port", "steps_per_s": rate, "ms_per_step_mean": 1e3 / rate,
                 "cores": cores, "sample": sample})
    out["configs"].append({"config": "C5 64 candidates x 200 closed-loop steps (MPC steps/s of the whole job)", "rows": rows})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

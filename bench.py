#!/usr/bin/env python
"""bench.py -- MPC steps/sec of the MPPI solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|c4|c5]
                    [--scaling strong|weak] [--precision auto|fp32|fp16|bf16]

One "step" = one MPPI solve (``MPPI.run`` of autompc/control/mppi.py:154-168): shift, K rollouts of
H steps through the MLP dynamics with QuadCost, exponentiated-cost-weighted update.  Workload at
N=1 (and sharded over samples at N>1): BASELINE.json config C3 = HalfCheetah-dim (nx=17, nu=6),
MLP[23-256-256-256-17] ReLU, K=16384, H=50 (SURVEY.md 8d recipe, synthetic weights).  ``--scaling strong``
(default, the BASELINE metric: K=16384 in total) or ``weak`` (K=16384 PER GPU).  Other workloads
(``--workload``): c2 = cartpole MPPI K=4096 H=30; c4 = cartpole IterativeLQR H=50 (replicas at N>1);
c5 = 64 candidate controllers x 200 closed-loop steps, device resident (candidates dealt over the ranks).

Prints ONE JSON line (rank 0).  ``value`` = solves/s with the observation already in HBM and
noise generated in-kernel (CUDA events, per-step, L2 flushed between steps, max over ranks);
``e2e`` = the same through the public ``Controller.run(state, new_obs)`` with host NumPy buffers
(H2D of the observation + D2H of the control inside the timed region);
``roofline`` = algorithmic MLP FLOPs per solve / kernel time against the measured bf16 peak;
``cpu_baseline`` = the float64 NumPy oracle port timed on this box's host cores (bounded sample).
``--impl reference`` times only that CPU arm (the reference is pure Python; for ctrl_dim > 1 the
unmodified reference raises, so the arm is the documented restatement, oracle/mppi_oracle.py).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MPC steps/sec (MPPI K=16384 H=50, MLP dyn)"
UNIT = "steps/s"
METRICS = {"c3": (METRIC, UNIT), "c2": ("MPC steps/sec (MPPI K=4096 H=30, cartpole MLP dyn)", UNIT),
           "c4": ("iLQR solves/sec (cartpole MLP dyn, H=50, 50 iterations)", "solves/s"),
           "c5": ("closed-loop MPC steps/sec (64 candidate MPPI controllers x 200 steps, surrogate sim)", UNIT)}


def trained_cartpole_weights():
    """The cartpole MLP 2x64 trained by the reference's own MLP.train (oracle/make_golden.py), frozen as a fixture."""
    from autompc_b200.mlp import MLPWeights
    return MLPWeights.from_npz(np.load(os.path.join(ROOT, "tests", "golden", "cartpole_mlp.npz")))


def workload(name, world=1, scaling="strong"):
    from autompc_b200.problems import cartpole_problem, halfcheetah_dim_problem
    if name == "c3":
        system, task, w, x0 = halfcheetah_dim_problem()
        K = 16384 * (world if scaling == "weak" else 1)
        return dict(name=name, system=system, task=task, weights=w, x0=x0, K=K, H=50, sigma=1.0, lmda=1.0,
                    label="C3 HalfCheetah-dim nx=17 nu=6 MLP[23-256-256-256-17] relu, MPPI K=%d H=50, "
                          "QuadCost Q=I R=0.01I F=10I, sigma=1 lmda=1" % K)
    if name == "c2":
        system, task, w, x0 = cartpole_problem()
        K = 4096 * (world if scaling == "weak" else 1)
        return dict(name=name, system=system, task=task, weights=w, x0=x0, K=K, H=30, sigma=1.0, lmda=1.0,
                    label="C2 cartpole nx=4 nu=1 MLP[5-64-64-4] relu, MPPI K=%d H=30" % K)
    if name == "c4":
        system, task, w, x0 = cartpole_problem(trained_cartpole_weights())
        return dict(name=name, system=system, task=task, weights=w, x0=x0, H=50, K=0,
                    label="C4 cartpole nx=4 nu=1 trained MLP[5-64-64-4] relu, IterativeLQR H=50, max_iter=50, 10-alpha "
                          "line search, re-solved from zeros every step (ilqr.py:281-288)")
    if name == "c5":
        system, task, w, x0 = cartpole_problem(trained_cartpole_weights())
        return dict(name=name, system=system, task=task, weights=w, x0=x0, H=0, K=0, n_cand=64, T=200,
                    label="C5 64 candidate MPPI controllers (horizon 5-30, num_path 100-1000, sigma 1e-4-2, lmda 0.1-2, "
                          "QuadCost Q/F/R log-uniform 1e-3..1e4: mppi.py:52-63, quad_cost_factory.py:46-58) x 200 "
                          "closed-loop steps on the trained cartpole MLP as model and surrogate, scored with the "
                          "benchmark's ThresholdCost (benchmarks/cartpole.py:38-60)")
    raise SystemExit("unknown workload %s" % name)


def c5_candidates(wl):
    """64 configurations from the reference's own hyper-parameter ranges, np.random.default_rng(100)."""
    from autompc_b200.plugin import QuadCost, Task, ThresholdCost
    system = wl["system"]
    rng = np.random.default_rng(100)
    out = []
    for i in range(wl["n_cand"]):
        kw = dict(horizon=int(rng.integers(5, 31)), num_path=int(rng.integers(100, 1001)),
                  sigma=float(rng.uniform(1e-4, 2.0)), lmda=float(rng.uniform(0.1, 2.0)), seed=i)
        g = np.exp(rng.uniform(np.log(1e-3), np.log(1e4), size=9))
        task = Task(system)
        task.set_ctrl_bound("u", -20.0, 20.0)
        task.set_cost(QuadCost(system, np.diag(g[:4]), np.diag(g[8:9]), np.diag(g[4:8]), goal=np.zeros(4)))
        out.append((kw, task))
    score = ThresholdCost(system, np.zeros(4), [0, 3], 0.2)      # benchmarks/cartpole.py:51: goal 0, obs_range (0,3), 0.2
    return out, score


def mlp_flops_per_solve(w, K, H):
    """F = 2 * K * H * sum_layers(in * out), un-padded, MLP only."""
    return 2.0 * K * H * sum(int(a) * int(b) for a, b in zip(w.dims[:-1], w.dims[1:]))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(bf16_burst=float(d["bf16_tflops"]), bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), source="MEASURED_PEAKS.json (of measured)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback (of fallback)")


def tc_build_note(ctl):
    """Which build of the tensor-core kernel the handle runs (debug tap of the C ABI); '' for the fp32 kernel."""
    try:
        from autompc_b200 import _abi
        mode = int(_abi.lib().ampc_mppi_debug_tc_mode(ctl._h))
    except Exception:
        return ""
    if not mode:
        return ""
    return ", cta_group::%d%s" % (mode & 3, ", dz build (input layer fed by the output accumulator through kind::tf32)"
                                  if mode & 16 else "")


def ncu_traffic(workload_name, precision, world):
    """(bytes, source): dram__bytes_read.sum + dram__bytes_write.sum per launch of the rollout kernel.  NOT measured by
    this run (a number taken under a profiler never is a bench value): it is read from profiles/traffic.json, which
    records the committed `ncu --set full` capture of this same command it came from; (None, None) when no capture
    exists for the configuration."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            for e in json.load(f)["captures"]:
                if e["workload"] == workload_name and precision in e["precisions"] and e["n_gpus"] == world:
                    return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"]), e["source"]
    except (OSError, KeyError, ValueError):
        pass
    return None, None


# --------------------------------------------------------------------------- CPU arm ---
def _oracle_side(wl):
    from oracle.mppi_oracle import MLPParams, QuadCostParams
    w, task = wl["weights"], wl["task"]
    Q, R, F = task.get_cost().get_cost_matrices()
    p = MLPParams(w.W, w.b, w.act, w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, w.nx, w.nu)
    b = task.get_ctrl_bounds()
    return p, QuadCostParams(Q, R, F, task.get_cost().get_goal()), b[:, 0], b[:, 1]


def _set_threads(n):
    try:
        import threadpoolctl
        return threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        return None


def _threads_now():
    try:
        import threadpoolctl
        return max([i.get("num_threads", 1) for i in threadpoolctl.threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_port_rate(wl, budget_s, min_solves=2, threads=None, faithful=False):
    """Times the float64 NumPy oracle (restatement of mppi.py:110-168) on the host.  ``faithful``: the stage cost is
    evaluated with the reference's Python loop over samples (mppi.py:73-78) instead of the vectorised form.
    The sample is a slice of the K samples at the full horizon; cost is linear in K, so the rate of a
    full-K step is sample_rate * K_sample / K.  Returns (steps_per_s, cores, sample_description)."""
    from oracle.mppi_oracle import MPPIOracle
    p, cost, umin, umax = _oracle_side(wl)
    K, H = wl["K"], wl["H"]

    def run(ks, n):
        np.random.seed(0)
        o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=ks, sigma=wl["sigma"], lmda=wl["lmda"],
                       faithful_loop=faithful)
        o.solve(wl["x0"])                                   # warm-up (BLAS threads, allocations)
        t0 = time.perf_counter()
        for _ in range(n):
            o.solve(wl["x0"])
        return (time.perf_counter() - t0) / n

    # all host threads by default, also under torchrun (which exports OMP_NUM_THREADS=1 for every rank)
    lim = _set_threads(threads or (os.cpu_count() or 1))
    try:
        probe_k = min(K, 64 if faithful else 512)
        t_probe = run(probe_k, 1)
        per_sample = t_probe / probe_k
        ks = int(min(K, max(probe_k, (budget_s / (min_solves + 1)) / per_sample)))
        ks = max(64, (ks // 64) * 64) if ks < K else K
        n = max(min_solves, int(budget_s / max(per_sample * ks, 1e-9)) - 1)
        n = min(n, 20)
        t = run(ks, n)
        cores = _threads_now()
    finally:
        if lim is not None and hasattr(lim, "restore_original_limits"):
            lim.restore_original_limits()
    rate = (1.0 / t) * (ks / K)
    sample = ("float64 NumPy port of MPPI.run (%s), %d of %d samples x full H=%d, %d solves, "
              "%.2f s/solve-sample; steps/s scaled by %d/%d"
              % ("Python loop over samples for the stage cost as mppi.py:73-78" if faithful else "vectorised cost loop",
                 ks, K, H, n, t, ks, K))
    return rate, int(cores), sample


def cpu_rows(wl, budget_s):
    """The rows BASELINE.md section 3 asks for, each on its own bounded sample: the vectorised port at all threads
    (= cpu_baseline.value, the most favourable CPU number) and at 1 thread, and the reference-faithful evaluation
    (Python loop over the K samples, mppi.py:73-78) at 1 thread."""
    rows = []
    for variant, threads, faithful, share in (("vectorised, all threads", None, False, 0.5), ("vectorised, 1 thread", 1, False, 0.25),
                                              ("reference-faithful Python K-loop, 1 thread", 1, True, 0.25)):
        rate, cores, sample = cpu_port_rate(wl, budget_s * share, threads=threads, faithful=faithful)
        rows.append({"variant": variant, "value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample})
    return rows


def cpu_ilqr_rate(wl, budget_s):
    from oracle.ilqr_oracle import ilqr_solve
    p, cost, umin, umax = _oracle_side(wl)
    lim = _set_threads(1)                                  # 5x5 matrices and batch-10 MLPs: threads only add overhead
    try:
        t0 = time.perf_counter()
        n = 0
        while True:
            r = ilqr_solve(p, cost, wl["system"].dt, wl["x0"], wl["H"], (umin, umax))
            n += 1
            if time.perf_counter() - t0 > budget_s or n >= 10:
                break
        t = (time.perf_counter() - t0) / n
    finally:
        if lim is not None and hasattr(lim, "restore_original_limits"):
            lim.restore_original_limits()
    return 1.0 / t, 1, ("float64 NumPy port of compute_ilqr_default (ilqr.py:100-265), %d full solves (%d iterations "
                        "each), %.2f s/solve" % (n, r["n_iter"], t))


def cpu_c5_rate(wl, budget_s):
    """Oracle closed loops (MPPI.run + MLP.pred per step) for the first candidates, a few steps each; MPC steps/s of
    the whole job = steps done / time (every candidate-step costs ~ num_path * horizon)."""
    from oracle.mppi_oracle import MLPParams, MPPIOracle, QuadCostParams, mlp_pred
    w = wl["weights"]
    p = MLPParams(w.W, w.b, w.act, w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, w.nx, w.nu)
    cands, _ = c5_candidates(wl)
    lim = _set_threads(os.cpu_count() or 1)
    work_done, work_total = 0.0, float(sum(kw["num_path"] * kw["horizon"] for kw, _ in cands)) * wl["T"]
    t0 = time.perf_counter()
    steps = 0
    try:
        for kw, task in cands:
            Q, R, F = task.get_cost().get_cost_matrices()
            np.random.seed(kw["seed"])
            o = MPPIOracle(p, QuadCostParams(Q, R, F, np.zeros(4)), [-20.0], [20.0], horizon=kw["horizon"],
                           num_path=kw["num_path"], sigma=kw["sigma"], lmda=kw["lmda"])
            x = wl["x0"].copy()
            for _ in range(5):
                u = o.solve(x)
                x = mlp_pred(p, x, u)
                steps += 1
                work_done += kw["num_path"] * kw["horizon"]
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        cores = _threads_now()
    finally:
        if lim is not None and hasattr(lim, "restore_original_limits"):
            lim.restore_original_limits()
    rate = (wl["n_cand"] * wl["T"]) / (dt * work_total / work_done)
    return rate, int(cores), ("float64 NumPy port: %d closed-loop steps (5 per candidate) in %.1f s = %.4g of the job's "
                              "sample-steps; steps/s of the whole job scaled by that share" % (steps, dt, work_done / work_total))


def cpu_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return "%s, %d logical cores" % (model, os.cpu_count() or 1)


def cpu_arm(wl, budget_s):
    if wl["name"] in ("c3", "c2"):
        return cpu_port_rate(wl, budget_s)
    if wl["name"] == "c4":
        return cpu_ilqr_rate(wl, budget_s)
    return cpu_c5_rate(wl, budget_s)


def reference_arm(args, wl, rank):
    if rank != 0:
        return
    metric, unit = METRICS[wl["name"]]
    budget = float(os.environ.get("AMPC_REF_BUDGET_S", "60"))
    rate, cores, sample = cpu_arm(wl, budget)
    line = {
        "impl": "reference", "metric": metric, "value": rate, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "host": cpu_info(),
                   "note": "reference is pure Python/NumPy/torch-CPU; for ctrl_dim=6 the unmodified reference raises "
                           "(mppi.py:139), so this arm is the documented restatement (oracle/mppi_oracle.py)"},
        "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ clock sampling ---
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU with NVML while the timed region runs."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
           "hw_power_brake_slowdown": 0x80}
    NOTE = {"sw_power_cap": 0x4, "gpu_idle": 0x1, "applications_clocks_setting": 0x2, "sync_boost": 0x10}

    def __init__(self, index, period_s=0.002):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in list(self.BAD.items()) + list(self.NOTE.items()):
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ GPU arms ---
class Dist:
    """torch.distributed plumbing shared by the arms (one process per GPU, NCCL)."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch, self.rank, self.world, self.local_rank = torch, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.group = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.group = dist.group.WORLD

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxed, summed):
        """max / sum over ranks of two float lists."""
        if self.world == 1:
            return list(maxed), list(summed)
        t = self.torch
        a = t.tensor(maxed, dtype=t.float64, device=self.dev)
        b = t.tensor(summed, dtype=t.float64, device=self.dev)
        self.dist.all_reduce(a, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM)
        return a.tolist(), b.tolist()

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def gpu_arm_mppi(args, wl, d):
    import torch
    from autompc_b200 import MPPI, B200MLP, _abi
    rank, world, local_rank, dev = d.rank, d.world, d.local_rank, d.dev
    system, task, w = wl["system"], wl["task"], wl["weights"]
    model = B200MLP(system, w, device=local_rank)
    np.random.seed(0)
    ctl = MPPI(system, task, model, horizon=wl["H"], num_path=wl["K"], sigma=wl["sigma"], lmda=wl["lmda"],
               seed=0, noise="philox", precision=args.precision, device=local_rank, group=d.group,
               exchange=args.exchange)
    nx, nu = w.nx, w.nu
    x0_dev = torch.tensor(wl["x0"], dtype=torch.float32, device=dev)
    u_dev = torch.zeros(nu, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    stream = torch.cuda.current_stream()

    def one_step():
        ctl.solve_device(x0_dev, u_dev, stream=stream.cuda_stream)

    # ---- multi-GPU parity, outside the timed region: the sharded solve against ONE GPU doing all K samples from the
    #      same action sequence, seed and counter (Philox is keyed by the global sample index)
    parity = None
    if world > 1:
        act0 = ctl.act_sequence
        one_step()
        torch.cuda.synchronize()
        u_sharded = u_dev.cpu().numpy().astype(np.float64)
        act_sharded = ctl.act_sequence
        np.random.seed(0)
        single = MPPI(system, task, model, horizon=wl["H"], num_path=wl["K"], sigma=wl["sigma"], lmda=wl["lmda"],
                      seed=0, noise="philox", precision=ctl.precision, device=local_rank)
        single.act_sequence = act0
        u_single = single.solve(wl["x0"]) / ctl.ctrl_scale
        du = float(np.abs(u_sharded / ctl.ctrl_scale - u_single).max())
        da = float(np.abs(act_sharded - single.act_sequence).max())
        single.close()
        (du, da), _ = d.reduce([du, da], [0.0])
        parity = {"max_abs_du_vs_single_gpu": du, "max_abs_dact_vs_single_gpu": da, "atol": 2e-5,
                  "what": "first sharded solve vs one GPU rolling out all K samples from the same action sequence, seed "
                          "and counter (normalised controls); the only difference is the fp32 summation order of the "
                          "softmax merge; max over ranks"}
        assert du < 2e-5 and da < 2e-5, "sharded solve differs from the single-GPU solve: %g / %g" % (du, da)

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        one_step()
    d.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    n0 = _abi.launch_count()
    with ClockSampler(local_rank) as clk:
        d.barrier()
        for i in range(args.steps):
            flush.zero_()
            starts[i].record(stream)
            one_step()
            ends[i].record(stream)
        d.barrier()
    launches = _abi.launch_count() - n0
    per_step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    total_ms = float(per_step_ms.sum())
    u_host = u_dev.cpu().numpy()
    assert np.all(np.isfinite(u_host)), "non-finite control from the solve"

    # ---- end to end through Controller.run with host buffers (H2D of obs, D2H of u inside)
    rng = np.random.default_rng(1)
    obs = [wl["x0"] + 0.01 * rng.normal(size=nx) for _ in range(args.steps + 3)]
    constate = np.concatenate([wl["x0"], np.zeros(nu)])
    for i in range(3):
        u, constate = ctl.run(constate, obs[i])
    d.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        u, constate = ctl.run(constate, obs[3 + i])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    (total_ms, e2e_s), (launches,) = d.reduce([total_ms, e2e_s], [float(launches)])
    # ---- N=1: the other tensor-core operand format beside the default, same workload (extra key, outside the timed region)
    alt = None
    if world == 1 and ctl.precision in ("fp16", "bf16") and not args.no_weak_probe:
        other = "bf16" if ctl.precision == "fp16" else "fp16"
        try:
            np.random.seed(0)
            ctl_o = MPPI(system, task, model, horizon=wl["H"], num_path=wl["K"], sigma=wl["sigma"], lmda=wl["lmda"], seed=0,
                         noise="philox", precision=other, device=local_rank)
            n_o = min(args.steps, 50)
            for _ in range(5):
                flush.zero_()
                ctl_o.solve_device(x0_dev, u_dev, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_o)]
            for a, b in ev:
                flush.zero_()
                a.record(stream)
                ctl_o.solve_device(x0_dev, u_dev, stream=stream.cuda_stream)
                b.record(stream)
            torch.cuda.synchronize()
            ms_o = float(sum(a.elapsed_time(b) for a, b in ev)) / n_o
            ctl_o.close()
            alt = {other: {"ms_per_step": ms_o, "value": 1e3 / ms_o, "steps": n_o,
                           "deviation_from_float64_oracle": ({"fp16": 2.3e-4, "bf16": 3.4e-3} if "dz build" in tc_build_note(ctl)
                                                             else {"fp16": 2.1e-4, "bf16": 2.1e-3})[other],
                           "what": "same workload and timing rule with %s operands; the deviation is the measured max |act - "
                                   "oracle| at this size (profiles/r02_precision_dz.jsonl with the dz build of the kernel, "
                                   "r02_precision.jsonl without)" % other}}
        except ValueError:
            alt = None
    # ---- N>1, default (strong) scaling: the same machine on the weak-scaled problem (K per GPU fixed), as an extra key;
    #      its launches are outside the timed region above and are not counted in gpu_launches
    weak = None
    if world > 1 and args.scaling == "strong" and not args.no_weak_probe:
        Kw = wl["K"] * world
        np.random.seed(0)
        ctl_w = MPPI(system, task, model, horizon=wl["H"], num_path=Kw, sigma=wl["sigma"], lmda=wl["lmda"], seed=0,
                     noise="philox", precision=ctl.precision, device=local_rank, group=d.group, exchange=ctl.exchange)
        n_w = min(args.steps, 50)
        for _ in range(5):
            flush.zero_()
            ctl_w.solve_device(x0_dev, u_dev, stream=stream.cuda_stream)
        d.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_w)]
        for a, b in ev:
            flush.zero_()
            a.record(stream)
            ctl_w.solve_device(x0_dev, u_dev, stream=stream.cuda_stream)
            b.record(stream)
        d.barrier()
        ms_w = float(sum(a.elapsed_time(b) for a, b in ev))
        (ms_w,), _ = d.reduce([ms_w], [0.0])
        ctl_w.close()
        weak = {"K_total": Kw, "K_per_gpu": wl["K"], "steps": n_w, "ms_per_step": ms_w / n_w,
                "base_problem_solves_per_s": world * n_w / (ms_w * 1e-3),
                "what": "the same run on the weak-scaled problem (K=%d per GPU, one solve of K=%d): solves of the base "
                        "problem per second = N x solves/s; `python bench.py --scaling weak` reports it as the line's value"
                        % (wl["K"], Kw)}
    if rank == 0:
        metric, unit = METRICS[wl["name"]]
        peaks = measured_peaks()
        ms_per_step = total_ms / args.steps
        value = args.steps / (total_ms * 1e-3)
        value_note = "solves per second"
        if args.scaling == "weak" and world > 1:
            # weak scaling: every solve rolls out N x the base problem's samples; the whole-job aggregate is counted in
            # units of the base problem (K per GPU), so that it is comparable across N
            value *= world
            value_note = ("weak scaling: %d x solves per second = solves of the base problem (K=%d) per second; one solve of "
                          "K=%d samples takes ms_per_step" % (world, wl["K"] // world, wl["K"]))
        flops = mlp_flops_per_solve(w, wl["K"], wl["H"])
        # captures are per kernel build: "<precision>" = the dz build, "<precision>-nodz" = the build without it
        traffic, traffic_src = ncu_traffic(wl["name"], ctl.precision + ("" if "dz build" in tc_build_note(ctl) or ctl.precision == "fp32"
                                                                        else "-nodz"), world)
        fused = world > 1 and ctl.exchange == "nvlink"
        # dominant kernel = the rollout kernel; at N=1 it is the whole step.  Its average duration over
        # the timed region is ms_per_step minus the (tiny) merge kernel at N>1, which we do not subtract.
        achieved = flops / world / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "fp16"}.get(ctl.precision, "f32"), "data": "synthetic",
            "config": {"workload": wl["label"], "value_is": value_note, "noise": "in-kernel Philox4x32-10",
                       "precision": ctl.precision,
                       "precision_note": {"fp16": "tcgen05 kind::f16 with IEEE-half operands (11-bit significands = the "
                                                  "operand precision of tf32), fp32 accumulate/state/cost; deviation of "
                                                  "the updated action sequence from the float64 oracle at this size: "
                                                  "2.3e-4 with the dz build, 2.1e-4 without (profiles/r02_precision_dz.jsonl, "
                                                  "r02_precision.jsonl)",
                                          "bf16": "tcgen05 kind::f16 with bf16 operands; deviation from the float64 "
                                                  "oracle at this size: 3.4e-3 with the dz build, 2.1e-3 without "
                                                  "(profiles/r02_precision_dz.jsonl, r02_precision.jsonl)",
                                          "fp32": "CUDA-core fp32 FMA; deviation from the float64 oracle: 1.3e-5"}[ctl.precision],
                       "parallelism": ("%d samples sharded over %d GPU(s) (%s scaling), %d-float softmax record exchanged by %s"
                                       % (wl["K"], world, args.scaling, 2 + wl["H"] * nu,
                                          "NVLink peer stores + flags inside the rollout kernel (one launch per solve)"
                                          if ctl.exchange == "nvlink" else "one NCCL all-gather + merge kernel"))
                       if world > 1 else "1 GPU, one kernel per solve",
                       "timing": "CUDA events per step on the launch stream, sum over steps, max over ranks",
                       "l2": "256 MiB memset between timed steps (L2 flushed)",
                       "ms_per_step_min_median_max": [float(per_step_ms.min()), float(np.median(per_step_ms)),
                                                      float(per_step_ms.max())]},
            "clocks": clk.summary(),
            "e2e": {"value": args.steps / e2e_s * (world if (args.scaling == "weak" and world > 1) else 1), "unit": unit,
                    "h2d_bytes_per_step": 4 * nx,
                    "d2h_bytes_per_step": 4 * nu, "api": "autompc_b200.MPPI.run(state, new_obs) with NumPy float64 buffers",
                    "transfer": "host observation -> pinned float32 -> kernel parameters (H2D with the launch); control "
                                "written by the kernel's last CTA into mapped pinned host memory (D2H) with a sequence "
                                "number behind it that the host spins on (no stream synchronise)"
                                if (world == 1 or fused) else "pinned H2D / D2H copies on the stream"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_burst"], "traffic": traffic, "traffic_source": traffic_src,
                         "flop_per_launch": flops / world, "peak_source": "bf16 dense burst (kind::f16 runs fp16 and bf16 "
                         "operands at the same rate), " + peaks["source"],
                         "frac_of_sustained": achieved / peaks["bf16_sustained"],
                         "kernel": "mppi_rollout (%s)%s" % (ctl.precision, tc_build_note(ctl))},
        }
        if parity is not None:
            line["parity"] = parity
        if weak is not None:
            line["weak_scaling"] = weak
        if alt is not None:
            line["alt_precisions"] = alt
        if world == 1 and not args.no_cpu:
            budget = float(os.environ.get("AMPC_CPU_BUDGET_S", "24"))
            rows = cpu_rows(wl, budget)
            line["cpu_baseline"] = {"value": rows[0]["value"], "unit": unit, "cores": rows[0]["cores"], "kind": "port",
                                    "sample": rows[0]["sample"], "host": cpu_info(), "rows": rows}
        print(json.dumps(line), flush=True)
    ctl.close()


def gpu_arm_ilqr(args, wl, d):
    """C4: one step = one full IterativeLQR solve from x0 (one launch).  Does not shard: N replicas."""
    import torch
    from autompc_b200 import IterativeLQR, B200MLP, _abi
    system, task, w, x0 = wl["system"], wl["task"], wl["weights"], wl["x0"]
    il = IterativeLQR(system, task, B200MLP(system, w, device=d.local_rank), horizon=wl["H"], device=d.local_rank)
    conv, *_ = il.compute_ilqr(x0)
    info = dict(il.last_info)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=d.dev)
    for _ in range(max(args.warmup, 3)):
        il.launch_device(stream.cuda_stream)
    d.barrier()
    steps = args.steps
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    n0 = _abi.launch_count()
    with ClockSampler(d.local_rank) as clk:
        d.barrier()
        for i in range(steps):
            flush.zero_()
            starts[i].record(stream)
            il.launch_device(stream.cuda_stream)
            ends[i].record(stream)
        d.barrier()
    launches = _abi.launch_count() - n0
    per = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    total_ms = float(per.sum())
    # e2e: IterativeLQR.run (ilqr.py:267-295) with host buffers; reuse_feedback=0 => a full solve every call
    constate = np.concatenate([x0, np.zeros(w.nu)])
    for _ in range(3):
        u, constate = il.run(constate, x0)
    d.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        u, constate = il.run(constate, x0)
    e2e_s = time.perf_counter() - t0
    (total_ms, e2e_s), (launches,) = d.reduce([total_ms, e2e_s], [float(launches)])
    if d.rank == 0:
        metric, unit = METRICS["c4"]
        peaks = measured_peaks()
        ms = total_ms / steps
        n_it = info["n_iter"]
        flops = 2.0 * sum(int(a) * int(b) for a, b in zip(w.dims[:-1], w.dims[1:])) * (wl["H"] * (1 + 10 * n_it))
        line = {"metric": metric, "value": d.world * steps / (total_ms * 1e-3), "unit": unit, "n_gpus": d.world, "steps": steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["label"], "iterations": n_it, "converged": bool(conv),
                           "alpha_idx": info["alpha_idx"][:8],
                           "parallelism": "replicas only: the solve is one sequential problem (%d independent solves)" % d.world,
                           "timing": "CUDA events per solve on the launch stream", "l2": "256 MiB memset between timed solves",
                           "ms_per_step_min_median_max": [float(per.min()), float(np.median(per)), float(per.max())]},
                "clocks": clk.summary(),
                "e2e": {"value": d.world * steps / e2e_s, "unit": unit, "h2d_bytes_per_step": 8 * w.nx,
                        "d2h_bytes_per_step": 8 * ((wl["H"] + 1) * w.nx + wl["H"] * (w.nu + w.nu * w.nx + w.nu)) + 4 * (3 + 50),
                        "api": "autompc_b200.IterativeLQR.run(state, new_obs)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_burst"],
                             "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_burst"], "traffic": None,
                             "note": "latency-bound float64 chain of %d sequential stages on one CTA: neither roofline is "
                                     "approachable (SURVEY.md 8d); the figure is the line search's MLP flops against the "
                                     "bf16 peak, for completeness" % (n_it * 2 * wl["H"]), "kernel": "ilqr_kernel (f64)"}}
        if d.world == 1 and not args.no_cpu:
            rate, cores, sample = cpu_ilqr_rate(wl, float(os.environ.get("AMPC_CPU_BUDGET_S", "20")))
            line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                                    "host": cpu_info()}
        print(json.dumps(line), flush=True)
    il.close()


def gpu_arm_c5(args, wl, d):
    """C5: one step = the whole candidate batch (64 closed loops x 200 MPC steps), candidates dealt over the ranks."""
    import torch
    from autompc_b200 import MPPI, B200MLP, _abi, evaluate_candidates
    system, w, x0 = wl["system"], wl["weights"], wl["x0"]
    model = B200MLP(system, w, device=d.local_rank)
    cands, score = c5_candidates(wl)
    np.random.seed(1)
    ctls = [MPPI(system, task, model, device=d.local_rank, precision=args.precision, **kw) for kw, task in cands]
    T = wl["T"]
    for _ in range(max(1, min(args.warmup, 2))):
        evaluate_candidates(ctls, x0, T, model, group=d.group, cost=score)
    steps = max(1, min(args.steps, 20))
    d.barrier()
    n0 = _abi.launch_count()
    times = []
    with ClockSampler(d.local_rank) as clk:
        for _ in range(steps):
            d.barrier()
            t0 = time.perf_counter()
            costs, _ = evaluate_candidates(ctls, x0, T, model, group=d.group, cost=score)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    launches = _abi.launch_count() - n0
    total = float(np.sum(times))
    (total,), (launches,) = d.reduce([total], [float(launches)])
    if d.rank == 0:
        metric, unit = METRICS["c5"]
        peaks = measured_peaks()
        n_steps = wl["n_cand"] * T
        sample_steps = float(sum(kw["num_path"] * kw["horizon"] for kw, _ in cands)) * T
        flops = 2.0 * sum(int(a) * int(b) for a, b in zip(w.dims[:-1], w.dims[1:])) * sample_steps
        ms = 1e3 * total / steps
        line = {"metric": metric, "value": n_steps * steps / total, "unit": unit, "n_gpus": d.world, "steps": steps,
                "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": ctls[0].precision, "data": "synthetic",
                "config": {"workload": wl["label"], "candidate_evals_per_s": wl["n_cand"] * steps / total,
                           "parallelism": "candidates dealt round-robin over %d rank(s), no data-path collective; one "
                                          "stream per closed loop, all in flight" % d.world,
                           "timing": "host clock around start..finish of all closed loops, device synchronised on both "
                                     "sides (one stream per candidate: no single stream sees the work), max over ranks",
                           "l2": "inputs are generated on the device step by step; working set per candidate < L2",
                           "score": "ThresholdCost of the cartpole benchmark; finite for %d of %d candidates, mean %.1f"
                                    % (int(np.isfinite(costs).sum()), len(costs), float(np.nanmean(costs)))},
                "clocks": clk.summary(),
                "e2e": {"value": n_steps * steps / total, "unit": unit, "h2d_bytes_per_step": 8 * w.nx * wl["n_cand"],
                        "d2h_bytes_per_step": 8 * wl["n_cand"] * ((T + 1) * w.nx + T * w.nu + 1),
                        "api": "autompc_b200.evaluate_candidates(controllers, init_obs, T, sim_model, cost=) -- the call "
                               "the tuner makes; the timed region IS the public call (host init_obs in, trajectories "
                               "and costs out)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_burst"],
                             "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_burst"], "traffic": None,
                             "note": "cartpole-sized solves (4 672 MACs per sample-step, <= 1000 samples): launch- and "
                                     "latency-bound, neither roofline is approachable (SURVEY.md 8d)",
                             "kernel": "mppi_rollout + sim_step"}}
        if d.world == 1 and not args.no_cpu:
            rate, cores, sample = cpu_c5_rate(wl, float(os.environ.get("AMPC_CPU_BUDGET_S", "20")))
            line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                                    "host": cpu_info()}
        print(json.dumps(line), flush=True)
    for c in ctls:
        c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N>1 MPPI workloads: strong = K fixed in total (the BASELINE metric), weak = K per GPU fixed")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp16", "bf16"])
    ap.add_argument("--exchange", default="nvlink", choices=["nvlink", "nccl"],
                    help="N>1: how the ranks' softmax records meet (fused NVLink peer stores, or NCCL all-gather)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-weak-probe", action="store_true", help="skip the extra probes outside the timed region (weak-scaled problem at N>1, other operand format at N=1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                         "--master-addr 127.0.0.1 --master-port 29500 bench.py --gpus %d ..." % (args.gpus, args.gpus))
    wl = workload(args.workload, world=max(world, 1), scaling=args.scaling)
    if args.impl == "reference":
        reference_arm(args, wl, rank)
        return
    d = Dist(rank, world, local_rank)
    if args.workload in ("c3", "c2"):
        gpu_arm_mppi(args, wl, d)
    elif args.workload == "c4":
        gpu_arm_ilqr(args, wl, d)
    else:
        gpu_arm_c5(args, wl, d)
    d.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- MPC steps/sec of the MPPI solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2]

One "step" = one MPPI solve (``MPPI.run`` of autompc/control/mppi.py:154-168): shift, K rollouts of
H steps through the MLP dynamics with QuadCost, exponentiated-cost-weighted update.  Workload at
N=1 (and sharded over samples at N>1, strong scaling): BASELINE.json config C3 = HalfCheetah-dim
(nx=17, nu=6), MLP[23-256-256-256-17] ReLU, K=16384, H=50 (SURVEY.md 8d recipe, synthetic weights).

Prints ONE JSON line (rank 0).  ``value`` = solves/s with the observation already in HBM and
noise generated in-kernel (CUDA events, per-step, L2 flushed between steps, max over ranks);
``e2e`` = the same through the public ``Controller.run(state, new_obs)`` with host NumPy buffers
(pinned H2D of the observation + D2H of the control inside the timed region);
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


def workload(name):
    from autompc_b200.problems import cartpole_problem, halfcheetah_dim_problem
    if name == "c3":
        system, task, w, x0 = halfcheetah_dim_problem()
        return dict(system=system, task=task, weights=w, x0=x0, K=16384, H=50, sigma=1.0, lmda=1.0,
                    label="C3 HalfCheetah-dim nx=17 nu=6 MLP[23-256-256-256-17] relu, MPPI K=16384 H=50, "
                          "QuadCost Q=I R=0.01I F=10I, sigma=1 lmda=1")
    if name == "c2":
        system, task, w, x0 = cartpole_problem()
        return dict(system=system, task=task, weights=w, x0=x0, K=4096, H=30, sigma=1.0, lmda=1.0,
                    label="C2 cartpole nx=4 nu=1 MLP[5-64-64-4] relu, MPPI K=4096 H=30")
    raise SystemExit("unknown workload %s" % name)


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


def ncu_traffic(precision, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the rollout kernel, from the committed
    `ncu --set full` capture of this same command at N=1 (profiles/r01_mppi_tc_v11_ncu_full.md: 503 296 B read,
    1 792 B written -- the clipped-noise scratch stays in L2).  None when no capture exists for the configuration."""
    return 503296 + 1792 if (precision in ("bf16", "fp16") and world == 1) else None


# --------------------------------------------------------------------------- CPU arm ---
def cpu_port_rate(wl, budget_s, min_solves=2):
    """Times the float64 NumPy oracle (vectorised restatement of mppi.py:110-168) on the host.
    The sample is a slice of the K samples at the full horizon; cost is linear in K, so the rate of a
    full-K step is sample_rate * K_sample / K.  Returns (steps_per_s, cores, sample_description)."""
    from oracle.mppi_oracle import MLPParams, MPPIOracle, QuadCostParams
    w, task = wl["weights"], wl["task"]
    Q, R, F = task.get_cost().get_cost_matrices()
    p = MLPParams(w.W, w.b, w.act, w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, w.nx, w.nu)
    cost = QuadCostParams(Q, R, F, task.get_cost().get_goal())
    b = task.get_ctrl_bounds()
    K, H = wl["K"], wl["H"]

    def run(ks, n):
        np.random.seed(0)
        o = MPPIOracle(p, cost, b[:, 0], b[:, 1], horizon=H, num_path=ks, sigma=wl["sigma"], lmda=wl["lmda"])
        o.solve(wl["x0"])                                   # warm-up (BLAS threads, allocations)
        t0 = time.perf_counter()
        for _ in range(n):
            o.solve(wl["x0"])
        return (time.perf_counter() - t0) / n

    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 for every rank)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass
    probe_k = min(K, 512)
    t_probe = run(probe_k, 1)
    per_sample = t_probe / probe_k
    ks = int(min(K, max(probe_k, (budget_s / (min_solves + 1)) / per_sample)))
    ks = max(128, (ks // 128) * 128) if ks < K else K
    n = max(min_solves, int(budget_s / max(per_sample * ks, 1e-9)) - 1)
    n = min(n, 20)
    t = run(ks, n)
    rate = (1.0 / t) * (ks / K)
    try:
        import threadpoolctl
        cores = max([i.get("num_threads", 1) for i in threadpoolctl.threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count() or 1
    sample = ("float64 NumPy port of MPPI.run (vectorised cost loop), %d of %d samples x full H=%d, %d solves, "
              "%.2f s/solve-sample; steps/s scaled by %d/%d" % (ks, K, H, n, t, ks, K))
    return rate, int(cores), sample


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


def reference_arm(args, wl, rank):
    if rank != 0:
        return
    budget = float(os.environ.get("AMPC_REF_BUDGET_S", "60"))
    rate, cores, sample = cpu_port_rate(wl, budget, min_solves=max(2, min(args.steps, 5)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "host": cpu_info(),
                   "note": "reference is pure Python/NumPy/torch-CPU; for ctrl_dim=6 the unmodified reference raises "
                           "(mppi.py:139), so this arm is the documented restatement (oracle/mppi_oracle.py)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


# ------------------------------------------------------------------------------ GPU arm ---
def gpu_arm(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from autompc_b200 import MPPI, B200MLP, _abi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    system, task, w = wl["system"], wl["task"], wl["weights"]
    model = B200MLP(system, w, device=local_rank)
    np.random.seed(0)
    ctl = MPPI(system, task, model, horizon=wl["H"], num_path=wl["K"], sigma=wl["sigma"], lmda=wl["lmda"],
               seed=0, noise="philox", precision=args.precision, device=local_rank, group=group,
               exchange=args.exchange)
    nx, nu = w.nx, w.nu
    x0_dev = torch.tensor(wl["x0"], dtype=torch.float32, device=dev)
    u_dev = torch.zeros(nu, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        ctl.solve_device(x0_dev, u_dev, stream=stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        one_step()
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    n0 = _abi.launch_count()
    with ClockSampler(local_rank) as clk:
        barrier()
        for i in range(args.steps):
            flush.zero_()
            starts[i].record(stream)
            one_step()
            ends[i].record(stream)
        barrier()
    launches = _abi.launch_count() - n0
    per_step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    total_ms = float(per_step_ms.sum())
    u_host = u_dev.cpu().numpy()
    assert np.all(np.isfinite(u_host)), "non-finite control from the solve"

    # ---- end to end through Controller.run with host buffers (pinned H2D of obs, D2H of u inside)
    rng = np.random.default_rng(1)
    obs = [wl["x0"] + 0.01 * rng.normal(size=nx) for _ in range(args.steps + 3)]
    constate = np.concatenate([wl["x0"], np.zeros(nu)])
    for i in range(3):
        u, constate = ctl.run(constate, obs[i])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        u, constate = ctl.run(constate, obs[3 + i])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([total_ms, e2e_s, float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, launches = float(tmax[0]), float(tmax[1]), int(t[2])
    if rank == 0:
        peaks = measured_peaks()
        ms_per_step = total_ms / args.steps
        value = args.steps / (total_ms * 1e-3)
        flops = mlp_flops_per_solve(w, wl["K"], wl["H"])
        # dominant kernel = the rollout kernel; at N=1 it is the whole step.  Its average duration over
        # the timed region is ms_per_step minus the (tiny) merge kernel at N>1, which we do not subtract.
        achieved = flops / world / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "fp16"}.get(ctl.precision, "f32"), "data": "synthetic",
            "config": {"workload": wl["label"], "noise": "in-kernel Philox4x32-10", "precision": ctl.precision,
                       "parallelism": ("samples sharded over %d GPU(s), %d-float softmax record exchanged by %s"
                                       % (world, 2 + wl["H"] * nu,
                                          "NVLink peer stores + flags inside the rollout kernel (one launch per solve)"
                                          if ctl.exchange == "nvlink" else "one NCCL all-gather + merge kernel"))
                       if world > 1 else "1 GPU, one kernel per solve",
                       "timing": "CUDA events per step on the launch stream, sum over steps, max over ranks",
                       "l2": "256 MiB memset between timed steps (L2 flushed)",
                       "ms_per_step_min_median_max": [float(per_step_ms.min()), float(np.median(per_step_ms)),
                                                      float(per_step_ms.max())]},
            "clocks": clk.summary(),
            "e2e": {"value": args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 4 * nx,
                    "d2h_bytes_per_step": 4 * nu, "api": "autompc_b200.MPPI.run(state, new_obs) with NumPy float64 buffers",
                    "transfer": "host observation -> pinned float32 -> kernel parameters (H2D with the launch); control "
                                "written by the kernel's last CTA into mapped pinned host memory (D2H), one stream "
                                "synchronise per step" if world == 1 else "pinned H2D / D2H copies on the stream"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_burst"], "traffic": ncu_traffic(ctl.precision, world),
                         "flop_per_launch": flops / world, "peak_source": "bf16 dense burst, " + peaks["source"],
                         "kernel": "mppi_rollout (%s)" % ctl.precision},
        }
        if world == 1 and not args.no_cpu:
            rate, cores, sample = cpu_port_rate(wl, float(os.environ.get("AMPC_CPU_BUDGET_S", "20")))
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample, "host": cpu_info()}
        print(json.dumps(line), flush=True)
    ctl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp16", "bf16"])
    ap.add_argument("--exchange", default="nvlink", choices=["nvlink", "nccl"],
                    help="N>1: how the ranks' softmax records meet (fused NVLink peer stores, or NCCL all-gather)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                         "--master-addr 127.0.0.1 --master-port 29500 bench.py --gpus %d ..." % (args.gpus, args.gpus))
    wl = workload(args.workload)
    if args.impl == "reference":
        reference_arm(args, wl, rank)
        return
    gpu_arm(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()

"""Timing of the float64 batch kernels (pred_batch_kernel / rollout_batch_kernel) through the C ABI:
    python scripts/mlp_batch_bench.py [path/to/libampc_b200.so]
Wall clock around the ABI call (H2D + one launch + D2H, pageable host buffers) and the kernel alone (CUDA events,
ampc_mlp_debug_last_kernel_ms; median of the timed calls).  AMPC_MLP_BLOCKED=0|1 forces one kernel form."""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from autompc_b200 import _abi
if len(sys.argv) > 1:
    _abi.LIB_PATH = os.path.abspath(sys.argv[1])
from autompc_b200 import B200MLP
from autompc_b200.problems import halfcheetah_dim_problem

system, task, w, x0 = halfcheetah_dim_problem()
m = B200MLP(system, w)
rng = np.random.default_rng(0)
print("lib", _abi.LIB_PATH)
ref = {}
for batch, horizon in ((1, 1), (512, 1), (8192, 1), (65536, 1), (512, 20), (8192, 20)):
    X = rng.normal(size=(batch, 17)); U = 0.3 * rng.normal(size=(horizon, batch, 6))
    f = (lambda: m.pred_batch(X, U[0])) if horizon == 1 else (lambda: m.rollout_batch(X, U))
    out = f(); f()
    n = 5 if batch * horizon > 20000 else 20
    kms, ms = [], ctypes.c_float()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
        if hasattr(_abi.lib(), "ampc_mlp_debug_last_kernel_ms"):
            _abi.check(_abi.lib().ampc_mlp_debug_last_kernel_ms(m._h, ctypes.byref(ms)))
            kms.append(ms.value)
    dt = (time.perf_counter() - t0) / n
    fl = 2.0 * 141312 * batch * horizon
    k = float(np.median(kms)) if kms else float("nan")
    print("batch %6d horizon %2d: %9.3f ms per call, kernel %8.4f ms = %6.2f TFLOP/s fp64  checksum %.17g"
          % (batch, horizon, dt * 1e3, k, fl / (k * 1e-3) / 1e12, float(np.sum(out))))

# Jacobian kernels (one sample / knot per CTA): pred_diff_batch and the direct-transcription callbacks at batch H
for batch in (1, 7, 50, 400):
    X = rng.normal(size=(batch, 17)); U = 0.3 * rng.normal(size=(batch, 6))
    m.pred_diff_batch(X, U)
    kms, ms = [], ctypes.c_float()
    t0 = time.perf_counter()
    for _ in range(10):
        xn, jx, ju = m.pred_diff_batch(X, U)
        if hasattr(_abi.lib(), "ampc_mlp_debug_last_kernel_ms"):
            _abi.check(_abi.lib().ampc_mlp_debug_last_kernel_ms(m._h, ctypes.byref(ms)))
            kms.append(ms.value)
    dt = (time.perf_counter() - t0) / 10
    fl = 2.0 * (141312 + 23 * (256 * 256 * 2 + 256 * 17)) * batch
    k = float(np.median(kms)) if kms else float("nan")
    print("pred_diff_batch %4d: %9.3f ms per call, kernel %8.4f ms = %6.2f TFLOP/s fp64  checksum %.17g"
          % (batch, dt * 1e3, k, fl / (k * 1e-3) / 1e12, float(np.sum(jx) + np.sum(ju))))

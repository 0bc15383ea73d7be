"""GPU parity: the CUDA MPPI solve (through the C ABI) against
 (1) fixtures recorded from the UNMODIFIED reference (ctrl_dim == 1), replaying the
     global NumPy noise stream from the recorded seed ("external eps" mode), and
 (2) the float64 oracle restatement on the same noise (ctrl_dim > 1, other
     activations, dense Q, Philox noise fetched back from the device).

Tolerances (stated).  The reference is float64; the engine
computes in float32 ("fp32": CUDA-core FMA) or bf16 x bf16 -> fp32 tensor-core
products with fp32 state/cost ("bf16").  The softmax amplifies ABSOLUTE cost error
by 1/lmda, so the tolerance on the updated action sequence is per (precision, lmda,
cost scale); the numbers below are for the listed cases:
   fp32 : costs rtol 2e-5, action sequence atol 2e-3 (normalised controls, |.| <= 1)
   fp16 : costs rtol 1e-4, action sequence atol 2e-3   (IEEE-half operands = tf32's 11-bit significands)
   bf16 : costs rtol 1e-3, action sequence atol 2e-2
   exact: argmin(costs) (trajectory index), the shift indexing, the RNG draw order.
Measured on a B200 (scripts/measure_precision.py -> profiles/r02_precision.jsonl), worst case over the synthetic
cases below / at C3 full size: fp32 7e-5 / 1.3e-5, fp16 8.6e-4 / 2.1e-4, bf16 7.4e-3 / 2.1e-3 on the action
sequence; 5e-7, 3.1e-5, 2.6e-4 relative on the costs.
"""
import os

import numpy as np
import pytest

from oracle import philox
from oracle.mppi_oracle import MPPIOracle, QuadCostParams, mlp_pred_batch
from tests.helpers import GOLDEN, load_cartpole, synthetic_mlp

pytestmark = pytest.mark.gpu

TOL = {"fp32": dict(cost_rtol=2e-5, act_atol=2e-3), "fp16": dict(cost_rtol=1e-4, act_atol=2e-3),
       "bf16": dict(cost_rtol=1e-3, act_atol=2e-2)}

# sharded vs unsharded solves run the same per-sample arithmetic (Philox is keyed by the global sample index);
# they differ only in the fp32 summation order of the softmax merge: a few ulp of O(1) controls
SHARD_ATOL = 5e-6


def _engine(p, cost, umin, umax, **kw):
    from autompc_b200 import MPPI
    from tests.gpu_helpers import problem_of
    system, task, model = problem_of(p, cost, umin, umax)
    return MPPI(system, task, model, **kw)


def _check_solve(ctl, o, x0, eps, tol, check_argmin=True):
    """One solve on both sides with the same unclipped noise; compares costs, argmin, act, u."""
    ctl.act_sequence = o.act_sequence           # same warm start on both sides (float32 copy on the device)
    u = ctl.solve(x0, eps=eps)
    u_o = o.solve(x0, eps=eps.copy())
    costs, term = ctl.last_costs()
    ref = o.last_costs - o.term_const          # the engine returns the reference's terminal scalar separately
    np.testing.assert_allclose(costs, ref, rtol=tol["cost_rtol"], atol=tol["cost_rtol"] * np.abs(ref).max())
    np.testing.assert_allclose(term, o.term_const, rtol=max(tol["cost_rtol"], 1e-5) * 10,
                               atol=tol["cost_rtol"] * np.abs(ref).max())
    if check_argmin:
        assert int(np.argmin(costs)) == int(np.argmin(ref))
    np.testing.assert_allclose(ctl.act_sequence, o.act_sequence, rtol=0, atol=tol["act_atol"])
    np.testing.assert_allclose(u, u_o, rtol=0, atol=tol["act_atol"] * np.abs(o.ctrl_scale).max())
    return u, u_o


@pytest.mark.parametrize("name", ["mppi_cartpole_K256_H20", "mppi_cartpole_K100_H5", "mppi_cartpole_K512_H30",
                                  "mppi_cartpole_K4096_H30"])
def test_mppi_matches_unmodified_reference_fixture(name):
    """ctrl_dim == 1: the checker is the unmodified reference's recorded output."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mlp, cost, umin, umax, _, _ = load_cartpole()
    K, H = int(z["K"]), int(z["H"])
    tol = TOL["fp32"]
    np.random.seed(int(z["seed"]))
    ctl = _engine(mlp, cost, umin, umax, horizon=H, num_path=K, sigma=float(z["sigma"]), lmda=float(z["lmda"]),
                  noise="numpy", precision="fp32")
    np.testing.assert_allclose(ctl.act_sequence, z["act0"], rtol=0, atol=1e-7)   # same draws, same order (mppi.py:99)
    constate = np.zeros(5)
    for s in range(int(z["n_steps"])):
        x = z["x0_%d" % s]
        if s > 0:
            # free-running float32 warm start stays near the reference's; then re-synchronise so that every
            # step is compared from an identical starting point (the learned cartpole dynamics near the
            # upright are unstable: H-step rollouts amplify a 1e-6 warm-start difference)
            np.testing.assert_allclose(ctl.act_sequence, z["act_%d" % (s - 1)], rtol=0, atol=10 * tol["act_atol"])
            ctl.act_sequence = z["act_%d" % (s - 1)]
        u, constate = ctl.run(constate, x)            # draws K*H normals from the global stream like mppi.py:126
        costs, term = ctl.last_costs()
        ref = z["costs_%d" % s]
        # the terminal scalar (last sample's final state, F up to 3000: mppi.py:79-82) is common to all
        # samples and cancels in the softmax; float32 state error after H steps is amplified by F there,
        # so the absolute level gets 1e-4 and the softmax-relevant differences the fp32 tolerance
        np.testing.assert_allclose(costs + term, ref, rtol=1e-4)
        np.testing.assert_allclose(costs - costs.min(), ref - ref.min(), rtol=0,
                                   atol=tol["cost_rtol"] * np.abs(ref).max())
        assert int(np.argmin(costs)) == int(z["argmin_%d" % s])                  # bit-exact trajectory index
        np.testing.assert_allclose(ctl.act_sequence, z["act_%d" % s], rtol=0, atol=tol["act_atol"])
        np.testing.assert_allclose(u, z["u_%d" % s], rtol=0, atol=tol["act_atol"] * 20.0)
        np.testing.assert_allclose(constate, np.concatenate([x, u]))
    ctl.close()


CASES = [
    # nx, nu, hidden, act, K, H, sigma, lmda, dense
    (17, 6, [256, 256, 256], "relu", 2048, 50, 1.0, 1.0, False),      # C3 dims at reduced K
    (17, 6, [64, 64], "tanh", 300, 12, 0.5, 0.7, True),
    (4, 1, [32], "sigmoid", 77, 7, 0.3, 0.5, False),
    (3, 2, [48, 24, 16], "selu", 129, 9, 0.8, 2.0, True),
    (4, 1, [64, 64], "relu", 1, 2, 1.0, 1.0, False),                  # K=1, minimum horizon
    (6, 3, [100, 60], "relu", 1000, 20, 1.0, 1.0, False),             # widths that are not powers of two
]


@pytest.mark.parametrize("nx,nu,hidden,act,K,H,sigma,lmda,dense", CASES)
def test_mppi_fp32_matches_oracle(nx, nu, hidden, act, K, H, sigma, lmda, dense):
    """ctrl_dim > 1 (restated oracle), all activations, dense Q/R/F, ragged K."""
    rng = np.random.default_rng(5)
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
    if dense:
        A, B, C = rng.normal(size=(nx, nx)), rng.normal(size=(nu, nu)), rng.normal(size=(nx, nx))
        cost = QuadCostParams(A @ A.T / nx, 0.01 * (B @ B.T), C @ C.T, goal=0.1 * rng.normal(size=nx))
    else:
        cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx), goal=0.05 * rng.normal(size=nx))
    umax = rng.uniform(0.5, 2.0, size=nu)
    umin = -umax * rng.uniform(0.5, 1.0, size=nu)
    np.random.seed(1)
    ctl = _engine(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda, noise="numpy",
                  precision="fp32")
    np.random.seed(1)
    o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda)
    x0 = rng.normal(size=nx)
    for _ in range(3):                                   # warm-started consecutive solves (shift indexing)
        eps = o.sample_eps()
        u, _ = _check_solve(ctl, o, x0, eps, TOL["fp32"], check_argmin=K > 1)
        x0 = mlp_pred_batch(p, x0[None], u[None])[0]
    ctl.close()


def test_mppi_per_sample_terminal_option():
    """terminal='per_sample' (explicit non-reference option): each sample pays its own terminal cost."""
    p = synthetic_mlp(5, 2, [32, 32], seed=9)
    cost = QuadCostParams(np.eye(5), 0.1 * np.eye(2), 50 * np.eye(5))
    umin, umax = [-1.0, -2.0], [1.0, 2.0]
    np.random.seed(4)
    ctl = _engine(p, cost, umin, umax, horizon=8, num_path=200, noise="numpy", precision="fp32",
                  terminal="per_sample")
    np.random.seed(4)
    o = MPPIOracle(p, cost, umin, umax, horizon=8, num_path=200)
    x0 = np.linspace(-1, 1, 5)
    eps = o.sample_eps()
    act = o.act_sequence.copy()
    ctl.solve(x0, eps=eps)
    costs_ref, eps_c = o.do_rollouts(x0, eps.copy())
    per = costs_ref - o.term_const + np.einsum("ki,ij,kj->k", o.last_path - cost.goal, cost.F, o.last_path - cost.goal)
    costs, _ = ctl.last_costs()
    np.testing.assert_allclose(costs, per, rtol=2e-5)
    S = np.exp(-(per - per.min()))
    shifted = np.concatenate([act[1:], act[-1:]])
    np.testing.assert_allclose(ctl.act_sequence, shifted + np.einsum("hkj,k->hj", eps_c, S / S.sum()), atol=2e-3)
    ctl.close()


def test_philox_stream_matches_restatement():
    """The in-kernel generator, restated in NumPy (oracle/philox.py): integer stream identical,
    float transforms to ~1e-5; and moments of a large draw."""
    p = synthetic_mlp(4, 6, [16], seed=1)
    cost = QuadCostParams(np.eye(4), np.eye(6), np.eye(4))
    ctl = _engine(p, cost, -np.ones(6), np.ones(6), horizon=11, num_path=4099, sigma=0.49, seed=1234567890123,
                  precision="fp32")
    e = ctl.philox_noise(counter=7)
    ref = philox.mppi_noise(1234567890123, 7, 11, 4099, 6, 0.49)
    np.testing.assert_allclose(e, ref, rtol=0, atol=2e-5)
    assert abs(e.mean()) < 5e-3 and abs(e.std() - 0.7) < 5e-3
    e2 = ctl.philox_noise(counter=8)
    assert abs(np.corrcoef(e.ravel(), e2.ravel())[0, 1]) < 0.01      # different solves, independent noise
    ctl.close()


@pytest.mark.parametrize("precision", ["fp32"])
def test_philox_mode_equals_external_eps_mode(precision):
    """Performance mode (in-kernel Philox) == parity mode fed the same noise, and == the oracle."""
    p = synthetic_mlp(17, 6, [64, 64], seed=2)
    cost = QuadCostParams(np.eye(17), 0.01 * np.eye(6), 10 * np.eye(17))
    kw = dict(horizon=15, num_path=777, sigma=1.0, lmda=1.0, seed=99, precision=precision)
    np.random.seed(0)
    a = _engine(p, cost, -np.ones(6), np.ones(6), **kw)
    np.random.seed(0)
    b = _engine(p, cost, -np.ones(6), np.ones(6), **kw)
    np.random.seed(0)
    o = MPPIOracle(p, cost, -np.ones(6), np.ones(6), horizon=15, num_path=777)
    x0 = np.random.default_rng(3).normal(size=17)
    for step in range(2):
        eps = a.philox_noise().astype(np.float64)        # noise of solve `cur_step`
        ua = a.solve(x0)                                 # Philox in-kernel
        ub = b.solve(x0, eps=eps)                        # same numbers uploaded
        np.testing.assert_allclose(ua, ub, rtol=0, atol=1e-6)
        np.testing.assert_allclose(a.last_costs()[0], b.last_costs()[0], rtol=1e-6)
        uo = o.solve(x0, eps=eps)
        np.testing.assert_allclose(ua, uo, rtol=0, atol=TOL[precision]["act_atol"])
    a.close()
    b.close()


def test_sharded_partials_merge_equals_single_handle():
    """Multi-GPU data path on one device: 3 shards (k_offset) -> rollout_partial -> merge
    == one handle over all samples (Philox keyed by GLOBAL sample index)."""
    import ctypes as C
    import torch
    from autompc_b200 import _abi
    p = synthetic_mlp(17, 6, [64, 64], seed=2)
    cost = QuadCostParams(np.eye(17), 0.01 * np.eye(6), 10 * np.eye(17))
    K, H = 1000, 10
    np.random.seed(0)
    full = _engine(p, cost, -np.ones(6), np.ones(6), horizon=H, num_path=K, seed=5, precision="fp32")
    act0 = full.act_sequence
    x0 = np.random.default_rng(1).normal(size=17)
    u_full = full.solve(x0)
    lib = _abi.lib()
    shards, recs = [], []
    dev = torch.device("cuda", 0)
    x0_d = torch.tensor(x0, dtype=torch.float32, device=dev)
    splits = [(0, 334), (334, 333), (667, 333)]
    for off, n in splits:                        # shard geometry through the C ABI directly
        cfg = _abi.MppiCfg(n, H, 17, 6, 1.0, 1.0, 0, 0, off, K, 0)
        h = C.c_void_p()
        _abi.check(lib.ampc_mppi_create(C.byref(h), C.byref(cfg), C.byref(full._mlp_holder.desc),
                                        C.byref(full._cost_holder.desc)))
        a = np.ascontiguousarray(act0)
        _abi.check(lib.ampc_mppi_set_act_seq(h, _abi.dptr(a)))
        rec = torch.zeros(lib.ampc_mppi_record_floats(h), dtype=torch.float32, device=dev)
        _abi.check(lib.ampc_mppi_rollout_partial(h, x0_d.data_ptr(), None, 5, 0, rec.data_ptr(), None))
        shards.append(h)
        recs.append(rec)
    torch.cuda.synchronize()
    allrec = torch.cat(recs).contiguous()
    u_d = torch.zeros(6, dtype=torch.float32, device=dev)
    for h in shards:
        _abi.check(lib.ampc_mppi_merge(h, allrec.data_ptr(), len(shards), u_d.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_allclose(u_d.cpu().numpy(), u_full, rtol=0, atol=SHARD_ATOL)
    for h in shards:
        a = np.empty((H, 6))
        _abi.check(lib.ampc_mppi_get_act_seq(h, _abi.dptr(a)))
        np.testing.assert_allclose(a, full.act_sequence, rtol=0, atol=SHARD_ATOL)
        lib.ampc_mppi_destroy(h)
    full.close()


def test_closed_loop_through_plugin_surface():
    """simulate()-style loop (utils/simulation.py:45-63) with the engine as Controller and B200MLP as
    the stepped model: fixed-size state, fresh arrays, reset() redraws (mppi.py:107-108)."""
    from autompc_b200 import MPPI, B200MLP
    from autompc_b200.problems import cartpole_problem
    from autompc_b200.mlp import MLPWeights
    z = np.load(os.path.join(GOLDEN, "cartpole_mlp.npz"))
    system, task, w, x0 = cartpole_problem(MLPWeights.from_npz(z))
    model = B200MLP(system, w)
    np.random.seed(0)
    ctl = MPPI(system, task, model, horizon=20, num_path=512, precision="fp32")
    assert ctl.state_dim == 5
    x, constate = x0.copy(), np.concatenate([x0, np.zeros(1)])
    for _ in range(5):
        u, constate = ctl.run(constate, x)
        assert u.shape == (1,) and constate.shape == (5,) and np.all(np.abs(u) <= 20.0 + 1e-5)
        x = model.pred(x, u)
    a = ctl.act_sequence
    ctl.reset()
    assert ctl.cur_step == 0 and not np.allclose(a, ctl.act_sequence)
    ctl.close()


# ----------------------------------------------------------------------------- tcgen05 (bf16) path ---
TC_CASES = [
    # nx, nu, hidden, act, K, H, sigma, lmda, force_cg
    (4, 1, [64, 64], "relu", 4096, 30, 1.0, 1.0, None),               # C2 dims (weights fit one CTA: cta_group::1)
    (4, 1, [64, 64], "relu", 300, 20, 1.0, 1.0, "2"),                 # same network as a CTA pair, ragged K
    (17, 6, [256, 256, 256], "relu", 2048, 50, 1.0, 1.0, None),       # C3 dims: cta_group::2 (half the weights per CTA)
    (17, 6, [128, 128], "tanh", 555, 12, 0.5, 0.7, "1"),
    (17, 6, [100, 60], "sigmoid", 129, 9, 0.8, 2.0, "2"),             # widths padded to 32
    (3, 2, [48, 24, 16, 40], "selu", 77, 7, 0.3, 0.5, None),          # 4 hidden layers
    (4, 1, [32], "relu", 1, 2, 1.0, 1.0, None),                       # K=1, minimum horizon
    # input-block layouts: NXP=32 with two control-only K-steps (stored ahead of the state), dense Q/R/F
    (32, 20, [256, 128], "relu", 513, 6, 0.6, 1.5, None, True),
    (16, 9, [64], "relu", 200, 5, 1.0, 1.0, None, True),              # NXP=16: no K-step mixes state and controls
    (8, 3, [128, 128, 128], "relu", 640, 10, 0.7, 1.0, None, True),   # NXP=8: the mixed K-step is the first one
]


@pytest.mark.parametrize("case", TC_CASES)
def test_mppi_bf16_tensor_core_matches_oracle(case, monkeypatch):
    """bf16 x bf16 -> fp32 tcgen05 products, fp32 state / cost / softmax: stated bf16 tolerance."""
    nx, nu, hidden, act, K, H, sigma, lmda, force_cg = case[:9]
    dense = len(case) > 9 and case[9]
    if force_cg:
        monkeypatch.setenv("AMPC_TC_FORCE_CG", force_cg)
    else:
        monkeypatch.delenv("AMPC_TC_FORCE_CG", raising=False)
    rng = np.random.default_rng(5)
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
    if dense:
        A, B, C = rng.normal(size=(nx, nx)), rng.normal(size=(nu, nu)), rng.normal(size=(nx, nx))
        cost = QuadCostParams(A @ A.T / nx, 0.01 * (B @ B.T) / nu, C @ C.T / nx, goal=0.1 * rng.normal(size=nx))
    else:
        cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx), goal=0.05 * rng.normal(size=nx))
    umax = rng.uniform(0.5, 2.0, size=nu)
    umin = -umax * rng.uniform(0.5, 1.0, size=nu)
    np.random.seed(1)
    ctl = _engine(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda, noise="numpy",
                  precision="bf16")
    assert ctl.precision == "bf16"
    np.random.seed(1)
    o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda)
    x0 = rng.normal(size=nx)
    for _ in range(3):
        eps = o.sample_eps()
        # argmin is exact unless two samples are closer than the bf16 cost tolerance
        ref_sorted = None
        u, _ = _check_solve(ctl, o, x0, eps, TOL["bf16"], check_argmin=False)
        costs, _t = ctl.last_costs()
        ref = o.last_costs - o.term_const
        ref_sorted = np.sort(ref)
        if K > 1 and ref_sorted[1] - ref_sorted[0] > 4 * TOL["bf16"]["cost_rtol"] * abs(ref_sorted[0]):
            assert int(np.argmin(costs)) == int(np.argmin(ref))
        x0 = mlp_pred_batch(p, x0[None], u[None])[0]
    ctl.close()


def test_mppi_bf16_philox_equals_external_eps_and_fp32():
    """Same Philox noise in both kernels: bf16 result within bf16 tolerance of the fp32 kernel and the oracle."""
    p = synthetic_mlp(17, 6, [256, 256, 256], seed=2)
    cost = QuadCostParams(np.eye(17), 0.01 * np.eye(6), 10 * np.eye(17))
    kw = dict(horizon=20, num_path=1000, sigma=1.0, lmda=1.0, seed=77)
    np.random.seed(0)
    a = _engine(p, cost, -np.ones(6), np.ones(6), precision="bf16", **kw)
    np.random.seed(0)
    b = _engine(p, cost, -np.ones(6), np.ones(6), precision="bf16", **kw)
    np.random.seed(0)
    c = _engine(p, cost, -np.ones(6), np.ones(6), precision="fp32", **kw)
    x0 = np.random.default_rng(3).normal(size=17)
    eps = a.philox_noise().astype(np.float64)
    np.testing.assert_array_equal(eps, c.philox_noise().astype(np.float64))     # one generator for both kernels
    ua, ub, uc = a.solve(x0), b.solve(x0, eps=eps), c.solve(x0)
    np.testing.assert_allclose(ua, ub, rtol=0, atol=1e-6)
    np.testing.assert_allclose(a.last_costs()[0], b.last_costs()[0], rtol=1e-6)
    np.testing.assert_allclose(a.last_costs()[0], c.last_costs()[0], rtol=TOL["bf16"]["cost_rtol"])
    np.testing.assert_allclose(ua, uc, rtol=0, atol=TOL["bf16"]["act_atol"])
    for e in (a, b, c):
        e.close()


def test_mppi_full_size_c3_properties():
    """BASELINE config 3 at full size (K=16384, H=50) on one GPU, checked through size-independent
    properties: (i) the update is a convex combination of the clipped noise => every entry of
    act_sequence - shift(act_sequence) lies inside the clip box; (ii) costs of the first 256 samples
    equal a K=256 solve on the same global sample indices (samples are independent);
    (iii) identical seeds => identical result; (iv) auto precision picks the tensor-core path.
    (The oracle comparison at this size is test_mppi_c3_full_size_matches_oracle.)"""
    from autompc_b200.problems import halfcheetah_dim_problem
    from autompc_b200 import MPPI, B200MLP
    system, task, w, x0 = halfcheetah_dim_problem()
    model = B200MLP(system, w)
    np.random.seed(0)
    big = MPPI(system, task, model, horizon=50, num_path=16384, seed=3)
    assert big.precision == "fp16"                 # 'auto': tensor-core kernel, IEEE-half operands
    act0 = big.act_sequence
    u1 = big.solve(x0)
    costs_big, _ = big.last_costs()
    shifted = np.concatenate([act0[1:], act0[-1:]])
    upd = big.act_sequence - shifted
    assert np.all(np.isfinite(upd)) and np.all(shifted + upd <= 1 + 1e-5) and np.all(shifted + upd >= -1 - 1e-5)
    np.random.seed(0)
    small = MPPI(system, task, model, horizon=50, num_path=256, seed=3)
    small.solve(x0)
    np.testing.assert_allclose(small.last_costs()[0], costs_big[:256], rtol=1e-6)
    np.random.seed(0)
    again = MPPI(system, task, model, horizon=50, num_path=16384, seed=3)
    u2 = again.solve(x0)
    np.testing.assert_array_equal(again.last_costs()[0], costs_big)
    np.testing.assert_allclose(u1, u2, rtol=0, atol=1e-6)      # merge order of CTA partials may differ
    for e in (big, small, again):
        e.close()


def test_fused_peer_exchange_two_shards_one_device():
    """The NVLink peer-exchange tail (stores into every rank's mailbox + flags + merge, ampc_mppi_solve_fused)
    exercised on ONE device: two shard handles connected in-process, launched on two streams so that both
    kernels are resident together; both must reproduce the single-handle solve."""
    import ctypes as C
    import torch
    from autompc_b200 import _abi
    p = synthetic_mlp(17, 6, [64, 64], seed=2)
    cost = QuadCostParams(np.eye(17), 0.01 * np.eye(6), 10 * np.eye(17))
    K, H = 600, 10
    np.random.seed(0)
    full = _engine(p, cost, -np.ones(6), np.ones(6), horizon=H, num_path=K, seed=5, precision="bf16")
    act0 = np.ascontiguousarray(full.act_sequence)
    x0 = np.random.default_rng(1).normal(size=17)
    lib = _abi.lib()
    dev = torch.device("cuda", 0)
    x0_d = torch.tensor(x0, dtype=torch.float32, device=dev)
    hs = (C.c_void_p * 2)()
    for r, (off, n) in enumerate([(0, 300), (300, 300)]):
        cfg = _abi.MppiCfg(n, H, 17, 6, 1.0, 1.0, 0, _abi.PREC_CODES["bf16"], off, K, 0)
        h = C.c_void_p()
        _abi.check(lib.ampc_mppi_create(C.byref(h), C.byref(cfg), C.byref(full._mlp_holder.desc),
                                        C.byref(full._cost_holder.desc)))
        _abi.check(lib.ampc_mppi_set_act_seq(h, _abi.dptr(act0)))
        hs[r] = h
    for r in range(2):
        _abi.check(lib.ampc_mppi_connect_peers_local(hs[r], 2, r, hs))
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    us = [torch.zeros(6, dtype=torch.float32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    u_ref = []
    for step in range(3):                                  # consecutive solves: both mailbox slots get reused
        u_ref.append(full.solve(x0))
        for r in range(2):
            _abi.check(lib.ampc_mppi_solve_fused(hs[r], x0_d.data_ptr(), None, 5, step, us[r].data_ptr(),
                                                 streams[r].cuda_stream))
        torch.cuda.synchronize()
        for r in range(2):
            np.testing.assert_allclose(us[r].cpu().numpy(), u_ref[-1], rtol=0, atol=SHARD_ATOL)
    for r in range(2):
        a = np.empty((H, 6))
        _abi.check(lib.ampc_mppi_get_act_seq(hs[r], _abi.dptr(a)))
        np.testing.assert_allclose(a, full.act_sequence, rtol=0, atol=SHARD_ATOL)
        lib.ampc_mppi_destroy(hs[r])
    full.close()


def _sumcost_problem(z):
    """cartpole task whose cost is QuadCost + QuadCost with DIFFERENT goals (a reference SumCost that is not is_quad)."""
    from autompc_b200 import B200MLP
    from autompc_b200.plugin import QuadCost, System, Task
    from tests.gpu_helpers import weights_of
    mlp, _, umin, umax, _, _ = load_cartpole()
    system = System(["theta", "omega", "x", "dx"], ["u"])
    system.dt = 0.05
    task = Task(system)
    task.set_ctrl_bounds(np.asarray(umin, dtype=np.float64), np.asarray(umax, dtype=np.float64))
    task.set_cost(QuadCost(system, z["Q1"], z["R1"], z["F1"], goal=z["g1"]) +
                  QuadCost(system, z["Q2"], z["R2"], z["F2"], goal=z["g2"]))
    return system, task, B200MLP(system, weights_of(mlp))


def test_mppi_sumcost_matches_unmodified_reference_fixture():
    """the reference evaluates the SumCost term by term (sum_cost.py:52-81); the engine folds it
    into one quadratic + constants on the host.  Checker: the unmodified reference's recorded costs / controls."""
    from autompc_b200 import MPPI
    z = np.load(os.path.join(GOLDEN, "mppi_cartpole_sumcost_K256_H20.npz"))
    system, task, model = _sumcost_problem(z)
    tol = TOL["fp32"]
    np.random.seed(int(z["seed"]))
    ctl = MPPI(system, task, model, horizon=int(z["H"]), num_path=int(z["K"]), sigma=float(z["sigma"]),
               lmda=float(z["lmda"]), noise="numpy", precision="fp32")
    np.testing.assert_allclose(ctl.act_sequence, z["act0"], rtol=1e-7, atol=1e-7)   # float32 copy on the device
    constate = np.zeros(5)
    for s in range(int(z["n_steps"])):
        if s > 0:
            ctl.act_sequence = z["act_%d" % (s - 1)]
        u, constate = ctl.run(constate, z["x0_%d" % s])
        costs, term = ctl.last_costs()                      # includes the fold's constants
        ref = z["costs_%d" % s]
        np.testing.assert_allclose(costs + term, ref, rtol=1e-4)
        np.testing.assert_allclose(costs - costs.min(), ref - ref.min(), rtol=0, atol=tol["cost_rtol"] * np.abs(ref).max())
        assert int(np.argmin(costs)) == int(z["argmin_%d" % s])
        np.testing.assert_allclose(ctl.act_sequence, z["act_%d" % s], rtol=0, atol=tol["act_atol"])
        np.testing.assert_allclose(u, z["u_%d" % s], rtol=0, atol=tol["act_atol"] * 20.0)
    ctl.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mppi_sumcost_per_sample_terminal_matches_oracle(precision):
    """Folded stage and terminal terms have different goals (ampc_quad_cost.goal_term): per-sample terminal mode."""
    from autompc_b200 import MPPI
    from oracle.mppi_oracle import SumQuadCostParams
    zf = np.load(os.path.join(GOLDEN, "mppi_cartpole_sumcost_K256_H20.npz"))
    # the fixture's stage terms, but mild terminal weights: F up to 3000 on the unstable cartpole model turns the
    # bf16 state error into O(10 %) terminal-cost error, which would test the dynamics, not the fold
    z = {k: zf[k] for k in ("Q1", "R1", "g1", "Q2", "R2", "g2")}
    z["F1"], z["F2"] = np.diag([2.0, 3.0, 0.5, 1.0]), np.diag([1.0, 0.0, 2.0, 0.5])
    system, task, model = _sumcost_problem(z)
    mlp, _, umin, umax, _, _ = load_cartpole()
    terms = [QuadCostParams(z["Q1"], z["R1"], z["F1"], z["g1"]), QuadCostParams(z["Q2"], z["R2"], z["F2"], z["g2"])]
    K, H = 384, 10
    np.random.seed(2)
    ctl = MPPI(system, task, model, horizon=H, num_path=K, sigma=0.5, lmda=5.0, noise="numpy", precision=precision,
               terminal="per_sample")
    np.random.seed(2)
    o = MPPIOracle(mlp, SumQuadCostParams(terms), umin, umax, horizon=H, num_path=K, sigma=0.5, lmda=5.0)
    x0 = np.array([0.3, 0.1, -0.2, 0.05])
    eps = o.sample_eps()
    ctl.act_sequence = o.act_sequence
    ctl.solve(x0, eps=eps)
    costs_o, eps_c = o.do_rollouts(x0, eps.copy())
    costs_o = costs_o - o.term_const + np.array([SumQuadCostParams(terms).eval_term_obs_cost(x) for x in o.last_path])
    costs, _ = ctl.last_costs()
    # the learned cartpole dynamics amplify rounding (see the fixture tests); the check here is the fold, so the
    # tolerance is the cost tolerance of the synthetic cases x10 (fp32) / x20 (bf16)
    k = 10 if precision == "fp32" else 20
    np.testing.assert_allclose(costs + ctl._term_const, costs_o, rtol=TOL[precision]["cost_rtol"] * k,
                               atol=TOL[precision]["cost_rtol"] * np.abs(costs_o).max())
    ctl.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_handles_with_different_shared_memory_needs_coexist(precision):
    """The dynamic shared-memory limit is a per-kernel attribute: a controller created earlier with the LARGER need
    (longer horizon) must still launch after a smaller one was created (regression: invalid-argument launch error
    when evaluate_candidates mixed horizons)."""
    p = synthetic_mlp(4, 1, [64, 64], seed=2)
    cost = QuadCostParams(np.eye(4), 0.1 * np.eye(1), np.eye(4))
    big = _engine(p, cost, [-1.0], [1.0], horizon=60, num_path=256, seed=1, precision=precision)
    small = _engine(p, cost, [-1.0], [1.0], horizon=5, num_path=256, seed=1, precision=precision)
    x0 = np.array([0.1, 0.2, -0.1, 0.0])
    u_small = small.solve(x0)
    u_big = big.solve(x0)                     # failed with cudaErrorInvalidValue before the limit became monotone
    assert np.all(np.isfinite(u_small)) and np.all(np.isfinite(u_big))
    big.close()
    small.close()


# ----------------------------------------------------------------------------- round-2 regressions / parity holes ---
def test_external_eps_after_peer_connect_and_closed_loop():
    """Regression (round-1 review): the external-eps staging branch of ampc_mppi_solve_host freed the peer mailboxes and
    the closed-loop buffers.  Sequence on ONE handle pair: connect peers -> fused solve -> external-eps solve (grows the
    staging buffers) -> fused solve again -> destroy; and closed loop -> external eps -> closed loop on another."""
    import ctypes as C
    import torch
    from autompc_b200 import MPPI, B200MLP, _abi, simulate
    from tests.gpu_helpers import problem_of
    p = synthetic_mlp(17, 6, [64, 64], seed=2)
    cost = QuadCostParams(np.eye(17), 0.01 * np.eye(6), 10 * np.eye(17))
    K, H = 512, 8
    np.random.seed(0)
    full = _engine(p, cost, -np.ones(6), np.ones(6), horizon=H, num_path=K, seed=5, precision="fp32")
    act0 = np.ascontiguousarray(full.act_sequence)
    x0 = np.random.default_rng(1).normal(size=17)
    lib = _abi.lib()
    dev = torch.device("cuda", 0)
    x0_d = torch.tensor(x0, dtype=torch.float32, device=dev)
    hs = (C.c_void_p * 2)()
    for r, (off, n) in enumerate([(0, 256), (256, 256)]):
        cfg = _abi.MppiCfg(n, H, 17, 6, 1.0, 1.0, 0, _abi.PREC_CODES["fp32"], off, K, 0)
        h = C.c_void_p()
        _abi.check(lib.ampc_mppi_create(C.byref(h), C.byref(cfg), C.byref(full._mlp_holder.desc),
                                        C.byref(full._cost_holder.desc)))
        _abi.check(lib.ampc_mppi_set_act_seq(h, _abi.dptr(act0)))
        hs[r] = h
    for r in range(2):
        _abi.check(lib.ampc_mppi_connect_peers_local(hs[r], 2, r, hs))
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    us = [torch.zeros(6, dtype=torch.float32, device=dev) for _ in range(2)]

    def fused(step):
        for r in range(2):
            _abi.check(lib.ampc_mppi_solve_fused(hs[r], x0_d.data_ptr(), None, 5, step, us[r].data_ptr(),
                                                 streams[r].cuda_stream))
        torch.cuda.synchronize()
        return [u.cpu().numpy() for u in us]

    u_ref = full.solve(x0)
    for u in fused(0):
        np.testing.assert_allclose(u, u_ref, rtol=0, atol=SHARD_ATOL)
    # external noise on the connected handles (this is the branch that used to free d_mail / d_rec / d_peer)
    eps = np.random.default_rng(2).normal(size=(H, 256, 6))
    for r in range(2):
        u = np.empty(6)
        _abi.check(lib.ampc_mppi_solve_host(hs[r], _abi.dptr(x0), _abi.dptr(eps), 5, 1, _abi.dptr(u)))
        assert np.all(np.isfinite(u))
    for r in range(2):                                     # the host solves updated each shard on its own: re-align
        _abi.check(lib.ampc_mppi_set_act_seq(hs[r], _abi.dptr(act0)))
    full.act_sequence = act0
    u_ref = full.solve(x0)                                 # cur_step 1 on `full`; fused() must use the same counter
    for u in fused(1):
        np.testing.assert_allclose(u, u_ref, rtol=0, atol=SHARD_ATOL)
    for r in range(2):
        _abi.check(lib.ampc_mppi_destroy(hs[r]))          # double free / double close showed up here
    full.close()
    # closed loop -> external eps -> closed loop
    system, task, model = problem_of(p, cost, -np.ones(6), np.ones(6))
    np.random.seed(3)
    ctl = MPPI(system, task, model, horizon=H, num_path=256, seed=9, precision="fp32")
    a = simulate(ctl, x0, sim_model=model, max_steps=6)
    ctl.solve(x0, eps=np.random.default_rng(4).normal(size=(H, 256, 6)))
    np.random.seed(3)
    ctl.reset()                                           # same draw as the constructor's, cur_step back to 0
    b = simulate(ctl, x0, sim_model=model, max_steps=6)   # T <= cl_T: reuses d_cl (was freed memory before the fix)
    np.testing.assert_array_equal(a.obs, b.obs)
    np.testing.assert_array_equal(a.ctrls, b.ctrls)
    torch.cuda.synchronize()
    ctl.close()


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_mppi_c3_full_size_matches_oracle(precision):
    """BASELINE config C3 at its OWN size (K=16384, H=50, MLP 23-256-256-256-17) on the tensor-core kernel against the
    float64 oracle on the same uploaded noise (mppi.py:120-152): costs, arg-min sample, action sequence, control."""
    from autompc_b200.problems import halfcheetah_dim_problem
    from autompc_b200 import MPPI, B200MLP
    from oracle.mppi_oracle import MLPParams
    system, task, w, x0 = halfcheetah_dim_problem()
    Q, R, F = task.get_cost().get_cost_matrices()
    p = MLPParams(w.W, w.b, w.act, w.xu_mean, w.xu_std, w.dy_mean, w.dy_std, w.nx, w.nu)
    cost = QuadCostParams(Q, R, F, task.get_cost().get_goal())
    b = task.get_ctrl_bounds()
    np.random.seed(0)
    ctl = MPPI(system, task, B200MLP(system, w), horizon=50, num_path=16384, noise="numpy", precision=precision)
    np.random.seed(0)
    o = MPPIOracle(p, cost, b[:, 0], b[:, 1], horizon=50, num_path=16384)
    eps = o.sample_eps()
    _check_solve(ctl, o, x0, eps, TOL[precision], check_argmin=False)
    costs, _ = ctl.last_costs()
    ref = o.last_costs - o.term_const
    srt = np.sort(ref)
    if srt[1] - srt[0] > 4 * TOL[precision]["cost_rtol"] * abs(srt[0]):
        assert int(np.argmin(costs)) == int(np.argmin(ref))
    ctl.close()


@pytest.mark.parametrize("case", TC_CASES)
def test_mppi_fp16_tensor_core_matches_oracle(case, monkeypatch):
    """IEEE-half operands on the same tcgen05 kernel: fp32-class tolerance (SURVEY 7.3 planned tf32 at 5e-3)."""
    nx, nu, hidden, act, K, H, sigma, lmda, force_cg = case[:9]
    dense = len(case) > 9 and case[9]
    if force_cg:
        monkeypatch.setenv("AMPC_TC_FORCE_CG", force_cg)
    else:
        monkeypatch.delenv("AMPC_TC_FORCE_CG", raising=False)
    rng = np.random.default_rng(5)
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
    if dense:
        A, B, C = rng.normal(size=(nx, nx)), rng.normal(size=(nu, nu)), rng.normal(size=(nx, nx))
        cost = QuadCostParams(A @ A.T / nx, 0.01 * (B @ B.T) / nu, C @ C.T / nx, goal=0.1 * rng.normal(size=nx))
    else:
        cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx), goal=0.05 * rng.normal(size=nx))
    umax = rng.uniform(0.5, 2.0, size=nu)
    umin = -umax * rng.uniform(0.5, 1.0, size=nu)
    np.random.seed(1)
    ctl = _engine(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda, noise="numpy",
                  precision="fp16")
    assert ctl.precision == "fp16"
    np.random.seed(1)
    o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda)
    x0 = rng.normal(size=nx)
    for _ in range(3):
        eps = o.sample_eps()
        u, _ = _check_solve(ctl, o, x0, eps, TOL["fp16"], check_argmin=False)
        costs, _t = ctl.last_costs()
        ref = o.last_costs - o.term_const
        srt = np.sort(ref)
        if K > 1 and srt[1] - srt[0] > 4 * TOL["fp16"]["cost_rtol"] * abs(srt[0]):
            assert int(np.argmin(costs)) == int(np.argmin(ref))
        x0 = mlp_pred_batch(p, x0[None], u[None])[0]
    ctl.close()


def test_fp16_refuses_out_of_range_weights_and_auto_falls_back():
    """precision='fp16' needs weights inside the half range; 'auto' then takes bf16 (same kernel, bf16 operands)."""
    p = synthetic_mlp(4, 1, [32], seed=1)
    p.weights[0][0, 0] = 1.0e5
    cost = QuadCostParams(np.eye(4), np.eye(1), np.eye(4))
    with pytest.raises(ValueError, match="half range"):
        _engine(p, cost, [-1.0], [1.0], horizon=5, num_path=64, precision="fp16")
    ctl = _engine(p, cost, [-1.0], [1.0], horizon=5, num_path=64, precision="auto")
    assert ctl.precision == "bf16"
    ctl.close()


def test_tc_refuses_input_block_overflow():
    """nx padded to 32 plus nu >= 31 does not fit the 64-column input block: the tensor-core path must refuse it
    (round-1 advisor finding: it silently dropped the last control / bias columns) and 'auto' must take fp32."""
    p = synthetic_mlp(30, 31, [64], seed=1)
    cost = QuadCostParams(np.eye(30), np.eye(31), np.eye(30))
    for prec in ("bf16", "fp16"):
        with pytest.raises(ValueError, match="input block"):
            _engine(p, cost, -np.ones(31), np.ones(31), horizon=4, num_path=64, precision=prec)
    np.random.seed(0)
    ctl = _engine(p, cost, -np.ones(31), np.ones(31), horizon=4, num_path=64, precision="auto", noise="numpy")
    assert ctl.precision == "fp32"
    np.random.seed(0)
    o = MPPIOracle(p, cost, -np.ones(31), np.ones(31), horizon=4, num_path=64)
    _check_solve(ctl, o, np.zeros(30), o.sample_eps(), TOL["fp32"], check_argmin=False)
    ctl.close()


# Unmodified-reference fixtures on the TENSOR-CORE kernel (round-1 review: they only ran on the fp32 kernel).
# The fixtures use the trained cartpole MLP, whose dynamics near the upright are unstable, with F = diag(2,3000,.15,.3):
# operand rounding is amplified along the H-step rollouts and then by F, so the stated tolerance is per fixture
# (measured deviations: profiles/r02_precision.jsonl).  K512_H30 in bf16 is not comparable at all (the arg-min sample
# flips: cost differences of 2.5 % of the cost scale at lmda = 1) and is excluded with this note; C2's own
# configuration (K4096_H30) passes in both modes.
FIXTURE_TC_TOL = {
    ("mppi_cartpole_K256_H20", "fp16"): dict(act=1e-3, cost_diff=1e-3),
    ("mppi_cartpole_K256_H20", "bf16"): dict(act=5e-3, cost_diff=6e-3),
    ("mppi_cartpole_K100_H5", "fp16"): dict(act=5e-3, cost_diff=1e-3),
    ("mppi_cartpole_K100_H5", "bf16"): dict(act=5e-2, cost_diff=1.5e-2),
    ("mppi_cartpole_K512_H30", "fp16"): dict(act=5e-2, cost_diff=4e-3),
    ("mppi_cartpole_K4096_H30", "fp16"): dict(act=1e-3, cost_diff=2e-3),
    ("mppi_cartpole_K4096_H30", "bf16"): dict(act=1e-3, cost_diff=8e-3),
}


@pytest.mark.parametrize("name,precision", sorted(FIXTURE_TC_TOL))
def test_tensor_core_kernel_matches_unmodified_reference_fixture(name, precision):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    mlp, cost, umin, umax, _, _ = load_cartpole()
    tol = FIXTURE_TC_TOL[(name, precision)]
    np.random.seed(int(z["seed"]))
    ctl = _engine(mlp, cost, umin, umax, horizon=int(z["H"]), num_path=int(z["K"]), sigma=float(z["sigma"]),
                  lmda=float(z["lmda"]), noise="numpy", precision=precision)
    assert ctl.precision == precision
    np.testing.assert_allclose(ctl.act_sequence, z["act0"], rtol=0, atol=1e-7)
    constate = np.zeros(5)
    for s_ in range(int(z["n_steps"])):
        if s_ > 0:
            ctl.act_sequence = z["act_%d" % (s_ - 1)]      # every step compared from the reference's warm start
        u, constate = ctl.run(constate, z["x0_%d" % s_])
        costs, term = ctl.last_costs()
        ref = z["costs_%d" % s_]
        np.testing.assert_allclose(costs - costs.min(), ref - ref.min(), rtol=0, atol=tol["cost_diff"] * np.abs(ref).max())
        assert int(np.argmin(costs)) == int(z["argmin_%d" % s_])
        np.testing.assert_allclose(ctl.act_sequence, z["act_%d" % s_], rtol=0, atol=tol["act"])
        np.testing.assert_allclose(u, z["u_%d" % s_], rtol=0, atol=tol["act"] * 20.0)
    ctl.close()


# "dz" build of the tensor-core kernel (ReLU networks with >= 2 hidden layers, when the extra tf32 image fits in shared
# memory): the next step's input layer is fed by the output layer's fp32 accumulator through kind::tf32 MMAs plus a
# 16-bit part built from the previous state.  Both builds must meet the same tolerance on the same cases.
@pytest.mark.parametrize("dz", ["1", "0"])
@pytest.mark.parametrize("ci", [0, 1, 2, 7, 9])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_tensor_core_builds_with_and_without_dz(ci, dz, precision, monkeypatch):
    from autompc_b200 import _abi
    nx, nu, hidden, act, K, H, sigma, lmda, force_cg = TC_CASES[ci][:9]
    monkeypatch.setenv("AMPC_TC_DZ", dz)
    if force_cg:
        monkeypatch.setenv("AMPC_TC_FORCE_CG", force_cg)
    else:
        monkeypatch.delenv("AMPC_TC_FORCE_CG", raising=False)
    rng = np.random.default_rng(5)
    p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
    cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx), goal=0.05 * rng.normal(size=nx))
    umax = rng.uniform(0.5, 2.0, size=nu)
    umin = -umax * rng.uniform(0.5, 1.0, size=nu)
    np.random.seed(1)
    ctl = _engine(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda, noise="numpy", precision=precision)
    mode = _abi.lib().ampc_mppi_debug_tc_mode(ctl._h)
    assert mode != 0 and (not force_cg or (mode & 3) == int(force_cg))
    assert bool(mode & 16) == (dz == "1"), "dz build expected for a ReLU network with two or more hidden layers"
    np.random.seed(1)
    o = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, sigma=sigma, lmda=lmda)
    x0 = rng.normal(size=nx)
    for _ in range(3):
        eps = o.sample_eps()
        u, _ = _check_solve(ctl, o, x0, eps, TOL[precision], check_argmin=False)
        x0 = mlp_pred_batch(p, x0[None], u[None])[0]
    ctl.close()


def test_dz_build_not_taken_for_one_hidden_layer_or_other_activations():
    from autompc_b200 import _abi
    for ci in (3, 6, 8):                                  # tanh; one hidden layer (twice)
        nx, nu, hidden, act, K, H, sigma, lmda, _ = TC_CASES[ci][:9]
        p = synthetic_mlp(nx, nu, hidden, act=act, seed=3)
        cost = QuadCostParams(np.eye(nx), 0.01 * np.eye(nu), 10 * np.eye(nx))
        ctl = _engine(p, cost, -np.ones(nu), np.ones(nu), horizon=H, num_path=K, precision="fp16")
        mode = _abi.lib().ampc_mppi_debug_tc_mode(ctl._h)
        assert mode in (1, 2)
        ctl.solve(np.zeros(nx))
        ctl.close()

/*
 * ampc_b200.h -- C ABI of the B200-native MPC solve engine (libampc_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of williamedwards/autompc: the
 * sampling / shooting MPC solve.  The reference has no FFI today (it is 100 %
 * Python); each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every function returns AMPC_OK (0) or a negative ampc_status; the message
 *    for the calling thread's last failure is ampc_last_error();
 *  - "host" pointers are plain CPU memory owned by the caller, float64 and
 *    row-major exactly as the reference keeps them (NumPy / torch state_dict);
 *  - "dev" pointers are CUDA device memory on the handle's device, owned by the
 *    caller (e.g. torch.Tensor.data_ptr()); calls taking dev pointers are
 *    asynchronous on the given cudaStream_t (passed as void*), calls taking
 *    host pointers are synchronous;
 *  - a handle is not thread-safe; distinct handles are independent;
 *  - there is NO CPU fallback: without a CUDA device create() fails.
 */
#ifndef AMPC_B200_H
#define AMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ampc_status {
  AMPC_OK = 0,
  AMPC_ERR_INVALID = -1,     /* bad shape / argument            -> ValueError in the shim   */
  AMPC_ERR_UNSUPPORTED = -2, /* model / size outside the engine -> ValueError in the shim   */
  AMPC_ERR_CUDA = -3,        /* CUDA runtime failure            -> RuntimeError in the shim */
  AMPC_ERR_NOMEM = -4
} ampc_status;

/* activation codes: torch.nn.{ReLU,Tanh,Sigmoid,SELU}, autompc/sysid/mlp.py:44-53 */
enum { AMPC_ACT_RELU = 0, AMPC_ACT_TANH = 1, AMPC_ACT_SIGMOID = 2, AMPC_ACT_SELU = 3 };

/* arithmetic of the MPPI rollout kernel */
enum {
  AMPC_PREC_FP32 = 0, /* CUDA-core fp32 FMA, any layer sizes <= 256                       */
  AMPC_PREC_BF16 = 1, /* tcgen05 (UMMA) bf16 x bf16 -> fp32 in TMEM; state/cost stay fp32  */
  AMPC_PREC_FP16 = 2  /* tcgen05 (UMMA) IEEE half x half -> fp32: 11-bit significands, i.e. the
                         operand precision of kind::tf32, at the full kind::f16 rate and half the
                         shared-memory footprint of tf32 (a tf32 image of the 3x256 network does not
                         fit a CTA pair).  Operands saturate at +-65504; create() refuses weights
                         outside that range (AMPC_ERR_UNSUPPORTED).  The reference network is
                         float64 (autompc/sysid/mlp.py:165); this is the fp32-class tensor-core mode */
};
/* Both tensor-core modes run the "dz" build of the kernel for ReLU networks with two or more hidden layers when its
 * extra image fits in shared memory: the input layer of step i+1 is the 16-bit GEMM on [z_i | u_{i+1} | 1] plus a
 * kind::tf32 GEMM on the output layer's fp32 accumulator (the increment of the normalised state), see DESIGN.md 3.1.
 * Same stated tolerances; AMPC_TC_DZ=0 in the environment at create() keeps the build without it. */

#define AMPC_MAX_LAYERS 5 /* <= 4 hidden + output, autompc/sysid/mlp.py:110-111 */
#define AMPC_MAX_WIDTH 256 /* hidden_size upper bound, autompc/sysid/mlp.py:112-119 */

/* The learned dynamics: what autompc.sysid.mlp.MLP.get_parameters() holds
 * (autompc/sysid/mlp.py:308-313).  All pointers are HOST float64.            */
typedef struct ampc_mlp_desc {
  int32_t n_layers;        /* hidden layers + output layer                                   */
  const int32_t *dims;     /* n_layers+1 entries: [nx+nu, h_1, ..., h_L, nx]                 */
  const double *const *W;  /* per layer, (out,in) row-major like torch.nn.Linear.weight      */
  const double *const *b;  /* per layer, (out,)                                              */
  int32_t act;             /* AMPC_ACT_*                                                     */
  const double *xu_mean, *xu_std; /* (nx+nu,)  z-score of the input, mlp.py:20-24           */
  const double *dy_mean, *dy_std; /* (nx,)     un-z-score of the output, mlp.py:26-30       */
} ampc_mlp_desc;

/* Quadratic cost (autompc/costs/quad_cost.py:7-51) and control box
 * (autompc/tasks/task.py:257-267).  HOST float64, full matrices.              */
typedef struct ampc_quad_cost {
  const double *Q;    /* (nx,nx) */
  const double *R;    /* (nu,nu) */
  const double *F;    /* (nx,nx) */
  const double *goal; /* (nx,)   */
  const double *umin; /* (nu,)   */
  const double *umax; /* (nu,)   must be finite and > 0: the reference normalises by umax, mppi.py:102 */
  const double *goal_term; /* (nx,) goal of the terminal term, NULL = goal.  A reference SumCost of quadratics with
                              different goals (autompc/costs/sum_cost.py) folds to ONE quadratic per term type plus a
                              constant; the stage and terminal folds have different goals in general.            */
} ampc_quad_cost;

/* ------------------------------------------------------------------ MPPI --- */
/* Replaces autompc.control.mppi.MPPI (autompc/control/mppi.py:66-181).         */
typedef struct ampc_mppi_cfg {
  int32_t K;             /* num_path owned by THIS handle (the local shard)                  */
  int32_t H;             /* horizon (>= 2: mppi.py:123 indexes act_sequence[-2])             */
  int32_t nx, nu;
  double sigma;          /* noise VARIANCE (scale = sqrt(sigma), mppi.py:18)                  */
  double lmda;
  int32_t terminal_mode; /* 0 = reference: terminal cost of the LAST sample added to all
                                samples (mppi.py:79-82,148) -- cancels in the softmax;
                            1 = per-sample terminal cost (explicit non-reference option)      */
  int32_t precision;     /* AMPC_PREC_*                                                       */
  int32_t k_offset;      /* first global sample index of this shard (multi-GPU), else 0       */
  int32_t K_global;      /* total samples over all shards, else K                             */
  int32_t device;        /* CUDA device ordinal                                               */
} ampc_mppi_cfg;

typedef struct ampc_mppi ampc_mppi;

/* MPPI.__init__ (mppi.py:67-105) minus the random act_sequence draw, which
 * stays on the host (global NumPy stream) and is uploaded with set_act_seq.    */
int ampc_mppi_create(ampc_mppi **out, const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp,
                     const ampc_quad_cost *cost);
int ampc_mppi_destroy(ampc_mppi *h);

/* self.act_sequence (H,nu) float64, normalised controls (mppi.py:97-99).       */
int ampc_mppi_set_act_seq(ampc_mppi *h, const double *host_act);
int ampc_mppi_get_act_seq(ampc_mppi *h, double *host_act);

/* One MPPI.run (mppi.py:154-168) = shift (:122-123) + K rollouts (:126-150) +
 * update (:110-118), with HOST buffers: copies x0 in, u out, synchronises.
 *   host_eps : NULL -> in-kernel Philox4x32-10 noise keyed by (seed, counter,
 *              global sample, step);  else (H,K,nu) float64 UNCLIPPED noise
 *              already scaled by sqrt(sigma) (what mppi.py:126 draws) --
 *              the "external eps" parity mode.
 *   host_u   : (nu,) = act_sequence[0] * umax after the update.                */
int ampc_mppi_solve_host(ampc_mppi *h, const double *host_x0, const double *host_eps,
                         uint64_t seed, uint64_t counter, double *host_u);

/* Same solve with DEVICE float32 buffers, asynchronous on `stream`.
 *   dev_eps: NULL or (H,K,nu) float32.                                         */
int ampc_mppi_solve(ampc_mppi *h, const float *dev_x0, const float *dev_eps, uint64_t seed,
                    uint64_t counter, float *dev_u, void *stream);

/* Threshold stage costs -- autompc.costs.thresh_cost.ThresholdCost (autompc/costs/thresh_cost.py:8-38) and
 * BoxThresholdCost (:40-83), alone or as terms of a SumCost next to the quadratic (autompc/costs/sum_cost.py:49-81).
 * Term b adds weight[b] (1.0 in the reference) to a sample's cost at every horizon step whose PRE-step state x
 * (mppi.py:142) leaves the box: x[j] < lo[b][j] or x[j] > hi[b][j] for some j (+-inf = unbounded).
 * ThresholdCost(goal, obs_range, threshold) is the box goal[j] -+ threshold on j in obs_range.  No control or
 * terminal part (thresh_cost.py:33-38, :78-83).  lo, hi: (n_terms, nx) HOST float64; n_terms <= AMPC_MAX_BOX_TERMS;
 * n_terms = 0 removes them.  Call after create(), before the next solve.                                         */
#define AMPC_MAX_BOX_TERMS 8
int ampc_mppi_set_box_costs(ampc_mppi *h, int32_t n_terms, const double *lo, const double *hi, const double *weight);

/* Parity taps (synchronous).  costs: (K,) what do_rollouts returns (mppi.py:152)
 * WITHOUT the terminal_mode-0 scalar, which is returned separately.            */
int ampc_mppi_get_costs(ampc_mppi *h, double *host_costs, double *term_const);
/* The exact unclipped noise the Philox path uses for (seed, counter): (H,K,nu). */
int ampc_mppi_get_noise(ampc_mppi *h, uint64_t seed, uint64_t counter, float *host_eps);

/* Multi-GPU (samples sharded over ranks).  rollout_partial
 * writes this shard's softmax partial record
 *     [ min cost m, sum_k exp(-(c_k-m)/lmda), sum_k exp(-(c_k-m)/lmda) * eps_k (H*nu) ]
 * (ampc_mppi_record_floats() float32) to dev_record; after the ranks exchange
 * records (one all-gather), merge applies mppi.py:115-118 on every rank.       */
int ampc_mppi_record_floats(const ampc_mppi *h);
int ampc_mppi_rollout_partial(ampc_mppi *h, const float *dev_x0, const float *dev_eps,
                              uint64_t seed, uint64_t counter, float *dev_record, void *stream);
int ampc_mppi_merge(ampc_mppi *h, const float *dev_records, int32_t n_records, float *dev_u,
                    void *stream);

/* Multi-GPU fast path: the exchange above fused into the rollout kernel's tail over NVLink peer memory.
 * Each rank allocates a mailbox (2 slots x world records + flags) and exports it (CUDA IPC, 64-byte handle);
 * after the handles of all ranks are gathered (any transport) connect_peers_ipc maps them.  solve_fused then is
 * ONE launch per solve: rollouts -> shard record -> stores into every rank's mailbox + system-scope flag ->
 * wait for the world's flags -> merge in rank order -> update (mppi.py:115-118).  connect_peers_local is the
 * same for handles living in one process.                                                                */
int ampc_mppi_mailbox_ipc(ampc_mppi *h, int32_t world, void *ipc_handle_64);
int ampc_mppi_connect_peers_ipc(ampc_mppi *h, int32_t world, int32_t rank, const void *ipc_handles);
int ampc_mppi_connect_peers_local(ampc_mppi *h, int32_t world, int32_t rank, ampc_mppi *const *handles);
int ampc_mppi_solve_fused(ampc_mppi *h, const float *dev_x0, const float *dev_eps, uint64_t seed,
                          uint64_t counter, float *dev_u, void *stream);
/* solve_fused with HOST buffers (what MPPI.run holds, mppi.py:154-168): like solve_host the observation rides in
 * the kernel parameters and the merged control lands in mapped pinned host memory; synchronous.  host_eps: NULL or
 * this shard's (H,K,nu) float64 slice of the unclipped noise.                                              */
int ampc_mppi_solve_fused_host(ampc_mppi *h, const double *host_x0, const double *host_eps, uint64_t seed,
                               uint64_t counter, double *host_u);

/* Debug tap, no reference counterpart: when the handle was created with AMPC_TC_TRACE=1 in the environment,
 * copies the tcgen05 kernel's timeline of CTA 0 ([warp][64] words = clock64 << 8 | tag) to host.       */
int ampc_mppi_debug_trace(ampc_mppi *h, unsigned long long *host, int32_t max_words);
/* Debug tap: 0 = the handle does not run the tcgen05 kernel; else cta_group (1 | 2) + 16 for the "dz" build. */
int ampc_mppi_debug_tc_mode(ampc_mppi *h);

/* ------------------------------------------------------- MLP model ops --- */
/* Replaces autompc.sysid.mlp.MLP.pred / pred_batch (mlp.py:219-236) and
 * pred_diff / pred_diff_batch (mlp.py:238-305).  float64 on the device (the
 * reference network is .double(), mlp.py:165).  HOST buffers, synchronous.     */
typedef struct ampc_mlp ampc_mlp;
int ampc_mlp_create(ampc_mlp **out, const ampc_mlp_desc *mlp, int32_t nx, int32_t nu, int32_t device);
int ampc_mlp_destroy(ampc_mlp *m);
/* X (m,nx), U (m,nu) -> Xn (m,nx) */
int ampc_mlp_pred_batch(ampc_mlp *m, int32_t batch, const double *X, const double *U, double *Xn);
/* k-step open-loop prediction, the inner loop of get_model_rmse (autompc/evaluation/model_metrics.py:33-35):
 * X0 (m,nx) window starts, U (horizon,m,nu) recorded controls -> Xh (m,nx) = pred_batch applied `horizon` times. */
int ampc_mlp_rollout_batch(ampc_mlp *m, int32_t batch, int32_t horizon, const double *X0, const double *U, double *Xh);
/* + Jx (m,nx,nx) = d Xn / d X, Ju (m,nx,nu) = d Xn / d U */
int ampc_mlp_pred_diff_batch(ampc_mlp *m, int32_t batch, const double *X, const double *U,
                             double *Xn, double *Jx, double *Ju);

/* Direct-transcription callbacks of autompc.control.nmpc.NonLinearMPCProblem for an MLP model (the NLP solver, IPOPT,
 * stays on the host).  x: the decision vector [ states (H+1, nx) | ctrls (H, nu) ] (nmpc.py:56-66).
 *   constraint: c (H*nx)  = get_constraint(x)  (nmpc.py:102-110):  c[i] = -state[i+1] + pred(state[i], ctrl[i]);
 *   jacobian:   jac (H*(nx*nx + nx*nu + nx)) = get_jacobian(x, False) (nmpc.py:170-187): per knot the dense state
 *               Jacobian (row-major), the dense control Jacobian, then -1 for state[i+1]; the (row, col) pattern of
 *               get_jacobian(x, True) (nmpc.py:148-169) is index arithmetic and stays on the host.                */
int ampc_mlp_nmpc_constraint(ampc_mlp *m, int32_t H, const double *x, double *c);
int ampc_mlp_nmpc_jacobian(ampc_mlp *m, int32_t H, const double *x, double *jac);
/* Measurement tap: device time (CUDA events) of the kernel the last call above launched on this handle, without the
 * host <-> device copies around it.  No reference counterpart. */
int ampc_mlp_debug_last_kernel_ms(ampc_mlp *m, float *ms);

/* ------------------------------------------------------- linear models --- */
/* Replaces ARX.pred / pred_batch (autompc/sysid/arx.py:146-154) and Koopman.pred / pred_batch
 * (autompc/sysid/koopman.py:165-173):  Xn = (A @ X.T + B @ U.T).T,  A (ns,ns), B (ns,nu) row-major HOST float64, ns =
 * the model's state_dim (history stack / lifted observation).  float64 on the device, HOST buffers, synchronous.    */
typedef struct ampc_linear ampc_linear;
int ampc_linear_create(ampc_linear **out, int32_t ns, int32_t nu, const double *A, const double *B, int32_t device);
int ampc_linear_destroy(ampc_linear *h);
int ampc_linear_pred_batch(ampc_linear *h, int32_t batch, const double *X, const double *U, double *Xn);

/* ------------------------------------------------- device-resident closed loop --- */
/* Replaces autompc.utils.simulation.simulate (autompc/utils/simulation.py:11-64) for an MPPI controller and an
 * MLP simulation model when term_cond is None:  T x [ u = controller.run(x) ; x = sim_model.pred(x, u) ]
 * (mppi.py:154-168, sysid/mlp.py:219-227) with no host round trip, plus the trajectory cost of Cost.__call__
 * (autompc/costs/cost.py:27-41) for the controller's QuadCost.  start() enqueues the T solves and plant steps on
 * the handle's own stream and returns; finish() waits and copies obs (T+1,nx), ctrl (T,nu), cost (1) to host
 * float64 buffers (any may be NULL).  Handles are independent: many closed loops can be in flight at once
 * (the tuner's candidate evaluations, tuning/pipeline_tuner.py:213-239).  Noise: in-kernel Philox, solve
 * counters counter0 .. counter0+T-1.                                                                        */
int ampc_mppi_closed_loop_start(ampc_mppi *h, ampc_mlp *sim, const double *x0, int32_t T, uint64_t seed,
                                uint64_t counter0);
/* The cost the closed loop accumulates over the trajectory (Cost.__call__, cost.py:27-41) defaults to the
 * controller's own.  The tuner scores a candidate with the TASK's cost instead, which differs from the cost the
 * controller optimises (tuning/pipeline_tuner.py:230-231; e.g. the cartpole benchmark's ThresholdCost,
 * benchmarks/cartpole.py:38-60): set_eval_cost replaces it by quad (NULL = no quadratic part; umin/umax ignored)
 * plus n_terms box terms as in ampc_mppi_set_box_costs.                                                       */
int ampc_mppi_set_eval_cost(ampc_mppi *h, const ampc_quad_cost *quad, int32_t n_terms, const double *lo,
                            const double *hi, const double *weight);
int ampc_mppi_closed_loop_finish(ampc_mppi *h, int32_t T, double *obs_out, double *ctrl_out, double *cost_out);

/* ------------------------------------------------------------------ iLQR --- */
/* Replaces autompc.control.ilqr.IterativeLQR.compute_ilqr_default
 * (autompc/control/ilqr.py:100-265) for MLP dynamics + QuadCost: the whole
 * solve (init rollout with Jacobians, up to max_iter x [backward Riccati,
 * 10-alpha line search, Jacobian refresh]) is one kernel launch, float64.      */
typedef struct ampc_ilqr_cfg {
  int32_t H, nx, nu;
  double dt;
  int32_t bounded;         /* clip controls to [umin,umax] (ilqr.py:203-204)   */
  int32_t max_iter;        /* 50  */
  int32_t ls_max_iter;     /* 10 (at most 20: one or two warps per line-search rollout) */
  double ls_discount;      /* 0.2 */
  double ls_cost_threshold;/* 0.3 */
  double u_threshold;      /* 1e-3 */
  int32_t device;
} ampc_ilqr_cfg;
typedef struct ampc_ilqr ampc_ilqr;
int ampc_ilqr_create(ampc_ilqr **out, const ampc_ilqr_cfg *cfg, const ampc_mlp_desc *mlp,
                     const ampc_quad_cost *cost);
int ampc_ilqr_destroy(ampc_ilqr *h);
/* host buffers: x0 (nx,), uguess (H,nu) or NULL (= zeros, ilqr.py:282);
 * out: states (H+1,nx), ctrls (H,nu), Ks (H,nu,nx), ks (H,nu),
 *      info[0]=converged, info[1]=iterations run, info[2]=line-search-failed,
 *      alpha_idx (max_iter,) adopted line-search index per iteration (-1 past the end). */
int ampc_ilqr_solve_host(ampc_ilqr *h, const double *x0, const double *uguess, double *states,
                         double *ctrls, double *Ks, double *ks, int32_t *info, int32_t *alpha_idx);
/* The same solve again from the x0 the last solve_host call uploaded (uguess = zeros), asynchronous on `stream`, no
 * host copies; results stay on the device.  Measurement hook: the kernel alone under CUDA events.               */
int ampc_ilqr_launch(ampc_ilqr *h, void *stream);
/* Debug tap, no reference counterpart: SM cycles of the last solve per phase: [0] set-up + initial rollout, [1] backward
 * passes, [2] line-search rollouts, [3] objectives + acceptance + bookkeeping, [4] Jacobian refreshes, [5] copy-out,
 * [6] total, [7] iterations run.                                                                                */
int ampc_ilqr_debug_profile(ampc_ilqr *h, unsigned long long *out8);

/* ------------------------------------------------------------------ misc --- */
const char *ampc_last_error(void);
const char *ampc_version(void);
/* number of kernels this library has launched in this process (bench evidence) */
uint64_t ampc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AMPC_B200_H */

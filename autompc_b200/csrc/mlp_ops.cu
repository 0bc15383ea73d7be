// float64 MLP inference and batched Jacobians on the device.
//
// Replaces autompc.sysid.mlp.MLP.pred/pred_batch (autompc/sysid/mlp.py:219-236) and
// pred_diff/pred_diff_batch (:238-305).  The reference network is float64
// (mlp.py:165); these kernels keep float64 so results agree to rounding.  The
// Jacobian is propagated forward through the layer stack,
//   J_0 = W_0 diag(1/xu_std),  J_l = W_l (act'(pre_{l-1}) * J_{l-1}),  J = diag(dy_std) J_L,
// which is the closed form of the reference's eye(nx) back-propagation.
//
// One CTA per sample (eight samples per CTA for pred_batch / k-step rollouts on batches
// beyond one wave, ampc_mlp_blocked); activations and the two Jacobian panels live in shared
// memory; weights are stored in-major ([in][out]) so neighbouring threads read
// neighbouring outputs.  The same device routines are reused by the iLQR kernel.
#include <cstdlib>
#include <vector>

#include "ampc_common.cuh"
#include "mlp_f64.cuh"

struct ampc_mlp {
  AmpcMlpF64 net;
  int device = 0;
  double *d_blob = nullptr;
  size_t smem_pred = 0, smem_batch = 0, smem_diff = 0;
  double *d_scratch = nullptr;       // grow-only staging for the host-buffer entry points (no cudaMalloc per call)
  size_t scratch_doubles = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // around the last launch (ampc_mlp_debug_last_kernel_ms)
};

namespace {

constexpr int NT = 128;
constexpr int NTJ = 512;   // Jacobian kernels: a [width x (nx+nu)] panel per layer is 5 888 dot products at 3x256

// threshold terms of a trajectory cost (thresh_cost.py:27-32, :73-77): n_box x [lo (nx) | hi (nx) | weight]
__device__ __forceinline__ double ampc_box_cost_f64(const double *box, int n_box, int nx, const double *x) {
  double c = 0.0;
  for (int b = 0; b < n_box; ++b) {
    const double *bx = box + (size_t)b * (2 * nx + 1);
    bool out = false;
    for (int j = 0; j < nx; ++j) out = out || (x[j] < bx[j]) || (x[j] > bx[nx + j]);
    if (out) c += bx[2 * nx];
  }
  return c;
}

// One sample per CTA: the batches of the reference's callers (1, H, a few hundred) leave SMs idle anyway.
__global__ void __launch_bounds__(NT) pred_batch_kernel(const AmpcMlpF64 net, int batch, const double *X,
                                                        const double *U, double *Xn) {
  extern __shared__ double sm_d[];
  const int s = blockIdx.x;
  if (s >= batch) return;
  double *h0 = sm_d, *h1 = sm_d + net.max_width;
  const int nx = net.nx, nu = net.nu;
  for (int j = threadIdx.x; j < nx + nu; j += NT) {
    const double v = j < nx ? X[(size_t)s * nx + j] : U[(size_t)s * nu + (j - nx)];
    h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
  }
  __syncthreads();
  const double *out = ampc_mlp_f64_forward(net, h0, h1, nullptr, threadIdx.x, NT);
  for (int j = threadIdx.x; j < nx; j += NT)
    Xn[(size_t)s * nx + j] = X[(size_t)s * nx + j] + (out[j] * net.dy_std[j] + net.dy_mean[j]);
}

// Batches beyond one wave of one-sample CTAs (ampc_mlp_blocked): a CTA owns AMPC_MLP_SB consecutive samples
// (mlp_f64.cuh: ampc_mlp_f64_forward_blocked), so the 1.1 MB of float64 weights is read from L2 once per eight samples
// instead of once per sample; same arithmetic per sample, bit-identical results (4.57 -> 1.91 ms at batch 65536, C3 net).
__global__ void __launch_bounds__(NT) pred_batch_blocked_kernel(const AmpcMlpF64 net, int batch, const double *X,
                                                                const double *U, double *Xn) {
  extern __shared__ double sm_d[];
  constexpr int SB = AMPC_MLP_SB;
  const int s0 = blockIdx.x * SB;
  const int nx = net.nx, nu = net.nu, nin = nx + nu, W = net.max_width;
  double *h0 = sm_d, *h1 = sm_d + SB * W;
  for (int t = threadIdx.x; t < SB * nin; t += NT) {
    const int q = t / nin, j = t - q * nin, s = s0 + q;
    double z = 0.0;
    if (s < batch) {
      const double v = j < nx ? X[(size_t)s * nx + j] : U[(size_t)s * nu + (j - nx)];
      z = (v - net.xu_mean[j]) / net.xu_std[j];
    }
    h0[q * W + j] = z;
  }
  __syncthreads();
  const double *out = ampc_mlp_f64_forward_blocked(net, h0, h1, W, threadIdx.x, NT);
  for (int t = threadIdx.x; t < SB * nx; t += NT) {
    const int q = t / nx, j = t - q * nx, s = s0 + q;
    if (s < batch) Xn[(size_t)s * nx + j] = X[(size_t)s * nx + j] + (out[q * W + j] * net.dy_std[j] + net.dy_mean[j]);
  }
}

__global__ void __launch_bounds__(NTJ) pred_diff_kernel(const AmpcMlpF64 net, int batch, const double *X,
                                                       const double *U, double *Xn, double *Jx, double *Ju) {
  extern __shared__ double sm_d[];
  const int s = blockIdx.x;
  if (s >= batch) return;
  const int nx = net.nx, nu = net.nu, nin = nx + nu;
  double *h0 = sm_d, *h1 = h0 + net.max_width, *g = h1 + net.max_width;
  double *J0 = g + net.max_width, *J1 = J0 + (size_t)net.max_width * nin;
  for (int j = threadIdx.x; j < nin; j += NTJ) {
    const double v = j < nx ? X[(size_t)s * nx + j] : U[(size_t)s * nu + (j - nx)];
    h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
  }
  __syncthreads();
  const double *J;
  const double *out = ampc_mlp_f64_forward_jac(net, h0, h1, g, J0, J1, &J, threadIdx.x, NTJ);
  for (int j = threadIdx.x; j < nx; j += NTJ)
    Xn[(size_t)s * nx + j] = X[(size_t)s * nx + j] + (out[j] * net.dy_std[j] + net.dy_mean[j]);
  for (int t = threadIdx.x; t < nx * nin; t += NTJ) {
    const int r = t / nin, c = t - r * nin;
    const double v = J[t] * net.dy_std[r];                     // mlp.py:298
    if (c < nx) Jx[((size_t)s * nx + r) * nx + c] = v + (r == c ? 1.0 : 0.0);   // mlp.py:303
    else Ju[((size_t)s * nx + r) * nu + (c - nx)] = v;
  }
}

// Direct-transcription callbacks (autompc/control/nmpc.py:102-110, :148-187) for an MLP model, one CTA per knot i.
// Decision vector x = [ states (H+1, nx) | ctrls (H, nu) ].  WANT_JAC = false: c[i] = -state[i+1] + pred(state[i], ctrl[i])
// (get_constraint).  WANT_JAC = true: the values of get_jacobian(x, False) in the reference's order, per knot
// [ d pred / d state (nx*nx, row-major) | d pred / d ctrl (nx*nu) | -1 x nx ].
template <bool WANT_JAC>
__global__ void __launch_bounds__(NTJ) nmpc_knot_kernel(const AmpcMlpF64 net, int H, const double *xvec, double *out) {
  extern __shared__ double sm_d[];
  const int i = blockIdx.x;
  if (i >= H) return;
  const int nx = net.nx, nu = net.nu, nin = nx + nu;
  const double *xs = xvec + (size_t)i * nx, *xn = xvec + (size_t)(i + 1) * nx;
  const double *us = xvec + (size_t)(H + 1) * nx + (size_t)i * nu;
  double *h0 = sm_d, *h1 = h0 + net.max_width, *g = h1 + net.max_width;
  for (int j = threadIdx.x; j < nin; j += NTJ) {
    const double v = j < nx ? xs[j] : us[j - nx];
    h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
  }
  __syncthreads();
  if constexpr (!WANT_JAC) {
    const double *o = ampc_mlp_f64_forward(net, h0, h1, nullptr, threadIdx.x, NTJ);
    for (int j = threadIdx.x; j < nx; j += NTJ)
      out[(size_t)i * nx + j] = -xn[j] + (xs[j] + (o[j] * net.dy_std[j] + net.dy_mean[j]));     // nmpc.py:109
  } else {
    double *J0 = g + net.max_width, *J1 = J0 + (size_t)net.max_width * nin;
    const double *J;
    ampc_mlp_f64_forward_jac(net, h0, h1, g, J0, J1, &J, threadIdx.x, NTJ);
    double *o = out + (size_t)i * (nx * nx + nx * nu + nx);
    for (int t = threadIdx.x; t < nx * nin; t += NTJ) {
      const int r = t / nin, c = t - r * nin;
      const double v = J[t] * net.dy_std[r];                                                  // mlp.py:298
      if (c < nx) o[r * nx + c] = v + (r == c ? 1.0 : 0.0);                                   // nmpc.py:180-181
      else o[nx * nx + r * nu + (c - nx)] = v;                                                // nmpc.py:182-183
    }
    for (int j = threadIdx.x; j < nx; j += NTJ) o[nx * nx + nx * nu + j] = -1.0;               // nmpc.py:184-185
  }
}

// k-step open-loop prediction: window s starts at X0[s] and is advanced `horizon` times with the recorded controls
// U[k][s] -- the inner loop of get_model_rmse (autompc/evaluation/model_metrics.py:33-35) as one launch.  Each step
// is pred_batch_kernel's arithmetic, so the result equals `horizon` chained pred_batch calls bit for bit.
__global__ void __launch_bounds__(NT) rollout_batch_kernel(const AmpcMlpF64 net, int batch, int horizon, const double *X0,
                                                           const double *U, double *Xh) {
  extern __shared__ double sm_d[];
  const int s = blockIdx.x;
  if (s >= batch) return;
  const int nx = net.nx, nu = net.nu;
  double *h0 = sm_d, *h1 = sm_d + net.max_width, *x = h1 + net.max_width;
  for (int j = threadIdx.x; j < nx; j += NT) x[j] = X0[(size_t)s * nx + j];
  __syncthreads();
  for (int k = 0; k < horizon; ++k) {
    const double *u = U + ((size_t)k * batch + s) * nu;
    for (int j = threadIdx.x; j < nx + nu; j += NT) {
      const double v = j < nx ? x[j] : u[j - nx];
      h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
    }
    __syncthreads();
    const double *out = ampc_mlp_f64_forward(net, h0, h1, nullptr, threadIdx.x, NT);
    for (int j = threadIdx.x; j < nx; j += NT) x[j] = x[j] + (out[j] * net.dy_std[j] + net.dy_mean[j]);
    __syncthreads();
  }
  for (int j = threadIdx.x; j < nx; j += NT) Xh[(size_t)s * nx + j] = x[j];
}

// sample-blocked form (see pred_batch_blocked_kernel): 13.4 -> 7.1 ms at batch 8192, horizon 20
__global__ void __launch_bounds__(NT) rollout_batch_blocked_kernel(const AmpcMlpF64 net, int batch, int horizon,
                                                                   const double *X0, const double *U, double *Xh) {
  extern __shared__ double sm_d[];
  constexpr int SB = AMPC_MLP_SB;
  const int s0 = blockIdx.x * SB;
  const int nx = net.nx, nu = net.nu, nin = nx + nu, W = net.max_width;
  double *h0 = sm_d, *h1 = h0 + SB * W, *x = h1 + SB * W;   // x: [SB][nx]
  for (int t = threadIdx.x; t < SB * nx; t += NT) {
    const int q = t / nx, j = t - q * nx, s = s0 + q;
    x[t] = s < batch ? X0[(size_t)s * nx + j] : 0.0;
  }
  __syncthreads();
  for (int k = 0; k < horizon; ++k) {
    for (int t = threadIdx.x; t < SB * nin; t += NT) {
      const int q = t / nin, j = t - q * nin, s = s0 + q;
      double z = 0.0;
      if (s < batch) {
        const double v = j < nx ? x[q * nx + j] : U[((size_t)k * batch + s) * nu + (j - nx)];
        z = (v - net.xu_mean[j]) / net.xu_std[j];
      }
      h0[q * W + j] = z;
    }
    __syncthreads();
    const double *out = ampc_mlp_f64_forward_blocked(net, h0, h1, W, threadIdx.x, NT);
    for (int t = threadIdx.x; t < SB * nx; t += NT) {
      const int q = t / nx, j = t - q * nx;
      x[t] = x[t] + (out[q * W + j] * net.dy_std[j] + net.dy_mean[j]);
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < SB * nx; t += NT) {
    const int q = t / nx, j = t - q * nx, s = s0 + q;
    if (s < batch) Xh[(size_t)s * nx + j] = x[t];
  }
}

// One closed-loop plant step on the device: x <- sim.pred(x, u) (mlp.py:219-227, float64), plus the float32 copy the
// next solve reads, the trajectory record and the running trajectory cost of Cost.__call__ (cost.py:27-41).
__global__ void __launch_bounds__(NT) sim_step_kernel(const AmpcMlpF64 net, double *x, const float *u, float *x32,
                                                      double *obs_next, double *ctrl_t, const double *Q, const double *R,
                                                      const double *goal, const double *box, int n_box, double *cost) {
  extern __shared__ double sm_d[];
  double *h0 = sm_d, *h1 = sm_d + net.max_width;
  const int nx = net.nx, nu = net.nu;
  for (int j = threadIdx.x; j < nx + nu; j += NT) {
    const double v = j < nx ? x[j] : (double)u[j - nx];
    h0[j] = (v - net.xu_mean[j]) / net.xu_std[j];
  }
  if (cost != nullptr && threadIdx.x == 0) {             // stage cost of (x_t, u_t): obst^T Q obst + u^T R u
    double c = 0.0;
    for (int j = 0; j < nx; ++j) {
      double col = 0.0;
      for (int i = 0; i < nx; ++i) col += (x[i] - goal[i]) * Q[i * nx + j];
      c += col * (x[j] - goal[j]);
    }
    for (int j = 0; j < nu; ++j) {
      double col = 0.0;
      for (int i = 0; i < nu; ++i) col += (double)u[i] * R[i * nu + j];
      c += col * (double)u[j];
    }
    *cost += c + ampc_box_cost_f64(box, n_box, nx, x);
  }
  __syncthreads();
  const double *out = ampc_mlp_f64_forward(net, h0, h1, nullptr, threadIdx.x, NT);
  for (int j = threadIdx.x; j < nx; j += NT) {
    const double xn = x[j] + (out[j] * net.dy_std[j] + net.dy_mean[j]);
    obs_next[j] = xn;
    x32[j] = (float)xn;
  }
  for (int j = threadIdx.x; j < nu; j += NT) ctrl_t[j] = (double)u[j];
  __syncthreads();
  for (int j = threadIdx.x; j < nx; j += NT) x[j] = obs_next[j];
}

// terminal part of Cost.__call__: obs cost of the last state (its control is zero) + terminal cost
__global__ void traj_cost_final_kernel(int nx, const double *x, const double *Q, const double *F, const double *goal,
                                       const double *goalF, const double *box, int n_box, double *cost) {
  if (threadIdx.x != 0) return;
  double c = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    const double *M = pass == 0 ? Q : F;
    const double *goal_p = pass == 0 ? goal : goalF;
    for (int j = 0; j < nx; ++j) {
      double col = 0.0;
      for (int i = 0; i < nx; ++i) col += (x[i] - goal_p[i]) * M[i * nx + j];
      c += col * (x[j] - goal_p[j]);
    }
  }
  *cost += c + ampc_box_cost_f64(box, n_box, nx, x);   // the last state's stage cost includes the threshold terms
}

}  // namespace

int ampc_mlp_device(const ampc_mlp *m) { return m->device; }
int ampc_mlp_nx(const ampc_mlp *m) { return m->net.nx; }
int ampc_mlp_nu(const ampc_mlp *m) { return m->net.nu; }

int ampc_mlp_sim_step_launch(ampc_mlp *m, double *d_x, const float *d_u, float *d_x32, double *d_obs_next, double *d_ctrl_t,
                             const double *d_Q, const double *d_R, const double *d_goal, const double *d_box, int n_box,
                             double *d_cost, cudaStream_t s) {
  sim_step_kernel<<<1, NT, m->smem_pred, s>>>(m->net, d_x, d_u, d_x32, d_obs_next, d_ctrl_t, d_Q, d_R, d_goal, d_box, n_box,
                                              d_cost);
  ampc_count_launch();
  AMPC_CUDA_CHECK(cudaGetLastError());
  return AMPC_OK;
}

int ampc_traj_cost_final_launch(int nx, const double *d_x, const double *d_Q, const double *d_F, const double *d_goal,
                                const double *d_goalF, const double *d_box, int n_box, double *d_cost, cudaStream_t s) {
  traj_cost_final_kernel<<<1, 32, 0, s>>>(nx, d_x, d_Q, d_F, d_goal, d_goalF, d_box, n_box, d_cost);
  ampc_count_launch();
  AMPC_CUDA_CHECK(cudaGetLastError());
  return AMPC_OK;
}

int ampc_mlp_f64_upload(const ampc_mlp_desc *mlp, int nx, int nu, AmpcMlpF64 *net, double **blob_out) {
  AMPC_REQUIRE(mlp && mlp->n_layers >= 2 && mlp->n_layers <= AMPC_MAX_LAYERS, AMPC_ERR_UNSUPPORTED,
               "MLP must have 1..%d hidden layers", AMPC_MAX_LAYERS - 1);
  AMPC_REQUIRE(mlp->dims[0] == nx + nu && mlp->dims[mlp->n_layers] == nx, AMPC_ERR_INVALID,
               "MLP dims do not match nx+nu -> nx");
  AMPC_REQUIRE(mlp->act >= 0 && mlp->act <= 3, AMPC_ERR_UNSUPPORTED, "unknown activation %d", mlp->act);
  memset(net, 0, sizeof(*net));
  net->n_layers = mlp->n_layers;
  net->act = mlp->act;
  net->nx = nx;
  net->nu = nu;
  size_t off = 0;
  size_t woff[AMPC_MAX_LAYERS], boff[AMPC_MAX_LAYERS];
  for (int l = 0; l <= mlp->n_layers; ++l) {
    net->dims[l] = mlp->dims[l];
    AMPC_REQUIRE(mlp->dims[l] >= 1 && mlp->dims[l] <= AMPC_MAX_WIDTH, AMPC_ERR_UNSUPPORTED, "layer width %d",
                 mlp->dims[l]);
    if (mlp->dims[l] > net->max_width) net->max_width = mlp->dims[l];
  }
  for (int l = 0; l < mlp->n_layers; ++l) {
    woff[l] = off; off += (size_t)mlp->dims[l] * mlp->dims[l + 1];
    boff[l] = off; off += mlp->dims[l + 1];
  }
  const size_t o_xm = off; off += nx + nu;
  const size_t o_xs = off; off += nx + nu;
  const size_t o_dm = off; off += nx;
  const size_t o_ds = off; off += nx;
  std::vector<double> hb(off);
  for (int l = 0; l < mlp->n_layers; ++l) {
    const int Kin = mlp->dims[l], N = mlp->dims[l + 1];
    for (int k = 0; k < Kin; ++k)
      for (int j = 0; j < N; ++j) hb[woff[l] + (size_t)k * N + j] = mlp->W[l][(size_t)j * Kin + k];
    for (int j = 0; j < N; ++j) hb[boff[l] + j] = mlp->b[l][j];
  }
  for (int j = 0; j < nx + nu; ++j) {
    AMPC_REQUIRE(mlp->xu_std[j] != 0.0, AMPC_ERR_INVALID, "xu_std[%d] == 0", j);
    hb[o_xm + j] = mlp->xu_mean[j];
    hb[o_xs + j] = mlp->xu_std[j];
  }
  for (int j = 0; j < nx; ++j) { hb[o_dm + j] = mlp->dy_mean[j]; hb[o_ds + j] = mlp->dy_std[j]; }
  double *d = nullptr;
  AMPC_CUDA_CHECK(cudaMalloc(&d, off * sizeof(double)));
  cudaError_t e = cudaMemcpy(d, hb.data(), off * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); AMPC_CUDA_CHECK(e); }
  for (int l = 0; l < mlp->n_layers; ++l) { net->Wt[l] = d + woff[l]; net->b[l] = d + boff[l]; }
  net->xu_mean = d + o_xm; net->xu_std = d + o_xs; net->dy_mean = d + o_dm; net->dy_std = d + o_ds;
  *blob_out = d;
  return AMPC_OK;
}

extern "C" int ampc_mlp_create(ampc_mlp **out, const ampc_mlp_desc *mlp, int32_t nx, int32_t nu, int32_t device) {
  AMPC_REQUIRE(out, AMPC_ERR_INVALID, "null out");
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  AMPC_REQUIRE(ce == cudaSuccess && ndev > 0, AMPC_ERR_CUDA, "no CUDA device: libampc_b200 has no CPU fallback (%s)",
               cudaGetErrorString(ce));
  AMPC_REQUIRE(device >= 0 && device < ndev, AMPC_ERR_INVALID, "device %d of %d", device, ndev);
  AMPC_CUDA_CHECK(cudaSetDevice(device));
  ampc_mlp *m = new ampc_mlp();
  m->device = device;
  int rc = ampc_mlp_f64_upload(mlp, nx, nu, &m->net, &m->d_blob);
  if (rc) { delete m; return rc; }
  m->smem_pred = 2 * (size_t)m->net.max_width * sizeof(double);                     // one sample (sim_step_kernel)
  m->smem_batch = 2 * (size_t)AMPC_MLP_SB * m->net.max_width * sizeof(double);      // AMPC_MLP_SB samples per CTA
  m->smem_diff = (3 * (size_t)m->net.max_width + 2 * (size_t)m->net.max_width * (nx + nu)) * sizeof(double);
  cudaError_t e = ampc_raise_smem_limit((const void *)pred_diff_kernel, m->smem_diff);
  if (e == cudaSuccess) e = ampc_raise_smem_limit((const void *)nmpc_knot_kernel<true>, m->smem_diff);
  if (e == cudaSuccess) e = ampc_raise_smem_limit((const void *)nmpc_knot_kernel<false>, m->smem_diff);
  if (e != cudaSuccess) {
    ampc_set_error("pred_diff kernel needs %zu B shared memory: %s", m->smem_diff, cudaGetErrorString(e));
    cudaFree(m->d_blob);
    delete m;
    return AMPC_ERR_UNSUPPORTED;
  }
  *out = m;
  return AMPC_OK;
}

extern "C" int ampc_mlp_destroy(ampc_mlp *m) {
  if (!m) return AMPC_OK;
  cudaSetDevice(m->device);
  cudaFree(m->d_blob);
  cudaFree(m->d_scratch);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  delete m;
  return AMPC_OK;
}

// Device staging of `n` doubles on the handle; grown (never shrunk) when a call needs more.
static int mlp_scratch(ampc_mlp *m, size_t n, double **out) {
  if (n > m->scratch_doubles) {
    cudaFree(m->d_scratch);
    m->d_scratch = nullptr;
    m->scratch_doubles = 0;
    AMPC_CUDA_CHECK(cudaMalloc(&m->d_scratch, n * sizeof(double)));
    m->scratch_doubles = n;
  }
  if (!m->ev0) {
    AMPC_CUDA_CHECK(cudaEventCreate(&m->ev0));
    AMPC_CUDA_CHECK(cudaEventCreate(&m->ev1));
  }
  *out = m->d_scratch;
  return AMPC_OK;
}

extern "C" int ampc_mlp_debug_last_kernel_ms(ampc_mlp *m, float *ms) {
  AMPC_REQUIRE(m && ms && m->ev0, AMPC_ERR_INVALID, "no launch recorded on this handle");
  AMPC_CUDA_CHECK(cudaSetDevice(m->device));
  AMPC_CUDA_CHECK(cudaEventElapsedTime(ms, m->ev0, m->ev1));
  return AMPC_OK;
}

// Sample-blocked kernels once the one-sample CTAs would no longer fit in one wave (8 CTAs of NT threads on each of the
// 148 SMs); below that the one-sample grid has more parallelism (kernel 0.126 vs 0.146 ms at batch 512; 0.625 vs 0.337 at 8192).
// AMPC_MLP_BLOCKED=0|1 forces one or the other (tests run both on the same inputs).
static bool ampc_mlp_blocked(int batch) {
  if (const char *f = getenv("AMPC_MLP_BLOCKED")) return f[0] == '1';
  return batch > 148 * 8;
}

static int run_mlp(ampc_mlp *m, int batch, const double *X, const double *U, double *Xn, double *Jx, double *Ju) {
  AMPC_REQUIRE(m && X && U && Xn && batch >= 0, AMPC_ERR_INVALID, "bad argument");
  if (batch == 0) return AMPC_OK;
  AMPC_CUDA_CHECK(cudaSetDevice(m->device));
  const int nx = m->net.nx, nu = m->net.nu;
  const size_t nX = (size_t)batch * nx, nU = (size_t)batch * nu;
  const size_t nJx = Jx ? nX * nx : 0, nJu = Jx ? nX * nu : 0;
  double *d = nullptr;
  if (int rc = mlp_scratch(m, 2 * nX + nU + nJx + nJu, &d)) return rc;
  double *dX = d, *dU = dX + nX, *dXn = dU + nU, *dJx = dXn + nX, *dJu = dJx + nJx;
  cudaError_t e = cudaMemcpy(dX, X, nX * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dU, U, nU * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cudaEventRecord(m->ev0);
    if (Jx) pred_diff_kernel<<<batch, NTJ, m->smem_diff>>>(m->net, batch, dX, dU, dXn, dJx, dJu);
    else if (ampc_mlp_blocked(batch))
      pred_batch_blocked_kernel<<<(batch + AMPC_MLP_SB - 1) / AMPC_MLP_SB, NT, m->smem_batch>>>(m->net, batch, dX, dU, dXn);
    else pred_batch_kernel<<<batch, NT, m->smem_pred>>>(m->net, batch, dX, dU, dXn);
    cudaEventRecord(m->ev1);
    ampc_count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(Xn, dXn, nX * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && Jx) e = cudaMemcpy(Jx, dJx, nJx * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && Jx) e = cudaMemcpy(Ju, dJu, nJu * sizeof(double), cudaMemcpyDeviceToHost);
  AMPC_CUDA_CHECK(e);
  return AMPC_OK;
}

extern "C" int ampc_mlp_pred_batch(ampc_mlp *m, int32_t batch, const double *X, const double *U, double *Xn) {
  return run_mlp(m, batch, X, U, Xn, nullptr, nullptr);
}

extern "C" int ampc_mlp_pred_diff_batch(ampc_mlp *m, int32_t batch, const double *X, const double *U, double *Xn,
                                        double *Jx, double *Ju) {
  AMPC_REQUIRE(Jx && Ju, AMPC_ERR_INVALID, "null Jacobian output");
  return run_mlp(m, batch, X, U, Xn, Jx, Ju);
}

extern "C" int ampc_mlp_rollout_batch(ampc_mlp *m, int32_t batch, int32_t horizon, const double *X0, const double *U,
                                      double *Xh) {
  AMPC_REQUIRE(m && X0 && U && Xh && batch >= 0 && horizon >= 1, AMPC_ERR_INVALID, "bad argument");
  if (batch == 0) return AMPC_OK;
  AMPC_CUDA_CHECK(cudaSetDevice(m->device));
  const int nx = m->net.nx, nu = m->net.nu;
  const size_t nX = (size_t)batch * nx, nU = (size_t)horizon * batch * nu;
  double *d = nullptr;
  if (int rc = mlp_scratch(m, 2 * nX + nU, &d)) return rc;
  double *dX = d, *dU = dX + nX, *dXh = dU + nU;
  cudaError_t e = cudaMemcpy(dX, X0, nX * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dU, U, nU * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cudaEventRecord(m->ev0);
    if (ampc_mlp_blocked(batch))
      rollout_batch_blocked_kernel<<<(batch + AMPC_MLP_SB - 1) / AMPC_MLP_SB, NT,
                                     m->smem_batch + (size_t)AMPC_MLP_SB * nx * sizeof(double)>>>(m->net, batch, horizon, dX,
                                                                                                  dU, dXh);
    else
      rollout_batch_kernel<<<batch, NT, m->smem_pred + nx * sizeof(double)>>>(m->net, batch, horizon, dX, dU, dXh);
    cudaEventRecord(m->ev1);
    ampc_count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(Xh, dXh, nX * sizeof(double), cudaMemcpyDeviceToHost);
  AMPC_CUDA_CHECK(e);
  return AMPC_OK;
}

static int run_nmpc(ampc_mlp *m, int32_t H, const double *x, double *out, bool jac) {
  AMPC_REQUIRE(m && x && out && H >= 1, AMPC_ERR_INVALID, "bad argument");
  AMPC_CUDA_CHECK(cudaSetDevice(m->device));
  const int nx = m->net.nx, nu = m->net.nu;
  const size_t n_in = (size_t)(H + 1) * nx + (size_t)H * nu;
  const size_t n_out = jac ? (size_t)H * (nx * nx + nx * nu + nx) : (size_t)H * nx;
  double *d = nullptr;
  if (int rc = mlp_scratch(m, n_in + n_out, &d)) return rc;
  cudaError_t e = cudaMemcpy(d, x, n_in * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    cudaEventRecord(m->ev0);
    if (jac) nmpc_knot_kernel<true><<<H, NTJ, m->smem_diff>>>(m->net, H, d, d + n_in);
    else nmpc_knot_kernel<false><<<H, NTJ, m->smem_diff>>>(m->net, H, d, d + n_in);
    cudaEventRecord(m->ev1);
    ampc_count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d + n_in, n_out * sizeof(double), cudaMemcpyDeviceToHost);
  AMPC_CUDA_CHECK(e);
  return AMPC_OK;
}

extern "C" int ampc_mlp_nmpc_constraint(ampc_mlp *m, int32_t H, const double *x, double *c) {
  return run_nmpc(m, H, x, c, false);
}

extern "C" int ampc_mlp_nmpc_jacobian(ampc_mlp *m, int32_t H, const double *x, double *jac) {
  return run_nmpc(m, H, x, jac, true);
}

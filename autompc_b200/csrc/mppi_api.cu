// C ABI for the MPPI solve: handle, weight packing, solve / merge / parity taps.
#include <map>
#include <mutex>
#include <utility>
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "ampc_common.cuh"

// ------------------------------------------------------------ error plumbing ---
static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void ampc_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void ampc_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

cudaError_t ampc_raise_smem_limit(const void *func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, size_t> limit;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t &cur = limit[std::make_pair(func, dev)];
  if (bytes <= cur) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

extern "C" const char *ampc_last_error(void) { return g_err; }
extern "C" const char *ampc_version(void) { return "ampc_b200 0.1 (sm_100a)"; }
extern "C" uint64_t ampc_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------ kernels implemented elsewhere ---
int ampc_mppi_fp32_configure(const AmpcMppiParams &p, bool *resident_out, size_t *smem_out);
int ampc_mppi_fp32_launch(const AmpcMppiParams &p, bool resident, size_t smem, cudaStream_t stream);
int ampc_mppi_fp32_grid(const AmpcMppiParams &p);

struct AmpcTcPlan;  // tcgen05 path (mppi_tc.cu)
int ampc_mppi_tc_supported(const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp, const char **why);
int ampc_mppi_tc_create(AmpcTcPlan **plan, const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp);
void ampc_mppi_tc_destroy(AmpcTcPlan *plan);
int ampc_mppi_tc_grid(const AmpcTcPlan *plan);
int ampc_mppi_tc_cta_group(const AmpcTcPlan *plan);
int ampc_mppi_tc_dz(const AmpcTcPlan *plan);
int ampc_mppi_tc_launch(AmpcTcPlan *plan, const AmpcMppiParams &p, cudaStream_t stream);
int ampc_mppi_tc_trace(AmpcTcPlan *plan, unsigned long long *host, int max_words);

// ------------------------------------------------------------------- handle ---
struct ampc_mppi {
  ampc_mppi_cfg cfg;
  AmpcMppiParams p;         // template, per-solve fields filled at launch
  int device = 0;
  int HN = 0;
  int n_partials = 0;
  bool resident = false;
  size_t smem = 0;
  AmpcTcPlan *tc = nullptr;
  cudaStream_t stream = nullptr;   // used by the *_host entry points
  float *d_wpack = nullptr, *d_consts = nullptr, *d_act = nullptr, *d_costs = nullptr, *d_term = nullptr;
  float *d_partials = nullptr, *d_x0 = nullptr, *d_u = nullptr, *d_eps = nullptr;
  unsigned int *d_ticket = nullptr;
  float *h_pin = nullptr;          // pinned staging: x0 (nx) | u (nu) | "control written" sequence number (one word)
  unsigned int host_seq = 0;       // sequence number of the last host-buffer solve
  float *d_pin = nullptr;          // the same buffer as the device sees it (mapped): the control is written straight to it
  float *h_eps = nullptr;          // pinned staging for external eps (lazily)
  size_t eps_elems = 0;
  // NVLink peer exchange
  float *d_mail = nullptr;         // own mailbox (cudaMalloc, exported over CUDA IPC)
  float *d_rec = nullptr;          // staging of the shard's record
  float **d_peer = nullptr;        // device array of the ranks' mailbox pointers
  std::vector<void *> ipc_opened;  // peers' mailboxes opened with cudaIpcOpenMemHandle
  int world = 1, rank = 0;
  unsigned int seq = 0;
  // device-resident closed loop
  double *d_cl = nullptr;          // [ x (nx) | cost (1) | Q | R | F | goal | goal_term | obs (T+1, nx) | ctrl (T, nu) ]
  int cl_T = 0;
  std::vector<double> h_cost;      // Q, R, F, goal, goal_term of the closed loop's trajectory cost (float64)
  // threshold (box) cost terms
  float *d_box = nullptr;          // rollout kernels: n_box x [lo (nx) | hi (nx) | weight], float32
  int n_box = 0;
  double *d_evalbox = nullptr;     // closed loop's trajectory cost: the same layout in float64
  size_t evalbox_cap = 0;          // doubles allocated at d_evalbox
  int n_evalbox = 0;
  bool eval_cost_set = false;      // ampc_mppi_set_eval_cost was called: h_cost / d_evalbox no longer follow the controller's cost
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

int pow2ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

bool is_diag(const double *M, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j && M[i * n + j] != 0.0) return false;
  return true;
}

__global__ void merge_kernel(const float *recs, int n_recs, int H, int nu, float inv_lmda, const float *scale,
                             float *act_seq, float *u_out) {
  extern __shared__ float sm[];
  float *s_act = sm;                                   // H*nu shifted
  float *s_scratch = sm + ((H * nu + 3) & ~3);
  const int HN = H * nu;
  for (int e = threadIdx.x; e < HN; e += blockDim.x) {
    const int i = e / nu, j = e - i * nu;
    const int src = (i + 1 < H) ? i + 1 : H - 1;
    s_act[e] = act_seq[src * nu + j];
  }
  __syncthreads();
  ampc_merge_records(recs, n_recs, 2 + HN, HN, nu, inv_lmda, s_act, scale, act_seq, u_out, nullptr, s_scratch);
}

__global__ void noise_kernel(float *eps, int H, int K, int nu, int k_offset, float sqrt_sigma, uint64_t seed,
                             uint64_t ctr) {
  const int nblk = (nu + 3) >> 2;
  const long long total = (long long)H * K * nblk;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int blk = (int)(t % nblk);
    const long long r = t / nblk;
    const int k = (int)(r % K), i = (int)(r / K);
    float n4[4];
    ampc_normal4(seed, ctr, (uint32_t)(k_offset + k), (uint32_t)i, (uint32_t)blk, n4);
    for (int q = 0; q < 4; ++q) {
      const int j = blk * 4 + q;
      if (j < nu) eps[((size_t)i * K + k) * nu + j] = n4[q] * sqrt_sigma;
    }
  }
}

int validate(const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp, const ampc_quad_cost *cost) {
  AMPC_REQUIRE(cfg && mlp && cost, AMPC_ERR_INVALID, "null argument");
  AMPC_REQUIRE(cfg->K >= 1 && cfg->H >= 2 && cfg->H < 65536, AMPC_ERR_INVALID,
               "need num_path >= 1 and 2 <= horizon < 65536 (got K=%d H=%d)", cfg->K, cfg->H);
  AMPC_REQUIRE(cfg->nx >= 1 && cfg->nu >= 1 && cfg->nx + cfg->nu <= AMPC_MAX_WIDTH, AMPC_ERR_INVALID,
               "bad dims nx=%d nu=%d", cfg->nx, cfg->nu);
  AMPC_REQUIRE(cfg->sigma > 0 && cfg->lmda > 0, AMPC_ERR_INVALID, "sigma and lmda must be > 0");
  AMPC_REQUIRE(cfg->terminal_mode == 0 || cfg->terminal_mode == 1, AMPC_ERR_INVALID, "terminal_mode");
  AMPC_REQUIRE(cfg->k_offset >= 0 && cfg->K_global >= cfg->k_offset + cfg->K, AMPC_ERR_INVALID,
               "shard [%d,%d) outside K_global=%d", cfg->k_offset, cfg->k_offset + cfg->K, cfg->K_global);
  AMPC_REQUIRE(mlp->n_layers >= 2 && mlp->n_layers <= AMPC_MAX_LAYERS, AMPC_ERR_UNSUPPORTED,
               "MLP must have 1..%d hidden layers (got n_layers=%d)", AMPC_MAX_LAYERS - 1, mlp->n_layers);
  AMPC_REQUIRE(mlp->dims[0] == cfg->nx + cfg->nu && mlp->dims[mlp->n_layers] == cfg->nx, AMPC_ERR_INVALID,
               "MLP dims do not match nx+nu -> nx");
  for (int l = 1; l < mlp->n_layers; ++l)
    AMPC_REQUIRE(mlp->dims[l] >= 1 && mlp->dims[l] <= AMPC_MAX_WIDTH, AMPC_ERR_UNSUPPORTED,
                 "hidden width %d outside 1..%d", mlp->dims[l], AMPC_MAX_WIDTH);
  AMPC_REQUIRE(mlp->act >= 0 && mlp->act <= 3, AMPC_ERR_UNSUPPORTED, "unknown activation %d", mlp->act);
  for (int j = 0; j < cfg->nu; ++j)
    AMPC_REQUIRE(isfinite(cost->umin[j]) && isfinite(cost->umax[j]) && cost->umax[j] > 0 &&
                     cost->umin[j] <= cost->umax[j],
                 AMPC_ERR_INVALID,
                 "control %d: MPPI needs finite bounds with umax > 0 (controls are normalised by umax, "
                 "mppi.py:100-102); got [%g, %g]", j, cost->umin[j], cost->umax[j]);
  for (int j = 0; j < cfg->nx + cfg->nu; ++j)
    AMPC_REQUIRE(mlp->xu_std[j] != 0.0, AMPC_ERR_INVALID, "xu_std[%d] == 0", j);
  return AMPC_OK;
}

void free_handle(ampc_mppi *h) {
  if (!h) return;
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);   // a host-buffer solve returns before its kernel has torn down
  if (h->tc) ampc_mppi_tc_destroy(h->tc);
  cudaFree(h->d_wpack); cudaFree(h->d_consts); cudaFree(h->d_act); cudaFree(h->d_costs); cudaFree(h->d_term);
  cudaFree(h->d_partials); cudaFree(h->d_x0); cudaFree(h->d_u); cudaFree(h->d_eps); cudaFree(h->d_ticket);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->h_eps) cudaFreeHost(h->h_eps);
  for (void *q : h->ipc_opened) cudaIpcCloseMemHandle(q);
  cudaFree(h->d_mail); cudaFree(h->d_rec); cudaFree(h->d_peer); cudaFree(h->d_cl);
  cudaFree(h->d_box); cudaFree(h->d_evalbox);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// fused = true: the multi-GPU exchange runs in the kernel's tail (ampc_mppi_connect_peers_* done before)
int launch_rollout(ampc_mppi *h, const float *dev_x0, const float *dev_eps, uint64_t seed, uint64_t counter,
                   float *dev_u, float *dev_record, cudaStream_t stream, const float *inline_x0 = nullptr,
                   bool fused = false, unsigned int *host_flag = nullptr, unsigned int host_seq = 0) {
  AmpcMppiParams p = h->p;
  p.host_flag = host_flag;
  p.host_seq = host_seq;
  p.x0_inline = 0;
  if (inline_x0) {
    p.x0_inline = 1;
    for (int j = 0; j < h->cfg.nx && j < 32; ++j) p.x0_val[j] = inline_x0[j];
  }
  p.x0 = dev_x0;
  p.eps = dev_eps;
  p.seed = seed;
  p.ctr = counter;
  p.u_out = dev_u;
  p.record_out = dev_record;
  p.peer_mail = nullptr;
  p.n_box = h->n_box;
  p.box = h->d_box;
  if (fused) {
    p.record_out = h->d_rec;
    p.peer_mail = h->d_peer;
    p.world = h->world;
    p.rank = h->rank;
    p.seq = ++h->seq;
  }
  if (h->tc) return ampc_mppi_tc_launch(h->tc, p, stream);
  return ampc_mppi_fp32_launch(p, h->resident, h->smem, stream);
}

// external noise (parity mode): float64 host (H,K,nu) -> pinned float32 staging -> device, on the handle's stream
int stage_eps(ampc_mppi *h, const double *host_eps, const float **d_eps_out) {
  const size_t n = (size_t)h->cfg.H * h->cfg.K * h->cfg.nu;
  if (h->eps_elems < n) {
    if (h->h_eps) cudaFreeHost(h->h_eps);
    cudaFree(h->d_eps);
    h->h_eps = nullptr; h->d_eps = nullptr; h->eps_elems = 0;
    AMPC_CUDA_CHECK(cudaMallocHost(&h->h_eps, n * sizeof(float)));
    AMPC_CUDA_CHECK(cudaMalloc(&h->d_eps, n * sizeof(float)));
    h->eps_elems = n;
  }
  for (size_t i = 0; i < n; ++i) h->h_eps[i] = (float)host_eps[i];
  AMPC_CUDA_CHECK(cudaMemcpyAsync(h->d_eps, h->h_eps, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  *d_eps_out = h->d_eps;
  return AMPC_OK;
}

// One solve with HOST buffers on the handle's stream (fused = with the peer exchange in the kernel's tail).
int solve_host_impl(ampc_mppi *h, const double *host_x0, const double *host_eps, uint64_t seed, uint64_t counter,
                    double *host_u, bool fused) {
  const int nx = h->cfg.nx, nu = h->cfg.nu;
  for (int j = 0; j < nx; ++j) h->h_pin[j] = (float)host_x0[j];
  const float *d_eps = nullptr;
  if (host_eps) {
    int rc = stage_eps(h, host_eps, &d_eps);
    if (rc) return rc;
  }
  if (nx <= 32 && !getenv("AMPC_NO_INLINE_IO")) {
    // The only host traffic besides (optional) external noise is the observation in (nx floats) and the control out
    // (nu floats).  The observation rides in the kernel parameters and the last CTA writes the control straight into
    // the mapped pinned buffer: one launch + one synchronise instead of copy -> launch -> copy on the stream.
    // The kernel's last CTA also publishes a sequence number behind the control (system-scope fence in between); the
    // host spins on it and returns when the result has landed -- cudaStreamSynchronize would also wait for the
    // kernel's teardown and costs a driver round trip.  The stream is polled now and then so that a failed launch
    // surfaces as an error instead of a hang.  (AMPC_NO_SPIN=1: synchronise the stream instead.)
    static const bool spin = !getenv("AMPC_NO_SPIN");
    volatile unsigned int *flag = reinterpret_cast<volatile unsigned int *>(h->h_pin + nx + nu);
    const unsigned int seq = ++h->host_seq ? h->host_seq : ++h->host_seq;     // never 0
    int rc = launch_rollout(h, h->d_x0, d_eps, seed, counter, h->d_pin + nx, nullptr, h->stream, h->h_pin, fused,
                            spin ? reinterpret_cast<unsigned int *>(h->d_pin + nx + nu) : nullptr, seq);
    if (rc) return rc;
    if (spin) {
      for (unsigned int it = 1;; ++it) {
        if (*flag == seq) break;
        if ((it & 0x3FFFu) == 0u) {
          const cudaError_t q = cudaStreamQuery(h->stream);
          if (q == cudaSuccess) { if (*flag != seq) AMPC_CUDA_CHECK(cudaStreamSynchronize(h->stream)); break; }
          if (q != cudaErrorNotReady) AMPC_CUDA_CHECK(q);
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
      }
    } else {
      AMPC_CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    for (int j = 0; j < nu; ++j) host_u[j] = h->h_pin[nx + j];
    return AMPC_OK;
  }
  AMPC_CUDA_CHECK(cudaMemcpyAsync(h->d_x0, h->h_pin, nx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  int rc = launch_rollout(h, h->d_x0, d_eps, seed, counter, h->d_u, nullptr, h->stream, nullptr, fused);
  if (rc) return rc;
  AMPC_CUDA_CHECK(cudaMemcpyAsync(h->h_pin + nx, h->d_u, nu * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  AMPC_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  for (int j = 0; j < nu; ++j) host_u[j] = h->h_pin[nx + j];
  return AMPC_OK;
}

}  // namespace

extern "C" int ampc_mppi_create(ampc_mppi **out, const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp,
                                const ampc_quad_cost *cost) {
  AMPC_REQUIRE(out, AMPC_ERR_INVALID, "null out");
  *out = nullptr;
  int rc = validate(cfg, mlp, cost);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  AMPC_REQUIRE(ce == cudaSuccess && ndev > 0, AMPC_ERR_CUDA,
               "no CUDA device: libampc_b200 has no CPU fallback (%s)", cudaGetErrorString(ce));
  AMPC_REQUIRE(cfg->device >= 0 && cfg->device < ndev, AMPC_ERR_INVALID, "device %d of %d", cfg->device, ndev);
  DeviceGuard g(cfg->device);
  cudaDeviceProp prop;
  AMPC_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
  AMPC_REQUIRE(prop.major == 10, AMPC_ERR_UNSUPPORTED,
               "libampc_b200 is built for sm_100a only; device %d is sm_%d%d", cfg->device, prop.major, prop.minor);

  ampc_mppi *h = new ampc_mppi();
  h->cfg = *cfg;
  h->device = cfg->device;
  const int nx = cfg->nx, nu = cfg->nu, H = cfg->H, K = cfg->K;
  h->HN = H * nu;
  AmpcMppiParams &p = h->p;
  memset(&p, 0, sizeof(p));
  p.K = K; p.H = H; p.nx = nx; p.nu = nu;
  p.k_offset = cfg->k_offset; p.K_global = cfg->K_global;
  p.terminal_mode = cfg->terminal_mode;
  p.q_diag = is_diag(cost->Q, nx);
  p.f_diag = is_diag(cost->F, nx);
  p.r_diag = is_diag(cost->R, nu);
  p.act = mlp->act;
  p.n_layers = mlp->n_layers;
  p.max_width = 0;
  for (int l = 0; l <= mlp->n_layers; ++l) {
    p.dims[l] = mlp->dims[l];
    if (mlp->dims[l] > p.max_width) p.max_width = mlp->dims[l];
  }
  p.inv_lmda = (float)(1.0 / cfg->lmda);
  p.sqrt_sigma = (float)sqrt(cfg->sigma);
  p.lam_over_sigma = (float)(cfg->lmda / cfg->sigma);

  h->h_cost.assign(cost->Q, cost->Q + nx * nx);
  h->h_cost.insert(h->h_cost.end(), cost->R, cost->R + nu * nu);
  h->h_cost.insert(h->h_cost.end(), cost->F, cost->F + nx * nx);
  h->h_cost.insert(h->h_cost.end(), cost->goal, cost->goal + nx);
  {
    const double *gt = cost->goal_term ? cost->goal_term : cost->goal;
    h->h_cost.insert(h->h_cost.end(), gt, gt + nx);
  }
  // constants block
  const AmpcConstLayout cl(nx, nu);
  std::vector<float> hc(cl.total, 0.f);
  for (int j = 0; j < nx + nu; ++j) {
    hc[cl.xu_mean + j] = (float)mlp->xu_mean[j];
    hc[cl.xu_inv + j] = (float)(1.0 / mlp->xu_std[j]);
  }
  for (int j = 0; j < nx; ++j) {
    hc[cl.dy_mean + j] = (float)mlp->dy_mean[j];
    hc[cl.dy_std + j] = (float)mlp->dy_std[j];
    hc[cl.goal + j] = (float)cost->goal[j];
    hc[cl.goalF + j] = (float)(cost->goal_term ? cost->goal_term[j] : cost->goal[j]);
  }
  for (int i = 0; i < nx * nx; ++i) { hc[cl.Q + i] = (float)cost->Q[i]; hc[cl.F + i] = (float)cost->F[i]; }
  for (int i = 0; i < nu * nu; ++i) hc[cl.R + i] = (float)cost->R[i];
  for (int j = 0; j < nu; ++j) {   // mppi.py:100-102, :137-138 (ctrl_scale = umax)
    hc[cl.lo + j] = (float)(cost->umin[j] / cost->umax[j]);
    hc[cl.hi + j] = (float)(cost->umax[j] / cost->umax[j]);
    hc[cl.scale + j] = (float)cost->umax[j];
  }

#define AMPC_CREATE_CHECK(expr)                                                                   \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ampc_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,      \
                     cudaGetErrorString(_e));                                                     \
      free_handle(h);                                                                             \
      return AMPC_ERR_CUDA;                                                                       \
    }                                                                                             \
  } while (0)

  AMPC_CREATE_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_consts, cl.total * sizeof(float)));
  AMPC_CREATE_CHECK(cudaMemcpy(h->d_consts, hc.data(), cl.total * sizeof(float), cudaMemcpyHostToDevice));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_act, h->HN * sizeof(float)));
  AMPC_CREATE_CHECK(cudaMemset(h->d_act, 0, h->HN * sizeof(float)));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_costs, (size_t)K * sizeof(float)));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_term, sizeof(float)));
  AMPC_CREATE_CHECK(cudaMemset(h->d_term, 0, sizeof(float)));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_x0, nx * sizeof(float)));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_u, nu * sizeof(float)));
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_ticket, sizeof(unsigned int)));
  AMPC_CREATE_CHECK(cudaMemset(h->d_ticket, 0, sizeof(unsigned int)));
  AMPC_CREATE_CHECK(cudaHostAlloc(&h->h_pin, (nx + nu + 1) * sizeof(float), cudaHostAllocMapped));
  memset(h->h_pin, 0, (nx + nu + 1) * sizeof(float));
  AMPC_CREATE_CHECK(cudaHostGetDevicePointer((void **)&h->d_pin, h->h_pin, 0));
  p.consts = h->d_consts; p.act_seq = h->d_act; p.costs = h->d_costs; p.term_out = h->d_term;
  p.ticket = h->d_ticket;

  if (cfg->precision == AMPC_PREC_BF16 || cfg->precision == AMPC_PREC_FP16) {
    const char *why = "";
    if (!ampc_mppi_tc_supported(cfg, mlp, &why)) {
      ampc_set_error("precision=%s (tcgen05) does not support this problem: %s",
                     cfg->precision == AMPC_PREC_FP16 ? "fp16" : "bf16", why);
      free_handle(h);
      return AMPC_ERR_UNSUPPORTED;
    }
    rc = ampc_mppi_tc_create(&h->tc, cfg, mlp);
    if (rc) { free_handle(h); return rc; }
    h->n_partials = ampc_mppi_tc_grid(h->tc);
  } else if (cfg->precision == AMPC_PREC_FP32) {
    // fp32 pack: per layer Wt[k][8*npt] (in-major, zero padded) then bias[8*npt]
    int off = 0;
    for (int l = 0; l < mlp->n_layers; ++l) {
      const int N = mlp->dims[l + 1], Kin = mlp->dims[l];
      int npt = pow2ceil((N + 7) / 8);
      if (npt == 0) npt = 1;
      p.npt[l] = npt;
      p.woff[l] = off; off += Kin * 8 * npt;
      p.boff[l] = off; off += 8 * npt;
      off = (off + 3) & ~3;
    }
    p.wpack_floats = off;
    std::vector<float> hw(off, 0.f);
    for (int l = 0; l < mlp->n_layers; ++l) {
      const int N = mlp->dims[l + 1], Kin = mlp->dims[l], NP = 8 * p.npt[l];
      for (int k = 0; k < Kin; ++k)
        for (int j = 0; j < N; ++j) hw[p.woff[l] + k * NP + j] = (float)mlp->W[l][(size_t)j * Kin + k];
      for (int j = 0; j < N; ++j) hw[p.boff[l] + j] = (float)mlp->b[l][j];
    }
    AMPC_CREATE_CHECK(cudaMalloc(&h->d_wpack, off * sizeof(float)));
    AMPC_CREATE_CHECK(cudaMemcpy(h->d_wpack, hw.data(), off * sizeof(float), cudaMemcpyHostToDevice));
    p.wpack = h->d_wpack;
    rc = ampc_mppi_fp32_configure(p, &h->resident, &h->smem);
    if (rc) { free_handle(h); return rc; }
    h->n_partials = ampc_mppi_fp32_grid(p);
  } else {
    ampc_set_error("unknown precision %d", cfg->precision);
    free_handle(h);
    return AMPC_ERR_INVALID;
  }
  AMPC_CREATE_CHECK(cudaMalloc(&h->d_partials, (size_t)h->n_partials * (2 + h->HN) * sizeof(float)));
  p.partials = h->d_partials;
  *out = h;
  return AMPC_OK;
}

extern "C" int ampc_mppi_destroy(ampc_mppi *h) {
  free_handle(h);
  return AMPC_OK;
}

extern "C" int ampc_mppi_set_act_seq(ampc_mppi *h, const double *host_act) {
  AMPC_REQUIRE(h && host_act, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  std::vector<float> t(h->HN);
  for (int i = 0; i < h->HN; ++i) t[i] = (float)host_act[i];
  AMPC_CUDA_CHECK(cudaMemcpy(h->d_act, t.data(), h->HN * sizeof(float), cudaMemcpyHostToDevice));
  return AMPC_OK;
}

extern "C" int ampc_mppi_get_act_seq(ampc_mppi *h, double *host_act) {
  AMPC_REQUIRE(h && host_act, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  std::vector<float> t(h->HN);
  AMPC_CUDA_CHECK(cudaDeviceSynchronize());
  AMPC_CUDA_CHECK(cudaMemcpy(t.data(), h->d_act, h->HN * sizeof(float), cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->HN; ++i) host_act[i] = t[i];
  return AMPC_OK;
}

extern "C" int ampc_mppi_solve(ampc_mppi *h, const float *dev_x0, const float *dev_eps, uint64_t seed,
                               uint64_t counter, float *dev_u, void *stream) {
  AMPC_REQUIRE(h && dev_x0 && dev_u, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  return launch_rollout(h, dev_x0, dev_eps, seed, counter, dev_u, nullptr, (cudaStream_t)stream);
}

extern "C" int ampc_mppi_solve_host(ampc_mppi *h, const double *host_x0, const double *host_eps, uint64_t seed,
                                    uint64_t counter, double *host_u) {
  AMPC_REQUIRE(h && host_x0 && host_u, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  return solve_host_impl(h, host_x0, host_eps, seed, counter, host_u, false);
}

namespace {
// (n, lo, hi, weight) -> host block n x [lo (nx) | hi (nx) | weight]
template <typename T>
int pack_box(int nx, int32_t n_terms, const double *lo, const double *hi, const double *weight, std::vector<T> *out) {
  AMPC_REQUIRE(n_terms >= 0 && n_terms <= AMPC_MAX_BOX_TERMS, AMPC_ERR_INVALID, "n_terms %d outside 0..%d", n_terms,
               AMPC_MAX_BOX_TERMS);
  AMPC_REQUIRE(n_terms == 0 || (lo && hi), AMPC_ERR_INVALID, "null box limits");
  out->assign((size_t)n_terms * (2 * nx + 1), T(0));
  for (int b = 0; b < n_terms; ++b) {
    T *row = out->data() + (size_t)b * (2 * nx + 1);
    for (int j = 0; j < nx; ++j) {
      AMPC_REQUIRE(!isnan(lo[b * nx + j]) && !isnan(hi[b * nx + j]), AMPC_ERR_INVALID, "NaN box limit");
      row[j] = (T)lo[b * nx + j];
      row[nx + j] = (T)hi[b * nx + j];
    }
    row[2 * nx] = weight ? (T)weight[b] : T(1);
  }
  return AMPC_OK;
}
}  // namespace

static int upload_evalbox(ampc_mppi *h, const std::vector<double> &blk, int n_terms);

extern "C" int ampc_mppi_set_box_costs(ampc_mppi *h, int32_t n_terms, const double *lo, const double *hi,
                                       const double *weight) {
  AMPC_REQUIRE(h, AMPC_ERR_INVALID, "null handle");
  DeviceGuard g(h->device);
  std::vector<float> blk;
  int rc = pack_box<float>(h->cfg.nx, n_terms, lo, hi, weight, &blk);
  if (rc) return rc;
  AMPC_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_box);
  h->d_box = nullptr;
  h->n_box = 0;
  if (n_terms > 0) {
    AMPC_CUDA_CHECK(cudaMalloc(&h->d_box, blk.size() * sizeof(float)));
    AMPC_CUDA_CHECK(cudaMemcpy(h->d_box, blk.data(), blk.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->n_box = n_terms;
  }
  if (!h->eval_cost_set) {                     // the closed loop's trajectory cost defaults to the controller's cost
    std::vector<double> blk64;
    rc = pack_box<double>(h->cfg.nx, n_terms, lo, hi, weight, &blk64);
    if (rc) return rc;
    return upload_evalbox(h, blk64, n_terms);
  }
  return AMPC_OK;
}

// (re)fills a device block without a device-wide synchronisation when it already has the capacity (cudaFree would
// wait for every closed loop in flight on the other handles' streams)
static int upload_evalbox(ampc_mppi *h, const std::vector<double> &blk, int n_terms) {
  if (n_terms > 0) {
    if (h->evalbox_cap < blk.size()) {
      cudaFree(h->d_evalbox);
      h->d_evalbox = nullptr;
      h->evalbox_cap = 0;
      const size_t cap = (size_t)AMPC_MAX_BOX_TERMS * (2 * h->cfg.nx + 1);
      AMPC_CUDA_CHECK(cudaMalloc(&h->d_evalbox, cap * sizeof(double)));
      h->evalbox_cap = cap;
    }
    // pageable source: the copy is staged before the call returns, `blk` may go out of scope
    AMPC_CUDA_CHECK(cudaMemcpyAsync(h->d_evalbox, blk.data(), blk.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  h->n_evalbox = n_terms;
  return AMPC_OK;
}

extern "C" int ampc_mppi_set_eval_cost(ampc_mppi *h, const ampc_quad_cost *quad, int32_t n_terms, const double *lo,
                                       const double *hi, const double *weight) {
  AMPC_REQUIRE(h, AMPC_ERR_INVALID, "null handle");
  DeviceGuard g(h->device);
  const int nx = h->cfg.nx, nu = h->cfg.nu;
  std::vector<double> blk64;
  int rc = pack_box<double>(nx, n_terms, lo, hi, weight, &blk64);
  if (rc) return rc;
  std::vector<double> hc((size_t)2 * nx * nx + (size_t)nu * nu + 2 * (size_t)nx, 0.0);
  if (quad) {
    AMPC_REQUIRE(quad->Q && quad->R && quad->F && quad->goal, AMPC_ERR_INVALID, "null cost matrix");
    double *q = hc.data();
    memcpy(q, quad->Q, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, quad->R, sizeof(double) * nu * nu); q += nu * nu;
    memcpy(q, quad->F, sizeof(double) * nx * nx); q += nx * nx;
    memcpy(q, quad->goal, sizeof(double) * nx); q += nx;
    memcpy(q, quad->goal_term ? quad->goal_term : quad->goal, sizeof(double) * nx);
  }
  h->h_cost = hc;                       // read by the next closed_loop_start (stream-ordered after this call)
  h->eval_cost_set = true;
  return upload_evalbox(h, blk64, n_terms);
}

extern "C" int ampc_mppi_get_costs(ampc_mppi *h, double *host_costs, double *term_const) {
  AMPC_REQUIRE(h && host_costs, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  std::vector<float> t(h->cfg.K);
  float term = 0.f;
  AMPC_CUDA_CHECK(cudaDeviceSynchronize());
  AMPC_CUDA_CHECK(cudaMemcpy(t.data(), h->d_costs, (size_t)h->cfg.K * sizeof(float), cudaMemcpyDeviceToHost));
  AMPC_CUDA_CHECK(cudaMemcpy(&term, h->d_term, sizeof(float), cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->cfg.K; ++i) host_costs[i] = t[i];
  if (term_const) *term_const = term;
  return AMPC_OK;
}

extern "C" int ampc_mppi_get_noise(ampc_mppi *h, uint64_t seed, uint64_t counter, float *host_eps) {
  AMPC_REQUIRE(h && host_eps, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  const size_t n = (size_t)h->cfg.H * h->cfg.K * h->cfg.nu;
  float *d = nullptr;
  AMPC_CUDA_CHECK(cudaMalloc(&d, n * sizeof(float)));
  noise_kernel<<<296, 256, 0, h->stream>>>(d, h->cfg.H, h->cfg.K, h->cfg.nu, h->cfg.k_offset, h->p.sqrt_sigma,
                                           seed, counter);
  ampc_count_launch();
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(host_eps, d, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  AMPC_CUDA_CHECK(e);
  return AMPC_OK;
}

extern "C" int ampc_mppi_record_floats(const ampc_mppi *h) { return h ? 2 + h->HN : 0; }

extern "C" int ampc_mppi_rollout_partial(ampc_mppi *h, const float *dev_x0, const float *dev_eps, uint64_t seed,
                                         uint64_t counter, float *dev_record, void *stream) {
  AMPC_REQUIRE(h && dev_x0 && dev_record, AMPC_ERR_INVALID, "null argument");
  DeviceGuard g(h->device);
  return launch_rollout(h, dev_x0, dev_eps, seed, counter, h->d_u, dev_record, (cudaStream_t)stream);
}

extern "C" int ampc_mppi_merge(ampc_mppi *h, const float *dev_records, int32_t n_records, float *dev_u,
                               void *stream) {
  AMPC_REQUIRE(h && dev_records && dev_u && n_records >= 1, AMPC_ERR_INVALID, "bad argument");
  DeviceGuard g(h->device);
  const AmpcConstLayout cl(h->cfg.nx, h->cfg.nu);
  const size_t smem = (((h->HN + 3) & ~3) + 64 + AMPC_MERGE_CACHE) * sizeof(float);
  merge_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(dev_records, n_records, h->cfg.H, h->cfg.nu, h->p.inv_lmda,
                                                       h->d_consts + cl.scale, h->d_act, dev_u);
  ampc_count_launch();
  AMPC_CUDA_CHECK(cudaGetLastError());
  return AMPC_OK;
}

// Debug tap (not part of the reference surface): timeline of CTA 0 of the tcgen05 kernel when the handle was
// created with AMPC_TC_TRACE=1 in the environment.  Returns the number of 64-bit words written.
extern "C" int ampc_mppi_debug_trace(ampc_mppi *h, unsigned long long *host, int32_t max_words) {
  if (!h || !h->tc || !host) return 0;
  DeviceGuard g(h->device);
  return ampc_mppi_tc_trace(h->tc, host, max_words);
}

// Debug tap: which build of the tcgen05 kernel the handle runs.  0 = not the tensor-core path; otherwise
// cta_group (1 | 2) + 16 if the "dz" build (input layer fed by the output layer's accumulator through kind::tf32).
extern "C" int ampc_mppi_debug_tc_mode(ampc_mppi *h) {
  if (!h || !h->tc) return 0;
  return ampc_mppi_tc_cta_group(h->tc) + (ampc_mppi_tc_dz(h->tc) ? 16 : 0);
}

// ------------------------------------------------------------ NVLink peer exchange ---
namespace {
size_t mail_floats(int world, int rec) { return (size_t)2 * world * rec + 2 * world; }

int ensure_mailbox(ampc_mppi *h, int world) {
  if (h->d_mail) return AMPC_OK;
  const size_t n = mail_floats(world, 2 + h->HN);
  AMPC_CUDA_CHECK(cudaMalloc(&h->d_mail, n * sizeof(float)));
  AMPC_CUDA_CHECK(cudaMemset(h->d_mail, 0, n * sizeof(float)));
  AMPC_CUDA_CHECK(cudaMalloc(&h->d_rec, (2 + h->HN) * sizeof(float)));
  AMPC_CUDA_CHECK(cudaMalloc(&h->d_peer, world * sizeof(float *)));
  AMPC_CUDA_CHECK(cudaDeviceSynchronize());
  return AMPC_OK;
}
}  // namespace

extern "C" int ampc_mppi_mailbox_ipc(ampc_mppi *h, int32_t world, void *ipc_handle_64) {
  AMPC_REQUIRE(h && ipc_handle_64 && world >= 1 && world <= 64, AMPC_ERR_INVALID, "bad argument");
  AMPC_REQUIRE(sizeof(cudaIpcMemHandle_t) == 64, AMPC_ERR_UNSUPPORTED, "unexpected cudaIpcMemHandle_t size");
  DeviceGuard g(h->device);
  int rc = ensure_mailbox(h, world);
  if (rc) return rc;
  cudaIpcMemHandle_t hd;
  AMPC_CUDA_CHECK(cudaIpcGetMemHandle(&hd, h->d_mail));
  memcpy(ipc_handle_64, &hd, 64);
  return AMPC_OK;
}

static int finish_connect(ampc_mppi *h, int world, int rank, const std::vector<float *> &ptrs) {
  AMPC_CUDA_CHECK(cudaMemcpy(h->d_peer, ptrs.data(), world * sizeof(float *), cudaMemcpyHostToDevice));
  h->world = world;
  h->rank = rank;
  h->seq = 0;
  return AMPC_OK;
}

extern "C" int ampc_mppi_connect_peers_ipc(ampc_mppi *h, int32_t world, int32_t rank, const void *ipc_handles) {
  AMPC_REQUIRE(h && ipc_handles && world >= 1 && world <= 64 && rank >= 0 && rank < world, AMPC_ERR_INVALID, "bad argument");
  DeviceGuard g(h->device);
  int rc = ensure_mailbox(h, world);
  if (rc) return rc;
  std::vector<float *> ptrs(world, nullptr);
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ptrs[r] = h->d_mail; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char *)ipc_handles + (size_t)r * 64, 64);
    void *q = nullptr;
    AMPC_CUDA_CHECK(cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(q);
    ptrs[r] = (float *)q;
  }
  return finish_connect(h, world, rank, ptrs);
}

// same-process variant: the ranks' handles live in this process (tests; multi-GPU from one process)
extern "C" int ampc_mppi_connect_peers_local(ampc_mppi *h, int32_t world, int32_t rank, ampc_mppi *const *handles) {
  AMPC_REQUIRE(h && handles && world >= 1 && world <= 64 && rank >= 0 && rank < world && handles[rank] == h,
               AMPC_ERR_INVALID, "bad argument");
  std::vector<float *> ptrs(world, nullptr);
  for (int r = 0; r < world; ++r) {
    AMPC_REQUIRE(handles[r] && handles[r]->HN == h->HN, AMPC_ERR_INVALID, "peer %d has a different horizon/ctrl_dim", r);
    DeviceGuard gr(handles[r]->device);
    int rc = ensure_mailbox(handles[r], world);
    if (rc) return rc;
    ptrs[r] = handles[r]->d_mail;
    if (handles[r]->device != h->device) {
      DeviceGuard gh(h->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(handles[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) AMPC_CUDA_CHECK(e);
      cudaGetLastError();
    }
  }
  DeviceGuard g(h->device);
  return finish_connect(h, world, rank, ptrs);
}

extern "C" int ampc_mppi_solve_fused(ampc_mppi *h, const float *dev_x0, const float *dev_eps, uint64_t seed,
                                     uint64_t counter, float *dev_u, void *stream) {
  AMPC_REQUIRE(h && dev_x0 && dev_u, AMPC_ERR_INVALID, "null argument");
  AMPC_REQUIRE(h->d_peer && h->world >= 1, AMPC_ERR_INVALID, "ampc_mppi_connect_peers_* has not been called");
  DeviceGuard g(h->device);
  return launch_rollout(h, dev_x0, dev_eps, seed, counter, dev_u, nullptr, (cudaStream_t)stream, nullptr, true);
}

extern "C" int ampc_mppi_solve_fused_host(ampc_mppi *h, const double *host_x0, const double *host_eps, uint64_t seed,
                                          uint64_t counter, double *host_u) {
  AMPC_REQUIRE(h && host_x0 && host_u, AMPC_ERR_INVALID, "null argument");
  AMPC_REQUIRE(h->d_peer && h->world >= 1, AMPC_ERR_INVALID, "ampc_mppi_connect_peers_* has not been called");
  DeviceGuard g(h->device);
  return solve_host_impl(h, host_x0, host_eps, seed, counter, host_u, true);
}

// ------------------------------------------------------------ device-resident closed loop ---
struct ampc_mlp;
int ampc_mlp_device(const ampc_mlp *m);
int ampc_mlp_nx(const ampc_mlp *m);
int ampc_mlp_nu(const ampc_mlp *m);
int ampc_mlp_sim_step_launch(ampc_mlp *m, double *d_x, const float *d_u, float *d_x32, double *d_obs_next, double *d_ctrl_t,
                             const double *d_Q, const double *d_R, const double *d_goal, const double *d_box, int n_box,
                             double *d_cost, cudaStream_t s);
int ampc_traj_cost_final_launch(int nx, const double *d_x, const double *d_Q, const double *d_F, const double *d_goal,
                                const double *d_goalF, const double *d_box, int n_box, double *d_cost, cudaStream_t s);

extern "C" int ampc_mppi_closed_loop_start(ampc_mppi *h, ampc_mlp *sim, const double *x0, int32_t T, uint64_t seed,
                                           uint64_t counter0) {
  AMPC_REQUIRE(h && sim && x0 && T >= 1, AMPC_ERR_INVALID, "bad argument");
  AMPC_REQUIRE(ampc_mlp_device(sim) == h->device && ampc_mlp_nx(sim) == h->cfg.nx && ampc_mlp_nu(sim) == h->cfg.nu,
               AMPC_ERR_INVALID, "simulation model and controller disagree on device / dimensions");
  AMPC_REQUIRE(h->cfg.k_offset == 0 && h->cfg.K_global == h->cfg.K, AMPC_ERR_UNSUPPORTED,
               "the device-resident closed loop runs on one GPU (unsharded controller)");
  DeviceGuard g(h->device);
  const int nx = h->cfg.nx, nu = h->cfg.nu;
  const size_t fixed = (size_t)nx + 1 + (size_t)2 * nx * nx + (size_t)nu * nu + 2 * (size_t)nx;
  if (h->cl_T < T) {
    cudaFree(h->d_cl);
    h->d_cl = nullptr;
    h->cl_T = 0;
    AMPC_CUDA_CHECK(cudaMalloc(&h->d_cl, (fixed + (size_t)(T + 1) * nx + (size_t)T * nu) * sizeof(double)));
    h->cl_T = T;
  }
  double *d_x = h->d_cl, *d_cost = d_x + nx, *d_Q = d_cost + 1, *d_R = d_Q + nx * nx, *d_F = d_R + nu * nu,
         *d_goal = d_F + nx * nx, *d_goalF = d_goal + nx, *d_obs = d_goalF + nx,
         *d_ctrl = d_obs + (size_t)(h->cl_T + 1) * nx;
  std::vector<double> init(fixed, 0.0);
  for (int j = 0; j < nx; ++j) init[j] = x0[j];
  for (size_t i = 0; i < h->h_cost.size(); ++i) init[nx + 1 + i] = h->h_cost[i];
  AMPC_CUDA_CHECK(cudaMemcpyAsync(d_x, init.data(), fixed * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  AMPC_CUDA_CHECK(cudaMemcpyAsync(d_obs, x0, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  for (int j = 0; j < nx; ++j) h->h_pin[j] = (float)x0[j];
  AMPC_CUDA_CHECK(cudaMemcpyAsync(h->d_x0, h->h_pin, nx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  AMPC_CUDA_CHECK(cudaStreamSynchronize(h->stream));      // `init` and h_pin are reused by the caller / next call
  for (int t = 0; t < T; ++t) {
    int rc = launch_rollout(h, h->d_x0, nullptr, seed, counter0 + (uint64_t)t, h->d_u, nullptr, h->stream);
    if (rc) return rc;
    rc = ampc_mlp_sim_step_launch(sim, d_x, h->d_u, h->d_x0, d_obs + (size_t)(t + 1) * nx, d_ctrl + (size_t)t * nu, d_Q,
                                  d_R, d_goal, h->d_evalbox, h->n_evalbox, d_cost, h->stream);
    if (rc) return rc;
  }
  return ampc_traj_cost_final_launch(nx, d_x, d_Q, d_F, d_goal, d_goalF, h->d_evalbox, h->n_evalbox, d_cost, h->stream);
}

extern "C" int ampc_mppi_closed_loop_finish(ampc_mppi *h, int32_t T, double *obs_out, double *ctrl_out, double *cost_out) {
  AMPC_REQUIRE(h && h->d_cl && T >= 1 && T <= h->cl_T, AMPC_ERR_INVALID, "no closed loop of that length in flight");
  DeviceGuard g(h->device);
  const int nx = h->cfg.nx, nu = h->cfg.nu;
  double *d_x = h->d_cl, *d_cost = d_x + nx, *d_obs = d_cost + 1 + 2 * nx * nx + nu * nu + 2 * nx,
         *d_ctrl = d_obs + (size_t)(h->cl_T + 1) * nx;
  AMPC_CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (obs_out) AMPC_CUDA_CHECK(cudaMemcpy(obs_out, d_obs, (size_t)(T + 1) * nx * sizeof(double), cudaMemcpyDeviceToHost));
  if (ctrl_out) AMPC_CUDA_CHECK(cudaMemcpy(ctrl_out, d_ctrl, (size_t)T * nu * sizeof(double), cudaMemcpyDeviceToHost));
  if (cost_out) AMPC_CUDA_CHECK(cudaMemcpy(cost_out, d_cost, sizeof(double), cudaMemcpyDeviceToHost));
  return AMPC_OK;
}

// Host side of the tcgen05 MPPI path: shape planning, 16-bit (bf16 or IEEE half) weight images, launch.  The kernel lives in
// mppi_tc_kernel.cuh and is instantiated in mppi_tc_inst_*.cu.
#include <algorithm>
#include <vector>

#include "mppi_tc_kernel.cuh"

using namespace ampc_tc;

namespace {

uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)((u >> 16) | ((u & 0xFFFFu) ? 0x40u : 0u));
  u += 0x7FFFu + ((u >> 16) & 1u);   // round to nearest even
  return (uint16_t)(u >> 16);
}

// IEEE half, round to nearest even, saturating at +-65504 (the kernel's conversions are .satfinite too)
uint16_t f32_to_f16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const uint32_t mag = u & 0x7FFFFFFFu;
  if (mag > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);           // NaN
  if (mag >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);          // >= 65520 rounds past the largest finite half
  if (mag < 0x33000001u) return sign;                                 // <= 2^-25: rounds to zero
  int e = (int)(mag >> 23) - 127;
  uint32_t m = (mag & 0x7FFFFFu) | 0x800000u;                         // 24-bit significand
  int shift = (e < -14) ? (13 + (-14 - e)) : 13;                      // subnormal halves lose more bits
  uint32_t h = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (h & 1u))) ++h;
  if (e < -14) return (uint16_t)(sign | h);                           // subnormal (a carry into 0x400 is the smallest normal)
  return (uint16_t)(sign | (uint16_t)(((uint32_t)(e + 15) << 10) + (h - 0x400u)));   // carry propagates into the exponent
}
float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  int e = (h >> 10) & 31;
  uint32_t m = h & 0x3FFu;
  uint32_t u;
  if (e == 0) {
    if (m == 0) { u = sign; }
    else {
      e = 1;
      while (!(m & 0x400u)) { m <<= 1; --e; }
      u = sign | ((uint32_t)(e + 112) << 23) | ((m & 0x3FFu) << 13);
    }
  } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
  else u = sign | ((uint32_t)(e + 112) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}
uint16_t f32_to_16(float f, bool f16) { return f16 ? f32_to_f16(f) : f32_to_bf16(f); }
float f16or_bf16_to_f32(uint16_t h, bool f16) {
  if (f16) return f16_to_f32(h);
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

size_t tc_smem_bytes(const TcArgs &a, int nx, int nu, int H) {
  const AmpcConstLayout cl(nx, nu);
  const size_t floats = (size_t)cl.total + ((H * nu + 3) & ~3) + (size_t)nx * TM +
                        (size_t)2 * nu * TM + 2 * 64 + 2 * 2 * 32 + 4 * 32 + 2 * TM + 32 + 64 + AMPC_MERGE_CACHE;
  return 1024 + a.w_bytes + floats * sizeof(float) + (2 * MAXG + 2) * sizeof(uint64_t) + 16 +
         (getenv("AMPC_TC_TRACE") ? (NTHR / 32) * TRACE_EV * sizeof(uint32_t) + 16 : 0);
}

int roundup(int v, int m) { return (v + m - 1) / m * m; }

int chunk_width(int npad, bool last) {
  // hidden layers are issued as N-chunks so that epilogues overlap the remaining MMAs; the output layer is one chunk
  if (last || npad < 128) return npad;
  return npad / 2;                    // hidden GEMMs of width >= 128 are issued as two N-halves
}

// fp32 -> tf32 (10 explicit mantissa bits), round to nearest even; the tensor core ignores the low 13 bits
float f32_to_tf32(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) != 0x7F800000u) u += 0xFFFu + ((u >> 13) & 1u);
  u &= ~0x1FFFu;
  memcpy(&f, &u, 4);
  return f;
}

void fill_args(TcArgs &a, const ampc_mlp_desc *mlp, int cg, bool f16, bool dz) {
  memset(&a, 0, sizeof(a));
  a.n_layers = mlp->n_layers;
  {
    const int nx = mlp->dims[mlp->n_layers];
    a.nxp = nx <= 4 ? 4 : (nx <= 8 ? 8 : (nx <= 16 ? 16 : (nx <= 24 ? 24 : 32)));
  }
  uint32_t off = 0;
  int boff = 0;
  for (int l = 0; l < mlp->n_layers; ++l) {
    // K of the layer's data; hidden layers also consume a constant-one K-step (bias): inside the input layer's
    // padding (two extra columns), as an extra K block for layers 1..L-2
    a.kpad[l] = (l == 0) ? roundup(a.nxp + (mlp->dims[0] - mlp->dims[mlp->n_layers]) + 2, 16) : a.npad[l - 1];
    {
      const int nl = mlp->dims[l + 1];
      a.npad[l] = (l == mlp->n_layers - 1) ? roundup(nl, 32) : (nl <= 64 ? 64 : (nl <= 128 ? 128 : 256));
    }
    const bool bias_k = (l >= 1);                 // layers 1..L-1 take their bias through an extra constant-one K block
    a.ones[l] = (l <= mlp->n_layers - 2) ? 1 : 0;  // ... written by the epilogue of the layer before them
    const int rows = a.npad[l] / cg, kblk = (a.kpad[l] + 63) / 64 + (bias_k ? 1 : 0);
    a.w_off[l] = off;
    // input layer with K <= 32 and two N-halves: a row of the 64-wide K block is half empty, so N-half 1 is stored in
    // bytes [64,128) of N-half 0's rows (its descriptor starts 64 bytes further, like K-steps 2 and 3 would)
    a.cw[l] = chunk_width(a.npad[l], l == mlp->n_layers - 1);
    if (l == 0 && a.kpad[0] <= 32 && a.cw[0] * 2 == a.npad[0]) a.l0_packed = 1;
    off += (uint32_t)kblk * rows * 128 / ((l == 0 && a.l0_packed) ? 2 : 1);
    a.b_off[l] = boff;
    boff += a.npad[l];
    // kind::f16: c=f32 (bit 4), a/b format at bits 7 / 10 (0 = f16, 1 = bf16), both K-major, N>>3 at 17, M>>4 at 24
    a.nch[l] = a.npad[l] / a.cw[l];
    a.nh[l] = a.nch[l];
    a.hwid[l] = a.cw[l];
    a.nkp[l] = (l == 0) ? 1 : a.nh[l - 1];
    a.awid[l] = (l == 0) ? 64 : a.hwid[l - 1];
    a.idesc[l] = (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(a.cw[l] >> 3) << 17) |
                 ((uint32_t)((TM * cg) >> 4) << 24);
  }
  // dz mode: fp32 (tf32-rounded) image of the input layer's state columns, one 128-byte row (32 K elements) per neuron
  a.nkx = (a.nxp + 7) / 8;
  a.idesc_x = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.cw[0] >> 3) << 17) | ((uint32_t)((TM * cg) >> 4) << 24);
  // (needs a hidden GEMM between the input layer's reads of its operand and the next stores to it, and the ReLU build)
  if (dz && mlp->n_layers >= 3 && mlp->act == AMPC_ACT_RELU) {
    a.dz = 1;
    // the early 16-bit input block goes where the last hidden GEMM's A buffer has 32 columns its MMAs never read
    a.ecol = (a.npad[mlp->n_layers - 3] == 256) ? 160 : 128;
    a.wx_off = off;
    off += (uint32_t)(a.npad[0] / cg) * 128;
  }
  a.w_bytes = off;
  a.bias_floats = boff;
}

}  // namespace

// the timeline build (AMPC_TC_TRACE=1) exists for the headline shape only: CTA pairs, NXP = 24, ReLU
static bool tc_trace_available(int cg, int nxp, int act) { return cg == 2 && nxp == 24 && act == AMPC_ACT_RELU; }
static TcKernel tc_kernel_ptr(int cg, int nxp, int act, bool f16, bool traced, bool dz) {
  if (traced && !f16 && tc_trace_available(cg, nxp, act))
    return dz ? ampc_tc_kernel_cg2_nxp24_relu1_f160_trace1_dz1() : ampc_tc_kernel_cg2_nxp24_relu1_f160_trace1_dz0();
  const bool relu = act == AMPC_ACT_RELU;
#define AMPC_TC_PICK_CG(N, F, D, R) \
  return cg == 1 ? ampc_tc_kernel_cg1_nxp##N##_relu##R##_f16##F##_trace0_dz##D() : ampc_tc_kernel_cg2_nxp##N##_relu##R##_f16##F##_trace0_dz##D()
#define AMPC_TC_PICK_F(N, F)                      \
  if (!relu) { AMPC_TC_PICK_CG(N, F, 0, 0); }     \
  else if (dz) { AMPC_TC_PICK_CG(N, F, 1, 1); }   \
  else { AMPC_TC_PICK_CG(N, F, 0, 1); }
#define AMPC_TC_PICK(N) \
  if (f16) { AMPC_TC_PICK_F(N, 1) } else { AMPC_TC_PICK_F(N, 0) }
  switch (nxp) {
    case 4: AMPC_TC_PICK(4)
    case 8: AMPC_TC_PICK(8)
    case 16: AMPC_TC_PICK(16)
    case 24: AMPC_TC_PICK(24)
    default: AMPC_TC_PICK(32)
  }
#undef AMPC_TC_PICK
#undef AMPC_TC_PICK_F
#undef AMPC_TC_PICK_CG
}

struct AmpcTcPlan {
  TcArgs args;
  int cg = 1;
  bool f16 = false;
  int act = AMPC_ACT_RELU;
  int grid = 0;
  size_t smem = 0;
  uint8_t *d_wimg = nullptr;
  float *d_bias = nullptr;
  float *d_epsc = nullptr;
  unsigned long long *d_trace = nullptr;
};

static bool tc_shape_ok(const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp, const char **why) {
  if (cfg->nx > 32) { *why = "nx > 32"; return false; }
  if (cfg->nu > 32) { *why = "nu > 32"; return false; }
  for (int l = 1; l < mlp->n_layers; ++l)
    if (mlp->dims[l] > 256) { *why = "hidden width > 256"; return false; }
  {
    // the input layer's A operand has 4 K-steps = 64 columns: padded state + controls + the two constant-one columns
    const int nxp = cfg->nx <= 4 ? 4 : (cfg->nx <= 8 ? 8 : (cfg->nx <= 16 ? 16 : (cfg->nx <= 24 ? 24 : 32)));
    if (nxp + cfg->nu + 2 > 64) { *why = "padded state + controls + 2 bias columns exceed the 64-column input block"; return false; }
  }
  if (cfg->precision == AMPC_PREC_FP16) {
    // IEEE half operands: every weight (output rows as scaled into z space) and bias must be finite in half
    const int L = mlp->n_layers, nx = mlp->dims[L];
    for (int l = 0; l < L; ++l) {
      const int Kl = mlp->dims[l], Nl = mlp->dims[l + 1];
      for (int n = 0; n < Nl; ++n) {
        const double k1 = (l == L - 1) ? mlp->dy_std[n] / mlp->xu_std[n] : 1.0;
        double mx = fabs((l == L - 1) ? (mlp->b[l][n] * mlp->dy_std[n] + mlp->dy_mean[n]) / mlp->xu_std[n] : mlp->b[l][n]);
        for (int k = 0; k < Kl; ++k) mx = fmax(mx, fabs(mlp->W[l][(size_t)n * Kl + k] * k1));
        if (!(mx < 6.0e4)) { *why = "a weight or bias exceeds the IEEE half range (precision=fp16)"; return false; }
      }
    }
    (void)nx;
  }
  return true;
}

// picks CG=1 when the whole bf16 weight image fits next to the working set, else CG=2 (half per CTA)
static bool tc_dz_wanted() {
  const char *e = getenv("AMPC_TC_DZ");                // A/B knob: AMPC_TC_DZ=0 keeps the owner hop between the output and the input layer
  return !(e && e[0] == '0');
}

static int tc_pick_cg(const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp, size_t cap, TcArgs *out, size_t *smem_out) {
  for (int cg = 1; cg <= 2; ++cg)
    for (int dz = tc_dz_wanted() ? 1 : 0; dz >= 0; --dz) {   // dz mode needs room for one more image: without it otherwise
      TcArgs a;
      fill_args(a, mlp, cg, cfg->precision == AMPC_PREC_FP16, dz != 0);
      const size_t need = tc_smem_bytes(a, cfg->nx, cfg->nu, cfg->H);
      if (need <= cap) {
        *out = a;
        *smem_out = need;
        return cg;
      }
    }
  return 0;
}

int ampc_mppi_tc_supported(const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp, const char **why) {
  if (!tc_shape_ok(cfg, mlp, why)) return 0;
  int dev = 0, max_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    *why = "cannot query the device";
    return 0;
  }
  TcArgs a;
  size_t smem = 0;
  if (!tc_pick_cg(cfg, mlp, (size_t)max_optin, &a, &smem)) {
    *why = "16-bit weight image does not fit in the shared memory of a CTA pair";
    return 0;
  }
  return 1;
}

void ampc_mppi_tc_destroy(AmpcTcPlan *plan) {
  if (!plan) return;
  cudaFree(plan->d_wimg);
  cudaFree(plan->d_bias);
  cudaFree(plan->d_epsc);
  cudaFree(plan->d_trace);
  delete plan;
}

int ampc_mppi_tc_create(AmpcTcPlan **out, const ampc_mppi_cfg *cfg, const ampc_mlp_desc *mlp) {
  *out = nullptr;
  int dev = 0, max_optin = 0;
  AMPC_CUDA_CHECK(cudaGetDevice(&dev));
  AMPC_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  AmpcTcPlan *pl = new AmpcTcPlan();
  int cg = tc_pick_cg(cfg, mlp, (size_t)max_optin, &pl->args, &pl->smem);
  const char *force = getenv("AMPC_TC_FORCE_CG");
  if (force && (force[0] == '1' || force[0] == '2')) {
    const int f = force[0] - '0';
    cg = 0;
    for (int dz = tc_dz_wanted() ? 1 : 0; dz >= 0 && !cg; --dz) {
      fill_args(pl->args, mlp, f, cfg->precision == AMPC_PREC_FP16, dz != 0);
      pl->smem = tc_smem_bytes(pl->args, cfg->nx, cfg->nu, cfg->H);
      cg = (pl->smem <= (size_t)max_optin) ? f : 0;
    }
  }
  if (!cg) {
    delete pl;
    ampc_set_error("tcgen05 MPPI path: weights do not fit in shared memory");
    return AMPC_ERR_UNSUPPORTED;
  }
  pl->cg = cg;
  pl->f16 = cfg->precision == AMPC_PREC_FP16;
  const bool f16 = pl->f16;
  pl->act = mlp->act;
  TcArgs &a = pl->args;
  const int tiles = (cfg->K + TM - 1) / TM;
  pl->grid = roundup(tiles, cg);
  a.Kc = pl->grid * TM;
  // ---- weight images: per CTA rank r, per layer, per 64-wide K block: rows x 128 B, 16-byte chunks XOR-swizzled
  std::vector<uint16_t> img((size_t)cg * a.w_bytes / 2, 0);
  std::vector<float> bias(a.bias_floats, 0.f);
  for (int l = 0; l < mlp->n_layers; ++l) {
    const int Kl = mlp->dims[l], Nl = mlp->dims[l + 1];
    const bool bias_k = (l >= 1);
    const bool outl = (l == mlp->n_layers - 1);
    const int kdata = (a.kpad[l] + 63) / 64;      // K blocks of data; the bias block (if any) follows
    const int rows = a.npad[l] / cg, kblk = kdata + (bias_k ? 1 : 0);
    const bool packed = (l == 0 && a.l0_packed);
    for (int r = 0; r < cg; ++r)
      for (int kb = 0; kb < kblk; ++kb)
        for (int n = 0; n < rows; ++n)
          for (int ch = 0; ch < (packed ? 4 : 8); ++ch) {
            const int crow = a.cw[l] / cg;      // rows of one chunk held by each CTA
            // packed input layer: local row n of N-half n / crow lives in physical row n % crow, 16-byte chunks 4..7 for half 1
            const int prow = packed ? n % crow : n, pch = packed ? ch + 4 * (n / crow) : ch;
            const size_t byte = (size_t)r * a.w_bytes + a.w_off[l] + (size_t)kb * rows * 128 + (size_t)prow * 128 +
                                (size_t)((pch ^ (prow & 7)) * 16);
            const int ng = (n / crow) * a.cw[l] + r * crow + (n % crow);   // D column (= neuron) of local row n
            for (int e = 0; e < 8; ++e) {
              const int k = kb * 64 + ch * 8 + e;
              int kin = k;                      // input layer: K columns [0,nxp) = state, [nxp,nxp+nu) = controls
              if (l == 0) {
                const int nx = mlp->dims[mlp->n_layers], nu = Kl - nx;
                kin = (k < a.nxp) ? (k < nx ? k : -1) : (k - a.nxp < nu ? nx + (k - a.nxp) : -1);
              }
              // output layer: rows scaled so that the GEMM yields the increment of the NORMALISED state directly:
              //   z' = z + k1 * (W h + b) + dy_mean / std,  k1 = dy_std / std   (mlp.py:26-30, :236 in z space)
              const double k1 = (outl && ng < Nl) ? mlp->dy_std[ng] / mlp->xu_std[ng] : 1.0;
              float v = (ng < Nl && kin >= 0 && kin < Kl && kb < kdata) ? (float)(mlp->W[l][(size_t)ng * Kl + kin] * k1) : 0.f;
              // bias of hidden layers: two bf16 terms (hi + lo) against the constant-one inputs
              int bsel = -1;
              if (l == 0 && l <= mlp->n_layers - 2) bsel = k - (a.nxp + (Kl - mlp->dims[mlp->n_layers]));
              if (bias_k && kb == kdata) bsel = k - kb * 64;
              if (ng < Nl && (bsel == 0 || bsel == 1)) {
                const float b = outl ? (float)((mlp->b[l][ng] * mlp->dy_std[ng] + mlp->dy_mean[ng]) / mlp->xu_std[ng])
                                     : (float)mlp->b[l][ng];
                const float bhi = f16or_bf16_to_f32(f32_to_16(b, f16), f16);
                v = (bsel == 0) ? bhi : (b - bhi);
              }
              img[byte / 2 + e] = f32_to_16(v, f16);
            }
          }
    for (int j = 0; j < Nl; ++j) bias[a.b_off[l] + j] = (float)mlp->b[l][j];
  }
  if (a.dz) {
    // tf32 image of W0's state columns (the A operand is the increment of the NORMALISED state, which is what W0 eats)
    const int Kl = mlp->dims[0], Nl = mlp->dims[1], nx = mlp->dims[mlp->n_layers];
    const int rows = a.npad[0] / cg, crow = a.cw[0] / cg;
    for (int r = 0; r < cg; ++r)
      for (int n = 0; n < rows; ++n) {
        const int ng = (n / crow) * a.cw[0] + r * crow + (n % crow);
        for (int k = 0; k < 32; ++k) {
          const float v = (ng < Nl && k < nx) ? f32_to_tf32((float)mlp->W[0][(size_t)ng * Kl + k]) : 0.f;
          const size_t byte = (size_t)r * a.w_bytes + a.wx_off + (size_t)n * 128 + (size_t)(((k >> 2) ^ (n & 7)) * 16) + (size_t)(k & 3) * 4;
          memcpy(reinterpret_cast<uint8_t *>(img.data()) + byte, &v, 4);
        }
      }
  }
  cudaError_t e = cudaMalloc(&pl->d_wimg, img.size() * 2);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_wimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_bias, bias.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_epsc, (size_t)cfg->H * cfg->nu * a.Kc * sizeof(float));
  if (e == cudaSuccess)
    e = ampc_raise_smem_limit((const void *)tc_kernel_ptr(cg, a.nxp, mlp->act, f16, false, a.dz != 0), pl->smem);
  if (e != cudaSuccess) {
    ampc_set_error("tcgen05 MPPI path create: %s", cudaGetErrorString(e));
    ampc_mppi_tc_destroy(pl);
    return AMPC_ERR_CUDA;
  }
  a.wimg = pl->d_wimg;
  a.bias = pl->d_bias;
  a.epsc = pl->d_epsc;
  a.trace = nullptr;
  if (a.dz && getenv("AMPC_TC_DZ") && getenv("AMPC_TC_DZ")[0] == '2') a.dz = 2;   // measurement knob, see the issuer
  a.defer_j = DEFER_J;
  if (const char *dj = getenv("AMPC_TC_DEFER")) {   // tuning knob
    const int v = atoi(dj);
    if (v == 2 || v == 4 || v == 6 || v == 8) a.defer_j = v;
  }
  if (getenv("AMPC_TC_TRACE") && !f16 && tc_trace_available(cg, a.nxp, mlp->act)) {
    ampc_raise_smem_limit((const void *)tc_kernel_ptr(cg, a.nxp, mlp->act, f16, true, a.dz != 0), pl->smem);
    if (cudaMalloc(&pl->d_trace, (NTHR / 32) * TRACE_EV * sizeof(unsigned long long)) == cudaSuccess) {
      cudaMemset(pl->d_trace, 0, (NTHR / 32) * TRACE_EV * sizeof(unsigned long long));
      a.trace = pl->d_trace;
    }
  }
  *out = pl;
  return AMPC_OK;
}

int ampc_mppi_tc_grid(const AmpcTcPlan *plan) { return plan->grid; }
int ampc_mppi_tc_cta_group(const AmpcTcPlan *plan) { return plan->cg; }
int ampc_mppi_tc_dz(const AmpcTcPlan *plan) { return plan->args.dz; }

int ampc_mppi_tc_launch(AmpcTcPlan *plan, const AmpcMppiParams &p, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan->grid);
  cfg.blockDim = dim3(NTHR);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan->cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc_kernel_ptr(plan->cg, plan->args.nxp, plan->act, plan->f16, plan->args.trace != nullptr, plan->args.dz != 0), p, plan->args);
  ampc_count_launch();
  AMPC_CUDA_CHECK(e);
  return AMPC_OK;
}

// debug: copies the timeline of CTA 0 (AMPC_TC_TRACE=1) to host; returns the number of 64-bit words
int ampc_mppi_tc_trace(AmpcTcPlan *plan, unsigned long long *host, int max_words) {
  const int n = (NTHR / 32) * TRACE_EV;
  if (!plan || !plan->d_trace || max_words < n) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(host, plan->d_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return n;
}

// losses.cu -- the training losses of NeuralLumen reduced in-kernel, together with their gradient seeds.
//
// Reference (relative to /root/reference/): projects/NeuralLumen/trainer.py:133-149 (_compute_loss, "render" =
// L1 * 3, psnr from mse), projects/neuralangelo/utils/misc.py:74-89 (eikonal / curvature with nan_to_num and the
// ~outside mask), projects/NeuralLumen/utils/utils.py:142-174 (intrinsic_loss with global min-max weight maps,
// regularize_re_loss), imaginaire/trainers/base.py:534-544 (total = sum weight_i * loss_i).
//
// Four small launches: (0) global min/max of the two pseudo-label maps, (1) per-ray terms + d_out seeds,
// (2) per-sample eikonal/curvature terms + seeds, (3) fixed-order final reduction (deterministic).
// Bound: HBM bandwidth (24 B in + 24 B out per sample).
#include "common.cuh"

namespace {

constexpr int kT = 256;
enum { P_RENDER = 0, P_MSE, P_REF, P_SHA, P_NEG, P_POS, P_RAY_TERMS };
enum { P_EIK = 0, P_CURV, P_SAMPLE_TERMS };

__device__ __forceinline__ float sgn(float x) { return (x > 0.0f) - (x < 0.0f); }

// weights_dev (device float[5]) replaces the five by-value loss weights: same kernels, but a captured CUDA graph then
// follows a weight schedule that moves every iteration
__device__ __forceinline__ void resolve_weights(mli_loss_cfg_t& cfg) {
  if (cfg.weights_dev != nullptr) {
    cfg.w_render = __ldg(cfg.weights_dev + 0); cfg.w_eikonal = __ldg(cfg.weights_dev + 1);
    cfg.w_curvature = __ldg(cfg.weights_dev + 2); cfg.w_intrinsic = __ldg(cfg.weights_dev + 3);
    cfg.w_regularize_re = __ldg(cfg.weights_dev + 4);
  }
}

__global__ void minmax_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ mm) {
  __shared__ float red[4][32];
  float lo0 = INFINITY, hi0 = -INFINITY, lo1 = INFINITY, hi1 = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    lo0 = fminf(lo0, a[i]); hi0 = fmaxf(hi0, a[i]);
    lo1 = fminf(lo1, b[i]); hi1 = fmaxf(hi1, b[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo0 = fminf(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = fmaxf(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
    lo1 = fminf(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = fmaxf(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = lo0; red[1][warp] = hi0; red[2][warp] = lo1; red[3][warp] = hi1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo0 = fminf(lo0, red[0][w]); hi0 = fmaxf(hi0, red[1][w]);
      lo1 = fminf(lo1, red[2][w]); hi1 = fmaxf(hi1, red[3][w]);
    }
    mm[0] = lo0; mm[1] = hi0; mm[2] = lo1; mm[3] = hi1;
  }
}

struct RayLossArgs {
  mli_loss_cfg_t cfg;
  const float* out; int n_out, col_or, col_os, n_os, col_ore;
  const float* image; const float* ref; const float* sha; const float* vis;
  const float* mm;
  int64_t R;
  float* d_out;
  float* part;  // [blocks][P_RAY_TERMS]
};

__global__ void __launch_bounds__(kT) ray_loss_kernel(RayLossArgs a) {
  __shared__ float red[32];
  resolve_weights(a.cfg);
  const int64_t r = (int64_t)blockIdx.x * kT + threadIdx.x;
  float t[P_RAY_TERMS];
#pragma unroll
  for (int k = 0; k < P_RAY_TERMS; ++k) t[k] = 0.0f;
  if (r < a.R) {
    const float* o = a.out + r * a.n_out;
    float* d = a.d_out + r * a.n_out;
    for (int c = 0; c < a.n_out; ++c) d[c] = 0.0f;
    const float invR3 = 1.0f / (3.0f * (float)a.R);
    for (int c = 0; c < 3; ++c) {
      const float e = o[c] - a.image[r * 3 + c];
      t[P_RENDER] += fabsf(e);
      t[P_MSE] += e * e;
      d[c] += a.cfg.w_render * 3.0f * invR3 * sgn(e);  // "render" = L1Loss * 3
    }
    if (a.cfg.has_intrinsic) {
      // weight maps: global min-max normalisation (utils.py:145-156), detached
      const float s = a.sha[r], v = a.vis[r];
      const float w_sha = a.cfg.range_sha[0] +
                          (s - a.mm[0]) / fmaxf(a.mm[1] - a.mm[0], 1e-6f) * (a.cfg.range_sha[1] - a.cfg.range_sha[0]);
      const float w_vis = a.cfg.range_vis[0] +
                          (v - a.mm[2]) / fmaxf(a.mm[3] - a.mm[2], 1e-6f) * (a.cfg.range_vis[1] - a.cfg.range_vis[0]);
      const float w_ref = fminf(w_vis, w_sha);
      for (int c = 0; c < 3; ++c) {
        const float e = o[a.col_or + c] - a.ref[r * 3 + c];
        t[P_REF] += fabsf(e) * w_ref;
        d[a.col_or + c] += a.cfg.w_intrinsic * a.cfg.factor_ref * invR3 * sgn(e) * w_ref;
      }
      const float invRs = 1.0f / ((float)a.n_os * (float)a.R);
      for (int c = 0; c < a.n_os; ++c) {
        const float e = o[a.col_os + c] - s;
        t[P_SHA] += fabsf(e) * w_sha;
        d[a.col_os + c] += a.cfg.w_intrinsic * a.cfg.factor_sha * invRs * sgn(e) * w_sha;
      }
    }
    if (a.col_ore >= 0) {
      for (int c = 0; c < 3; ++c) {
        const float x = o[a.col_ore + c];
        if (x < 0.0f) {
          t[P_NEG] += -x;
          d[a.col_ore + c] += a.cfg.w_regularize_re * a.cfg.factor_negative * invR3 * (-1.0f);
        } else {
          const float ex = a.cfg.exponent_positive;
          t[P_POS] += ex == 1.0f ? x : powf(x, ex);
          const float dp = ex == 1.0f ? 1.0f : ex * powf(x, ex - 1.0f);
          d[a.col_ore + c] += a.cfg.w_regularize_re * a.cfg.factor_positive * invR3 * dp;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < P_RAY_TERMS; ++k) {
    const float s = mli_block_sum(t[k], red);
    if (threadIdx.x == 0) a.part[(size_t)blockIdx.x * P_RAY_TERMS + k] = s;
  }
}

__global__ void __launch_bounds__(kT) sample_loss_kernel(mli_loss_cfg_t cfg, const float* __restrict__ gradients,
                                                         const float* __restrict__ hessians,
                                                         const uint8_t* __restrict__ outside, int64_t M, int N,
                                                         float* __restrict__ d_gradients, float* __restrict__ d_hessians,
                                                         float* __restrict__ part) {
  __shared__ float red[32];
  resolve_weights(cfg);
  const int64_t m = (int64_t)blockIdx.x * kT + threadIdx.x;
  float eik = 0.0f, curv = 0.0f;
  if (m < M) {
    const bool in = !outside[m / N];
    const float invM = 1.0f / (float)M;
    const float gx = gradients[m * 3], gy = gradients[m * 3 + 1], gz = gradients[m * 3 + 2];
    const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
    const float err = (nrm - 1.0f) * (nrm - 1.0f);
    float dg[3] = {0.f, 0.f, 0.f};
    if (isfinite(err) && in) {  // nan_to_num(0) then * ~outside
      eik = err;
      if (nrm > 0.0f) {  // torch norm backward: 0 at the origin
        const float k = cfg.w_eikonal * invM * 2.0f * (nrm - 1.0f) / nrm;
        dg[0] = k * gx; dg[1] = k * gy; dg[2] = k * gz;
      }
    }
    d_gradients[m * 3] = dg[0]; d_gradients[m * 3 + 1] = dg[1]; d_gradients[m * 3 + 2] = dg[2];
    if (hessians) {
      const float lap = hessians[m * 3] + hessians[m * 3 + 1] + hessians[m * 3 + 2];
      float dh = 0.0f;
      if (isfinite(lap) && in) {
        curv = fabsf(lap);
        dh = cfg.w_curvature * invM * sgn(lap);
      }
      d_hessians[m * 3] = dh; d_hessians[m * 3 + 1] = dh; d_hessians[m * 3 + 2] = dh;
    }
  }
  eik = mli_block_sum(eik, red);
  if (threadIdx.x == 0) part[(size_t)blockIdx.x * P_SAMPLE_TERMS + P_EIK] = eik;
  curv = mli_block_sum(curv, red);
  if (threadIdx.x == 0) part[(size_t)blockIdx.x * P_SAMPLE_TERMS + P_CURV] = curv;
}

__global__ void final_loss_kernel(mli_loss_cfg_t cfg, const float* __restrict__ part_ray, int nb_ray,
                                  const float* __restrict__ part_smp, int nb_smp, int64_t R, int64_t M, int n_os,
                                  int has_ore, int has_hess, float* __restrict__ losses) {
  __shared__ float red[32];
  resolve_weights(cfg);
  float sums[P_RAY_TERMS + P_SAMPLE_TERMS];
  for (int k = 0; k < P_RAY_TERMS; ++k) {
    float v = 0.0f;
    for (int i = threadIdx.x; i < nb_ray; i += blockDim.x) v += part_ray[(size_t)i * P_RAY_TERMS + k];
    sums[k] = mli_block_sum(v, red);
  }
  for (int k = 0; k < P_SAMPLE_TERMS; ++k) {
    float v = 0.0f;
    for (int i = threadIdx.x; i < nb_smp; i += blockDim.x) v += part_smp[(size_t)i * P_SAMPLE_TERMS + k];
    sums[P_RAY_TERMS + k] = mli_block_sum(v, red);
  }
  if (threadIdx.x == 0) {
    const float R3 = 3.0f * (float)R;
    const float render = 3.0f * sums[P_RENDER] / R3;
    const float mse = sums[P_MSE] / R3;
    const float eik = sums[P_RAY_TERMS + P_EIK] / (float)M;
    const float curv = has_hess ? sums[P_RAY_TERMS + P_CURV] / (float)M : 0.0f;
    const float intr = cfg.has_intrinsic ? cfg.factor_ref * sums[P_REF] / R3 +
                                               cfg.factor_sha * sums[P_SHA] / ((float)n_os * (float)R)
                                         : 0.0f;
    const float reg = has_ore ? cfg.factor_negative * sums[P_NEG] / R3 + cfg.factor_positive * sums[P_POS] / R3 : 0.0f;
    losses[MLI_LOSS_RENDER] = render;
    losses[MLI_LOSS_EIKONAL] = eik;
    losses[MLI_LOSS_CURVATURE] = curv;
    losses[MLI_LOSS_INTRINSIC] = intr;
    losses[MLI_LOSS_REG_RE] = reg;
    losses[MLI_LOSS_MSE] = mse;
    losses[7] = 0.0f;
    losses[MLI_LOSS_TOTAL] = cfg.w_render * render + cfg.w_eikonal * eik + cfg.w_curvature * curv +
                             (cfg.has_intrinsic ? cfg.w_intrinsic * intr : 0.0f) +
                             (has_ore ? cfg.w_regularize_re * reg : 0.0f);
  }
}

}  // namespace

extern "C" int64_t mli_losses_ws_bytes(int64_t R, int64_t M) {
  return (4 + (int64_t)mli_cdiv(R, kT) * P_RAY_TERMS + (int64_t)mli_cdiv(M, kT) * P_SAMPLE_TERMS) * sizeof(float);
}

extern "C" int mli_losses_fwd_bwd(const mli_loss_cfg_t* cfg, int32_t mode, const float* out, const float* gradients,
                                  const float* hessians, const uint8_t* outside, int64_t R, int32_t N,
                                  const float* image, const float* pseudo_ref, const float* pseudo_sha,
                                  const float* pseudo_vis, float* losses, float* d_out, float* d_gradients,
                                  float* d_hessians, void* ws, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(cfg && out && gradients && outside && image && losses && d_out && d_gradients && ws, "losses: NULL argument");
  MLI_REQUIRE(R >= 1 && N >= 1, "losses: bad R/N");
  MLI_REQUIRE(hessians == nullptr || d_hessians != nullptr, "losses: d_hessians required with hessians");
  RayLossArgs a;
  a.cfg = *cfg;
  a.out = out; a.image = image; a.ref = pseudo_ref; a.sha = pseudo_sha; a.vis = pseudo_vis; a.R = R; a.d_out = d_out;
  switch (mode) {
    case MLI_MODE_RGB: a.n_out = 3; a.col_or = a.col_os = a.col_ore = -1; a.n_os = 1; break;
    case MLI_MODE_RGB_R_S: a.n_out = 10; a.col_or = 3; a.col_os = 6; a.n_os = 1; a.col_ore = 7; break;
    case MLI_MODE_RGB_R: a.n_out = 9; a.col_or = 3; a.col_os = 6; a.n_os = 3; a.col_ore = -1; break;
    case MLI_MODE_R_S: a.n_out = 9; a.col_or = 3; a.col_os = 6; a.n_os = 3; a.col_ore = -1; break;
    case MLI_MODE_R_S_RE: a.n_out = 12; a.col_or = 3; a.col_os = 6; a.n_os = 3; a.col_ore = 9; break;
    default: mli_set_error("losses: unknown network_mode %d", mode); return MLI_EINVAL;
  }
  if (a.cfg.has_intrinsic) {
    MLI_REQUIRE(a.col_or >= 0 && pseudo_ref && pseudo_sha && pseudo_vis, "losses: intrinsic loss needs o_r/o_s + pseudo labels");
  }
  const int64_t M = R * N;
  float* mm = (float*)ws;
  float* part_ray = mm + 4;
  const int nb_ray = (int)mli_cdiv(R, kT), nb_smp = (int)mli_cdiv(M, kT);
  float* part_smp = part_ray + (size_t)nb_ray * P_RAY_TERMS;
  a.mm = mm; a.part = part_ray;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.cfg.has_intrinsic) {
    minmax_kernel<<<1, 1024, 0, st>>>(pseudo_sha, pseudo_vis, R, mm);
    MLI_LAUNCH_OK();
  }
  ray_loss_kernel<<<nb_ray, kT, 0, st>>>(a);
  MLI_LAUNCH_OK();
  sample_loss_kernel<<<nb_smp, kT, 0, st>>>(*cfg, gradients, hessians, outside, M, N, d_gradients, d_hessians, part_smp);
  MLI_LAUNCH_OK();
  final_loss_kernel<<<1, 1024, 0, st>>>(*cfg, part_ray, nb_ray, part_smp, nb_smp, R, M, a.n_os, a.col_ore >= 0,
                                        hessians != nullptr, losses);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

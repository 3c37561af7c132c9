// composite.cu -- NeuS SDF->alpha conversion, front-to-back alpha compositing and the per-network_mode merge,
// forward and backward, as warp-level scans over each ray's samples.
//
// Reference (relative to /root/reference/): projects/neuralangelo/model.py:492-515 (compute_neus_alphas,
// _get_iter_cos), projects/nerf/utils/render.py:87-112 (alpha_compositing_weights, composite),
// projects/NeuralLumen/model.py:266-323 (mode merge, white background, o_re), :365-368 and :101-104 (eval extras).
//
// One warp per ray, lane l owns samples [l*SPL, (l+1)*SPL) (contiguous -> vectorisable, coalesced across the
// warp).  Transmittance = exclusive product scan (shuffle-up) over the lanes' local products; the backward's
// suffix recurrence S_{i-1} = dw_i a_i + (1-a_i) S_i is an affine-map suffix scan (shuffle-down), so neither
// direction divides by (1-alpha).  Bound: HBM bandwidth (~48 B in, ~8 B out per sample forward).
#include "common.cuh"

namespace {

constexpr int kMaxSPL = 8;   // samples per lane -> N <= 256
constexpr int kMaxCh = 9;
constexpr int kWarps = 4;

struct CompArgs {
  const float* s_var;
  const float* sdf;        // [M] centre plane (after outside overwrite)
  const float* gradients;  // [M,3]
  const float* ray_unit;   // [R,3]
  const float* dists; int64_t ld_d;
  const float* far;        // [R]
  const float* S; int64_t lds;  // [M, lds] per-sample head outputs (post activation)
  int64_t R;
  int N, n_ch, mode, white_bg, eval_extras;
  float anneal;
  const float* anneal_dev;  // device override of `anneal` (graph replay with a moving schedule)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exclusive product scan across lanes
__device__ __forceinline__ float warp_excl_prod(float p, int lane) {
  float incl = p;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  const float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane == 0 ? 1.0f : excl;
}

__device__ __forceinline__ int n_out_of_mode(int mode) {
  switch (mode) {
    case MLI_MODE_RGB_R_S: return 10;
    case MLI_MODE_RGB_R: return 9;
    case MLI_MODE_R_S: return 9;
    case MLI_MODE_R_S_RE: return 12;
    default: return 3;
  }
}

// composited channels -> output row of the mode (NeuralLumen/model.py:269-310)
__device__ __forceinline__ void mode_merge(int mode, const float* c, float* o) {
  switch (mode) {
    case MLI_MODE_RGB_R_S:  // c = rgb3 | o_r3 | o_s1 -> rgb3 | o_r3 | o_s1 | o_re3
      for (int k = 0; k < 7; ++k) o[k] = c[k];
      for (int k = 0; k < 3; ++k) o[7 + k] = c[k] - c[3 + k] * c[6];
      break;
    case MLI_MODE_RGB_R:  // c = rgb3 | o_r3 -> rgb3 | o_r3 | o_s3 = rgb/o_r
      for (int k = 0; k < 6; ++k) o[k] = c[k];
      for (int k = 0; k < 3; ++k) o[6 + k] = c[k] / c[3 + k];
      break;
    case MLI_MODE_R_S:  // c = o_r3 | o_s3 -> rgb3 = o_r*o_s | o_r3 | o_s3
      for (int k = 0; k < 3; ++k) { o[k] = c[k] * c[3 + k]; o[3 + k] = c[k]; o[6 + k] = c[3 + k]; }
      break;
    case MLI_MODE_R_S_RE:  // c = o_r3 | o_s3 | o_re3 -> rgb3 = o_r*o_s + o_re | o_r3 | o_s3 | o_re3
      for (int k = 0; k < 3; ++k) {
        o[k] = c[k] * c[3 + k] + c[6 + k]; o[3 + k] = c[k]; o[6 + k] = c[3 + k]; o[9 + k] = c[6 + k];
      }
      break;
    default:
      for (int k = 0; k < 3; ++k) o[k] = c[k];
  }
}

// d(output row) -> d(composited channels); c = composited channels of the forward
__device__ __forceinline__ void mode_merge_bwd(int mode, const float* c, const float* d_o, float* d_c) {
  switch (mode) {
    case MLI_MODE_RGB_R_S: {
      float dos = d_o[6];
      for (int k = 0; k < 3; ++k) {
        d_c[k] = d_o[k] + d_o[7 + k];
        d_c[3 + k] = d_o[3 + k] - d_o[7 + k] * c[6];
        dos -= d_o[7 + k] * c[3 + k];
      }
      d_c[6] = dos;
    } break;
    case MLI_MODE_RGB_R:
      for (int k = 0; k < 3; ++k) {
        d_c[k] = d_o[k] + d_o[6 + k] / c[3 + k];
        d_c[3 + k] = d_o[3 + k] - d_o[6 + k] * c[k] / (c[3 + k] * c[3 + k]);
      }
      break;
    case MLI_MODE_R_S:
      for (int k = 0; k < 3; ++k) {
        d_c[k] = d_o[3 + k] + d_o[k] * c[3 + k];
        d_c[3 + k] = d_o[6 + k] + d_o[k] * c[k];
      }
      break;
    case MLI_MODE_R_S_RE:
      for (int k = 0; k < 3; ++k) {
        d_c[k] = d_o[3 + k] + d_o[k] * c[3 + k];
        d_c[3 + k] = d_o[6 + k] + d_o[k] * c[k];
        d_c[6 + k] = d_o[9 + k] + d_o[k];
      }
      break;
    default:
      for (int k = 0; k < 3; ++k) d_c[k] = d_o[k];
  }
}

template <int SPL>
__global__ void __launch_bounds__(32 * kWarps) composite_fwd_kernel(CompArgs a, float* __restrict__ alphas,
                                                                    float* __restrict__ weights,
                                                                    float* __restrict__ out, float* __restrict__ extras) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= a.R) return;
  const float inv_s = expf(a.s_var[0]);
  const float anneal = a.anneal_dev ? __ldg(a.anneal_dev) : a.anneal;
  const float rv[3] = {a.ray_unit[r * 3], a.ray_unit[r * 3 + 1], a.ray_unit[r * 3 + 2]};
  const float far = a.far[r];
  const int64_t base = r * a.N;
  float al[SPL];
  float prod = 1.0f;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const int i = lane * SPL + k;
    al[k] = 0.0f;
    if (i < a.N) {
      const float d0 = a.dists[r * a.ld_d + i];
      const float d1 = (i + 1 < a.N) ? a.dists[r * a.ld_d + i + 1] : far;
      const float g[3] = {a.gradients[(base + i) * 3], a.gradients[(base + i) * 3 + 1], a.gradients[(base + i) * 3 + 2]};
      al[k] = mli_neus_alpha(a.sdf[base + i], g, rv, d1 - d0, inv_s, anneal).alpha;
      prod *= 1.0f - al[k];
    }
  }
  float T = warp_excl_prod(prod, lane);
  float acc[kMaxCh], opa = 0.0f, eg[3] = {0.f, 0.f, 0.f}, ed = 0.0f;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const int i = lane * SPL + k;
    if (i < a.N) {
      const float w = al[k] * T;
      T *= 1.0f - al[k];
      if (alphas) alphas[base + i] = al[k];
      weights[base + i] = w;
      opa += w;
      const float* s = a.S + (base + i) * a.lds;
#pragma unroll
      for (int c = 0; c < kMaxCh; ++c)
        if (c < a.n_ch) acc[c] = fmaf(w, s[c], acc[c]);
      if (a.eval_extras) {
        eg[0] = fmaf(w, a.gradients[(base + i) * 3], eg[0]);
        eg[1] = fmaf(w, a.gradients[(base + i) * 3 + 1], eg[1]);
        eg[2] = fmaf(w, a.gradients[(base + i) * 3 + 2], eg[2]);
        ed = fmaf(w, a.dists[r * a.ld_d + i], ed);
      }
    }
  }
  opa = warp_sum(opa);
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c)
    if (c < a.n_ch) acc[c] = warp_sum(acc[c]);
  if (a.eval_extras) { eg[0] = warp_sum(eg[0]); eg[1] = warp_sum(eg[1]); eg[2] = warp_sum(eg[2]); ed = warp_sum(ed); }
  if (lane == 0) {
    const float white = a.white_bg ? 1.0f - opa : 0.0f;
    float comp[kMaxCh], o[12];
#pragma unroll
    for (int c = 0; c < kMaxCh; ++c) comp[c] = acc[c] + white;
    mode_merge(a.mode, comp, o);
    const int n_out = n_out_of_mode(a.mode);
    for (int c = 0; c < n_out; ++c) out[r * n_out + c] = o[c];
    if (a.eval_extras && extras) {
      extras[r * 5 + 0] = opa; extras[r * 5 + 1] = eg[0]; extras[r * 5 + 2] = eg[1]; extras[r * 5 + 3] = eg[2];
      extras[r * 5 + 4] = ed;
    }
  }
}

template <int SPL>
__global__ void __launch_bounds__(32 * kWarps) composite_bwd_kernel(CompArgs a, const float* __restrict__ weights,
                                                                    const float* __restrict__ d_out,
                                                                    const float* __restrict__ d_weights,
                                                                    uint32_t act_mask, float* __restrict__ dS_pre,
                                                                    float* __restrict__ d_sdf,
                                                                    float* __restrict__ d_gradients,
                                                                    float* __restrict__ d_inv_s_partial) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (r >= a.R) return;
  const float inv_s = expf(a.s_var[0]);
  const float anneal = a.anneal_dev ? __ldg(a.anneal_dev) : a.anneal;
  const float rv[3] = {a.ray_unit[r * 3], a.ray_unit[r * 3 + 1], a.ray_unit[r * 3 + 2]};
  const float far = a.far[r];
  const int64_t base = r * a.N;

  // forward quantities again: alphas (with intermediates), transmittance, composited channels
  mli_alpha_t al[SPL];
  float sdfv[SPL];
  float prod = 1.0f;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const int i = lane * SPL + k;
    al[k].alpha = 0.0f;
    if (i < a.N) {
      const float d0 = a.dists[r * a.ld_d + i];
      const float d1 = (i + 1 < a.N) ? a.dists[r * a.ld_d + i + 1] : far;
      const float g[3] = {a.gradients[(base + i) * 3], a.gradients[(base + i) * 3 + 1], a.gradients[(base + i) * 3 + 2]};
      sdfv[k] = a.sdf[base + i];
      al[k] = mli_neus_alpha(sdfv[k], g, rv, d1 - d0, inv_s, anneal);
      prod *= 1.0f - al[k].alpha;
    }
  }
  const float T0 = warp_excl_prod(prod, lane);
  float acc[kMaxCh], opa = 0.0f;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const int i = lane * SPL + k;
    if (i < a.N) {
      const float w = weights[base + i];
      opa += w;
      const float* s = a.S + (base + i) * a.lds;
#pragma unroll
      for (int c = 0; c < kMaxCh; ++c)
        if (c < a.n_ch) acc[c] = fmaf(w, s[c], acc[c]);
    }
  }
  opa = warp_sum(opa);
  const float white = a.white_bg ? 1.0f - opa : 0.0f;
  float comp[kMaxCh], d_comp[kMaxCh], d_o[12];
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c) { comp[c] = (c < a.n_ch ? warp_sum(acc[c]) : 0.0f) + white; d_comp[c] = 0.0f; }
  const int n_out = n_out_of_mode(a.mode);
  for (int c = 0; c < 12; ++c) d_o[c] = c < n_out ? d_out[r * n_out + c] : 0.0f;
  mode_merge_bwd(a.mode, comp, d_o, d_comp);
  float d_comp_sum = 0.0f;
#pragma unroll
  for (int c = 0; c < kMaxCh; ++c)
    if (c < a.n_ch) d_comp_sum += d_comp[c];

  // per-sample: d(pre-activation head output) and d(weight)
  float dw[SPL];
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    const int i = lane * SPL + k;
    dw[k] = 0.0f;
    if (i < a.N) {
      const float w = weights[base + i];
      const float* s = a.S + (base + i) * a.lds;
      float* ds = dS_pre + (base + i) * a.lds;
      float t = 0.0f;
#pragma unroll
      for (int c = 0; c < kMaxCh; ++c) {
        if (c < a.n_ch) {
          const float sv = s[c];
          t = fmaf(d_comp[c], sv, t);
          ds[c] = w * d_comp[c] * (((act_mask >> c) & 1u) ? sv * (1.0f - sv) : 1.0f);
        }
      }
      for (int c = a.n_ch; c < a.lds; ++c) ds[c] = 0.0f;
      // out_c = sum_i w_i s_ic + white*(1 - sum_i w_i)
      dw[k] = t - (a.white_bg ? d_comp_sum : 0.0f) + (d_weights ? d_weights[base + i] : 0.0f);
    }
  }

  // suffix recurrence S_{i-1} = B_i + A_i S_i with A_i = 1-a_i, B_i = dw_i a_i  (affine maps, composed right to left)
  float GA = 1.0f, GB = 0.0f;  // lane-local map S_hi -> S_{lo-1}: G = f_lo o ... o f_hi
#pragma unroll
  for (int k = SPL - 1; k >= 0; --k) {
    const int i = lane * SPL + k;
    if (i < a.N) {
      // G_new = G_old o f_i  is wrong order; we need f_lo o (... o f_hi): build from hi down: G <- f_i o G
      const float A = 1.0f - al[k].alpha, B = dw[k] * al[k].alpha;
      GB = fmaf(A, GB, B);
      GA = A * GA;
    }
  }
  // inclusive suffix scan over lanes: H_l = G_l o G_{l+1} o ... o G_31
  float HA = GA, HB = GB;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float tA = __shfl_down_sync(0xffffffffu, HA, o), tB = __shfl_down_sync(0xffffffffu, HB, o);
    if (lane + o < 32) { HB = fmaf(HA, tB, HB); HA = HA * tA; }
  }
  // S at the last sample of this lane = (H_{l+1})(0) = HB of lane l+1
  float Sx = __shfl_down_sync(0xffffffffu, HB, 1);
  if (lane == 31) Sx = 0.0f;

  // walk the lane's samples from high to low; T_i recomputed from the exclusive product
  float Tk[SPL];
  {
    float T = T0;
#pragma unroll
    for (int k = 0; k < SPL; ++k) { Tk[k] = T; T *= 1.0f - al[k].alpha; }
  }
  float d_inv_s = 0.0f;
#pragma unroll
  for (int k = SPL - 1; k >= 0; --k) {
    const int i = lane * SPL + k;
    if (i < a.N) {
      const float d_alpha = Tk[k] * (dw[k] - Sx);
      Sx = fmaf(1.0f - al[k].alpha, Sx, dw[k] * al[k].alpha);
      float ds = 0.0f, dg[3] = {0.f, 0.f, 0.f};
      d_inv_s += mli_neus_alpha_bwd(al[k], sdfv[k], rv, inv_s, anneal, d_alpha, &ds, dg);
      d_sdf[base + i] = ds;
      d_gradients[(base + i) * 3 + 0] += dg[0];
      d_gradients[(base + i) * 3 + 1] += dg[1];
      d_gradients[(base + i) * 3 + 2] += dg[2];
    }
  }
  d_inv_s = warp_sum(d_inv_s);
  if (lane == 0 && d_inv_s_partial) d_inv_s_partial[r] = d_inv_s * inv_s;  // inv_s = exp(s_var)
}

__global__ void sum_partials_kernel(const float* __restrict__ part, int64_t n, float* __restrict__ out, int accumulate) {
  __shared__ float red[32];
  float v = 0.0f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v += part[i];  // fixed assignment: deterministic
  v = mli_block_sum(v, red);
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.0f) + v;
}

int make_args(CompArgs* a, const mli_composite_cfg_t* cfg, const float* s_var, const float* sdf, const float* gradients,
              const float* ray_unit, const float* dists, int64_t ld_d, const float* far, const float* S, int64_t lds,
              int64_t R) {
  MLI_REQUIRE(cfg != nullptr, "composite: cfg is NULL");
  MLI_REQUIRE(cfg->N >= 1 && cfg->N <= 32 * kMaxSPL, "composite: N must be in 1..%d", 32 * kMaxSPL);
  MLI_REQUIRE(cfg->mode >= MLI_MODE_RGB && cfg->mode <= MLI_MODE_R_S_RE, "composite: unknown network_mode %d", cfg->mode);
  const int n_ch_mode[5] = {3, 7, 6, 6, 9};
  MLI_REQUIRE(lds >= n_ch_mode[cfg->mode] && ld_d >= cfg->N, "composite: bad lds/ld_d");
  a->s_var = s_var; a->sdf = sdf; a->gradients = gradients; a->ray_unit = ray_unit; a->dists = dists; a->ld_d = ld_d;
  a->far = far; a->S = S; a->lds = lds; a->R = R; a->N = cfg->N; a->n_ch = n_ch_mode[cfg->mode]; a->mode = cfg->mode;
  a->white_bg = cfg->white_bg; a->eval_extras = cfg->eval_extras; a->anneal = cfg->anneal_ratio; a->anneal_dev = cfg->anneal_dev;
  return MLI_OK;
}

}  // namespace

#define DISPATCH_SPL(N, CALL)                                  \
  do {                                                         \
    const int spl = ((N) + 31) / 32;                           \
    if (spl <= 1) { constexpr int SPL = 1; CALL; }             \
    else if (spl <= 2) { constexpr int SPL = 2; CALL; }        \
    else if (spl <= 4) { constexpr int SPL = 4; CALL; }        \
    else { constexpr int SPL = 8; CALL; }                      \
  } while (0)

extern "C" int mli_composite_fwd(const mli_composite_cfg_t* cfg, const float* s_var, const float* sdf_center,
                                 const float* gradients, const float* ray_unit, const float* dists, int64_t ld_d,
                                 const float* far, const float* S, int64_t lds, int64_t R, float* alphas,
                                 float* weights, float* out, float* extras, void* stream) {
  MLI_ENTRY();
  CompArgs a;
  if (int e = make_args(&a, cfg, s_var, sdf_center, gradients, ray_unit, dists, ld_d, far, S, lds, R)) return e;
  if (R <= 0) return MLI_OK;
  DISPATCH_SPL(cfg->N, (composite_fwd_kernel<SPL><<<mli_cdiv(R, kWarps), 32 * kWarps, 0, (cudaStream_t)stream>>>(
                           a, alphas, weights, out, extras)));
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_composite_bwd(const mli_composite_cfg_t* cfg, const float* s_var, const float* sdf_center,
                                 const float* gradients, const float* ray_unit, const float* dists, int64_t ld_d,
                                 const float* far, const float* S, int64_t lds, int64_t R, const float* weights,
                                 const float* d_out, const float* d_weights, uint32_t act_mask, float* dS_pre,
                                 float* d_sdf_center, float* d_gradients, float* d_s_var, int32_t accumulate_s_var,
                                 void* ws, void* stream) {
  MLI_ENTRY();
  CompArgs a;
  if (int e = make_args(&a, cfg, s_var, sdf_center, gradients, ray_unit, dists, ld_d, far, S, lds, R)) return e;
  MLI_REQUIRE(d_s_var == nullptr || ws != nullptr, "composite_bwd: workspace (R floats) required for d_s_var");
  if (R <= 0) return MLI_OK;
  DISPATCH_SPL(cfg->N, (composite_bwd_kernel<SPL><<<mli_cdiv(R, kWarps), 32 * kWarps, 0, (cudaStream_t)stream>>>(
                           a, weights, d_out, d_weights, act_mask, dS_pre, d_sdf_center, d_gradients,
                           d_s_var ? (float*)ws : nullptr)));
  MLI_LAUNCH_OK();
  if (d_s_var) {
    sum_partials_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const float*)ws, R, d_s_var, accumulate_s_var);
    MLI_LAUNCH_OK();
  }
  return MLI_OK;
}

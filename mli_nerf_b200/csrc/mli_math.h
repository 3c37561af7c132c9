// mli_math.h -- scalar/per-ray math of the render path, shared by every kernel.
//
// Everything here is __host__ __device__ so the exact arithmetic the kernels run can also be compiled with
// g++ into tests/hostsim (a TEST-ONLY harness that checks these formulas against the oracle on a box with
// no GPU).  The product never runs them on the host: the only callers in libmli_b200.so are __global__
// kernels.  Reference citations are relative to /root/reference/.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/mli_b200.h"

#if defined(__CUDACC__)
#define MLI_HD __host__ __device__ __forceinline__
#else
#define MLI_HD inline
#endif

// Non-contracting fp32 ops: torch evaluates a*b+c as two rounded ops; where bit-compatibility of a later
// integer decision depends on it we must not let nvcc fuse them into an FMA.
#if defined(__CUDA_ARCH__)
MLI_HD float mli_mul(float a, float b) { return __fmul_rn(a, b); }
MLI_HD float mli_add(float a, float b) { return __fadd_rn(a, b); }
MLI_HD float mli_sub(float a, float b) { return __fsub_rn(a, b); }
MLI_HD float mli_div(float a, float b) { return __fdiv_rn(a, b); }
#else
MLI_HD float mli_mul(float a, float b) { volatile float r = a * b; return r; }
MLI_HD float mli_add(float a, float b) { volatile float r = a + b; return r; }
MLI_HD float mli_sub(float a, float b) { volatile float r = a - b; return r; }
MLI_HD float mli_div(float a, float b) { volatile float r = a / b; return r; }
#endif

// ---------------------------------------------------------------------------------------------------------
// activations (forward from pre-activation; derivative from the *output*, so nothing extra is stored)
// ---------------------------------------------------------------------------------------------------------
MLI_HD float mli_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100   (misc.py:92-107, mlp cfg)
MLI_HD float mli_softplus100(float x) {
  float bx = 100.0f * x;
  return bx > 20.0f ? x : log1pf(expf(bx)) * 0.01f;
}
// d softplus/dx = sigmoid(100x) = 1 - exp(-100 y); beyond the threshold exactly 1.
MLI_HD float mli_softplus100_grad_from_out(float y) {
  return y > 0.2f ? 1.0f : -expm1f(-100.0f * y);
}
MLI_HD float mli_act(float x, int act) {
  switch (act) {
    case MLI_ACT_RELU: return x > 0.0f ? x : 0.0f;
    case MLI_ACT_SOFTPLUS100: return mli_softplus100(x);
    case MLI_ACT_SIGMOID: return mli_sigmoid(x);
    default: return x;
  }
}
MLI_HD float mli_dact_from_out(float y, int act) {
  switch (act) {
    case MLI_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
    case MLI_ACT_SOFTPLUS100: return mli_softplus100_grad_from_out(y);
    case MLI_ACT_SIGMOID: return y * (1.0f - y);
    default: return 1.0f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// hash grid (tcnn grid.h semantics, see oracle/torch_hashgrid.py)
// ---------------------------------------------------------------------------------------------------------
#define MLI_PRIME_Y 2654435761u
#define MLI_PRIME_Z 805459861u

MLI_HD uint32_t mli_grid_index(const mli_level_t& lv, uint32_t gx, uint32_t gy, uint32_t gz) {
  uint32_t index;
  if (lv.hashed) {
    index = gx ^ (gy * MLI_PRIME_Y) ^ (gz * MLI_PRIME_Z);
  } else {
    // dense walk: index += g[d]*stride while stride <= size
    uint32_t stride = 1;
    index = 0;
    if (stride <= lv.size) { index += gx * stride; stride *= lv.res; }
    if (stride <= lv.size) { index += gy * stride; stride *= lv.res; }
    if (stride <= lv.size) { index += gz * stride; stride *= lv.res; }
  }
  // index % size without the ~20-instruction runtime-divisor modulo in the common cases (level-uniform branches):
  // hashed levels have a power-of-two size (mask), dense indices of in-range points are already < size
  if ((lv.size & (lv.size - 1u)) == 0u) return index & (lv.size - 1u);
  return index < lv.size ? index : index % lv.size;
}

struct mli_cell_t {
  uint32_t g[3];
  float w[3];
};

MLI_HD mli_cell_t mli_grid_cell(const mli_level_t& lv, float x, float y, float z) {
  mli_cell_t c;
  float p[3] = {fmaf(lv.scale, x, 0.5f), fmaf(lv.scale, y, 0.5f), fmaf(lv.scale, z, 0.5f)};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float f = floorf(p[d]);
    c.g[d] = (uint32_t)(int)f;
    c.w[d] = p[d] - f;
  }
  return c;
}

MLI_HD void mli_corner(const mli_level_t& lv, const mli_cell_t& c, int corner, uint32_t* row, float* weight) {
  float w = 1.0f;
  uint32_t g[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if ((corner >> d) & 1) { w *= c.w[d]; g[d] = c.g[d] + 1u; }
    else { w *= 1.0f - c.w[d]; g[d] = c.g[d]; }
  }
  *row = lv.offset + mli_grid_index(lv, g[0], g[1], g[2]);
  *weight = w;
}

// sample point + tap offset exactly as torch does it: p = c + r*d (two rounded ops), then p + k*e.
// plane 0 = centre; 4 taps: k in {(1,-1,-1),(-1,-1,1),(-1,1,-1),(1,1,1)} (modules.py:159-162);
// 6 taps: +x,-x,+y,-y,+z,-z (modules.py:135-143).
MLI_HD void mli_tap_sign(int taps, int plane, float* k) {
  k[0] = k[1] = k[2] = 0.0f;
  if (plane == 0) return;
  if (taps == 4) {  // planes 1..4 = k1..k4
    k[0] = (plane == 1 || plane == 4) ? 1.0f : -1.0f;
    k[1] = (plane >= 3) ? 1.0f : -1.0f;
    k[2] = (plane == 2 || plane == 4) ? 1.0f : -1.0f;
  } else {
    int axis = (plane - 1) >> 1;
    k[axis] = ((plane - 1) & 1) ? -1.0f : 1.0f;
  }
}

MLI_HD void mli_sample_point(const float* c, const float* r, float d, int taps, int plane, float tap_eps, float* p) {
  float k[3];
  mli_tap_sign(taps, plane, k);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float v = mli_add(c[a], mli_mul(r[a], d));
    if (plane) v = mli_add(v, mli_mul(k[a], tap_eps));
    p[a] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// real spherical harmonics, 16 coefficients (spherical_harmonics.py:47-70)
// ---------------------------------------------------------------------------------------------------------
MLI_HD void mli_sh16(float x, float y, float z, float* v) {
  const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
  float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  v[0] = C0;
  v[1] = -C1 * y; v[2] = C1 * z; v[3] = -C1 * x;
  v[4] = 1.0925484305920792f * xy;
  v[5] = -1.0925484305920792f * yz;
  v[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
  v[7] = -1.0925484305920792f * xz;
  v[8] = 0.5462742152960396f * (xx - yy);
  v[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
  v[10] = 2.890611442640554f * xy * z;
  v[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
  v[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
  v[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
  v[14] = 1.445305721320277f * z * (xx - yy);
  v[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// ---------------------------------------------------------------------------------------------------------
// bounds (neuralangelo/model.py:420-430; nerf_util.py:199-205; NeuralLumen/utils/utils.py:86-123)
// ---------------------------------------------------------------------------------------------------------
MLI_HD void mli_bounds_sphere(const float* c, const float* r, float* near, float* far, uint8_t* outside) {
  // torch: (c*c).sum(-1): sequential over 3 elements, products rounded individually
  float ctc = mli_add(mli_add(mli_mul(c[0], c[0]), mli_mul(c[1], c[1])), mli_mul(c[2], c[2]));
  float ctv = mli_add(mli_add(mli_mul(c[0], r[0]), mli_mul(c[1], r[1])), mli_mul(c[2], r[2]));
  float disc = mli_sub(mli_mul(ctv, ctv), mli_sub(ctc, 1.0f));
  float sq = sqrtf(disc);  // NaN when the ray misses
  float n = mli_sub(-ctv, sq), f = mli_add(-ctv, sq);
  n = (n != n) ? n : (n > 0.0f ? n : 0.0f);  // relu_ keeps NaN
  bool out = n != n;
  *near = out ? 1.0f : n;
  *far = out ? 1.2f : f;
  *outside = out ? 1 : 0;
}

MLI_HD void mli_bounds_aabb(const float* c, const float* r, const float* aabb, float* near, float* far,
                            uint8_t* outside) {
  float tmin = -INFINITY, tmax = INFINITY;
  bool nan_seen = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float t0 = mli_div(mli_sub(aabb[a], c[a]), r[a]);
    float t1 = mli_div(mli_sub(aabb[3 + a], c[a]), r[a]);
    // torch amin/amax propagate NaN
    float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
    if (t0 != t0 || t1 != t1) nan_seen = true;
    tmin = fmaxf(tmin, lo);
    tmax = fminf(tmax, hi);
  }
  if (nan_seen) { tmin = NAN; tmax = NAN; }
  // clamp(min=0,max=1e10) propagates NaN
  if (tmin == tmin) tmin = fminf(fmaxf(tmin, 0.0f), 1e10f);
  if (tmax == tmax) tmax = fminf(fmaxf(tmax, 0.0f), 1e10f);
  bool out = tmax <= tmin;  // false for NaN, like torch
  *near = out ? 1.0f : tmin;
  *far = out ? 1.2f : tmax;
  *outside = out ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// hierarchical resampling, one ray (neuralangelo/model.py:467-484; nerf_util.py:41-68; render.py:87-99)
//
// torch-CPU rounding behaviour mirrored here (measured on torch 2.11 CPU, SURVEY.md section 7 "hard parts"):
//   cumprod / cumsum : sequential, float64 accumulator, each output rounded to float32
//   F.normalize(p=1) : denominator = sequential float32 sum of |w|, clamped at 1e-12
//   searchsorted(right=True) : upper bound
// `scratch` needs n floats (the cdf, n = number of input samples; cdf[0] = 0).
// ---------------------------------------------------------------------------------------------------------
MLI_HD void mli_weights_to_cdf(const float* w, int n_w, float* cdf) {
  float denom = 0.0f;
  for (int i = 0; i < n_w; ++i) denom = mli_add(denom, fabsf(w[i]));
  denom = fmaxf(denom, 1e-12f);
  double acc = 0.0;
  cdf[0] = 0.0f;
  for (int i = 0; i < n_w; ++i) {
    acc += (double)mli_div(w[i], denom);
    cdf[i + 1] = (float)acc;
  }
}

MLI_HD int mli_upper_bound(const float* a, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// unif_j = 0.5*(grid[j]+grid[j+1]) with grid = linspace(0,1,n_fine+1) in float32 (nerf_util.py:55-56)
MLI_HD float mli_linspace01(int j, int steps) {
  // torch.linspace float32 CPU kernel: step = (end-start)/(steps-1); value = start + step*j for j < steps/2,
  // end - step*(steps-1-j) otherwise.
  float step = 1.0f / (float)(steps - 1);
  if (j < steps / 2) return mli_mul(step, (float)j);
  return mli_sub(1.0f, mli_mul(step, (float)(steps - 1 - j)));
}
MLI_HD float mli_unif(int j, int n_fine) {
  return mli_mul(0.5f, mli_add(mli_linspace01(j, n_fine + 1), mli_linspace01(j + 1, n_fine + 1)));
}

MLI_HD float mli_sample_bin(const float* dists, const float* cdf, int n, float u, int* idx_o, int* low_o, int* high_o) {
  int idx = mli_upper_bound(cdf, n, u);
  int low = idx - 1 < 0 ? 0 : idx - 1;
  int high = idx > n - 1 ? n - 1 : idx;
  float c_lo = cdf[low], c_hi = cdf[high];
  float t = mli_div(mli_sub(u, c_lo), mli_add(mli_sub(c_hi, c_lo), 1e-8f));
  float d_lo = dists[low], d_hi = dists[high];
  if (idx_o) { *idx_o = idx; *low_o = low; *high_o = high; }
  return mli_add(d_lo, mli_mul(t, mli_sub(d_hi, d_lo)));
}

// finite-difference cosine of interval i (between samples i and i+1)
MLI_HD float mli_hier_cos(const float* d, const float* s, int i) {
  return mli_div(mli_sub(s[i + 1], s[i]), mli_add(mli_sub(d[i + 1], d[i]), 1e-5f));
}
// alpha of interval i for the hierarchical pass (robust=True: min with the previous interval's cosine)
MLI_HD float mli_hier_alpha(const float* d, const float* s, int i, float inv_s) {
  float mid = mli_mul(mli_add(s[i], s[i + 1]), 0.5f);
  float intv = mli_sub(d[i + 1], d[i]);
  float cosv = mli_hier_cos(d, s, i);
  float prev_cos = i > 0 ? mli_hier_cos(d, s, i - 1) : 0.0f;
  float c = fminf(prev_cos, cosv);
  if (prev_cos != prev_cos || cosv != cosv) c = NAN;  // torch.min propagates NaN
  float half = mli_mul(mli_mul(c, intv), 0.5f);
  float p = mli_sigmoid(mli_mul(mli_sub(mid, half), inv_s));
  float q = mli_sigmoid(mli_mul(mli_add(mid, half), inv_s));
  float a = mli_div(mli_sub(p, q), mli_add(p, 1e-5f));
  return fminf(fmaxf(a, 0.0f), 1.0f);
}
// in place: alphas a[0..n_w) -> compositing weights (float64 cumprod accumulator, float32 outputs)
MLI_HD void mli_alphas_to_weights(float* a, int n_w) {
  double T = 1.0;
  for (int i = 0; i < n_w; ++i) {
    float ai = a[i];
    a[i] = mli_mul(ai, (float)T);
    T *= (double)mli_sub(1.0f, ai);
  }
}
// weights of the n-1 intervals of one ray for the hierarchical pass -> w[0..n-2]
MLI_HD void mli_hier_weights(const float* d, const float* s, int n, float inv_s, float* w) {
  for (int i = 0; i < n - 1; ++i) w[i] = mli_hier_alpha(d, s, i, inv_s);
  mli_alphas_to_weights(w, n - 1);
}

// ---------------------------------------------------------------------------------------------------------
// NeuS alpha for one sample and its backward (neuralangelo/model.py:492-515)
// ---------------------------------------------------------------------------------------------------------
struct mli_alpha_t {
  float alpha, p, q, raw, iter_cos, true_cos, intv;
};

MLI_HD mli_alpha_t mli_neus_alpha(float sdf, const float* g, const float* r, float intv, float inv_s, float anneal) {
  mli_alpha_t o;
  o.true_cos = r[0] * g[0] + r[1] * g[1] + r[2] * g[2];
  float a1 = fmaxf(-o.true_cos * 0.5f + 0.5f, 0.0f), a2 = fmaxf(-o.true_cos, 0.0f);
  o.iter_cos = -(a1 * (1.0f - anneal) + a2 * anneal);
  o.intv = intv;
  float half = o.iter_cos * intv * 0.5f;
  o.p = mli_sigmoid((sdf - half) * inv_s);
  o.q = mli_sigmoid((sdf + half) * inv_s);
  o.raw = (o.p - o.q) / (o.p + 1e-5f);
  o.alpha = fminf(fmaxf(o.raw, 0.0f), 1.0f);
  return o;
}

// d_alpha -> d_sdf, d_g[3] (added), d_inv_s (returned)
MLI_HD float mli_neus_alpha_bwd(const mli_alpha_t& o, float sdf, const float* r, float inv_s, float anneal,
                                float d_alpha, float* d_sdf, float* d_g) {
  // clip passes gradient on the closed interval [0,1]
  float d_raw = (o.raw >= 0.0f && o.raw <= 1.0f) ? d_alpha : 0.0f;
  float den = o.p + 1e-5f;
  float d_p = d_raw * (1.0f / den - (o.p - o.q) / (den * den));
  float d_q = -d_raw / den;
  float d_zp = d_p * o.p * (1.0f - o.p), d_zq = d_q * o.q * (1.0f - o.q);  // z = (sdf -/+ half) * inv_s
  float half = o.iter_cos * o.intv * 0.5f;
  float d_inv_s = d_zp * (sdf - half) + d_zq * (sdf + half);
  float d_ep = d_zp * inv_s, d_eq = d_zq * inv_s;
  *d_sdf = d_ep + d_eq;
  float d_half = d_eq - d_ep;
  float d_iter = d_half * o.intv * 0.5f;
  // iter_cos = -(relu(-c/2+1/2)(1-a) + relu(-c) a); relu' = 0 at 0
  float d_c = 0.0f;
  if (-o.true_cos * 0.5f + 0.5f > 0.0f) d_c += 0.5f * (1.0f - anneal);
  if (-o.true_cos > 0.0f) d_c += anneal;
  float d_true = d_iter * d_c;
  d_g[0] += d_true * r[0]; d_g[1] += d_true * r[1]; d_g[2] += d_true * r[2];
  return d_inv_s;
}

// geometry.cu -- per-sample SDF stencil -> gradient / Hessian diagonal / normal, and the non-feature part of the
// head inputs (points, SH of the view direction, normal, SH of the light *position*).
//
// Reference (relative to /root/reference/): projects/neuralangelo/utils/modules.py:131-177 (numerical 4/6 taps),
// projects/NeuralLumen/model.py:343-349 (outside overwrite before the stencil, normalize, expand),
// projects/neuralangelo/utils/spherical_harmonics.py:47-70, projects/NeuralLumen/utils/modules.py:106-109.
// Elementwise, one thread per sample; HBM-bandwidth bound (~(1+taps)*4 B in, ~220 B out per sample).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

struct GeoConst {
  int N, taps;
  float div_grad;   // float32(4e) (4 taps) or float32(2 eps) (6 taps)
  float div_hess;   // float32(e^2) / float32(eps^2)
  float outside_val;
};

__global__ void __launch_bounds__(256) geometry_fwd_kernel(float* __restrict__ sdf, int64_t M, GeoConst g,
                                                           const uint8_t* __restrict__ outside,
                                                           const float* __restrict__ center,
                                                           const float* __restrict__ ray_unit,
                                                           const float* __restrict__ pts_light,
                                                           const float* __restrict__ dists, int64_t ld_d,
                                                           float* __restrict__ gradients, float* __restrict__ hessians,
                                                           float* __restrict__ XH, int64_t ldxh, int xh_off,
                                                           int sdf_is_delta, __nv_bfloat16* __restrict__ xh_tcl,
                                                           int xh_chunks, int xh_chunk0) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t ray = m / g.N;
  const int i = (int)(m - ray * g.N);
  float s0 = sdf[m];
  // delta mode: the tap planes hold d_i = sdf_i - sdf_centre.  Gradient and Hessian only depend on sdf_i - s0 (the
  // tap vectors sum to zero), so evaluate the reference's formulas on (d_i + shift, 0 + shift') with the centre
  // moved to the origin; `shift` restores the literal behaviour for outside rays, whose centre alone is overwritten.
  float shift = 0.0f;
  if (outside[ray]) {  // model.py:343 (in place, before the stencil)
    if (sdf_is_delta) shift = s0 - g.outside_val;
    s0 = g.outside_val;
    sdf[m] = s0;
  }
  if (sdf_is_delta) s0 = 0.0f;
  float grad[3], hess[3];
  if (g.taps == 4) {
    const float s1 = sdf[M + m] + shift, s2 = sdf[2 * M + m] + shift, s3 = sdf[3 * M + m] + shift, s4 = sdf[4 * M + m] + shift;
    // (k1*s1 + k2*s2 + k3*s3 + k4*s4) / (4e), summed left to right (modules.py:167)
    grad[0] = mli_div(mli_add(mli_add(mli_add(s1, -s2), -s3), s4), g.div_grad);
    grad[1] = mli_div(mli_add(mli_add(mli_add(-s1, -s2), s3), s4), g.div_grad);
    grad[2] = mli_div(mli_add(mli_add(mli_add(-s1, s2), -s3), s4), g.div_grad);
    // ((s1+s2+s3+s4)/2 - 2 sdf)/e^2, then [h,h,h]/3 (modules.py:172-173)
    const float sum = mli_add(mli_add(mli_add(s1, s2), s3), s4);
    const float hxx = mli_div(mli_sub(mli_div(sum, 2.0f), mli_mul(2.0f, s0)), g.div_hess);
    hess[0] = hess[1] = hess[2] = mli_div(hxx, 3.0f);
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float sp = sdf[(1 + 2 * a) * M + m] + shift, sn = sdf[(2 + 2 * a) * M + m] + shift;
      grad[a] = mli_div(mli_sub(sp, sn), g.div_grad);
      hess[a] = mli_div(mli_sub(mli_add(sp, sn), mli_mul(2.0f, s0)), g.div_hess);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) gradients[m * 3 + a] = grad[a];
  if (hessians) {
#pragma unroll
    for (int a = 0; a < 3; ++a) hessians[m * 3 + a] = hess[a];
  }
  if (XH || xh_tcl) {
    const float c[3] = {center[ray * 3], center[ray * 3 + 1], center[ray * 3 + 2]};
    const float r[3] = {ray_unit[ray * 3], ray_unit[ray * 3 + 1], ray_unit[ray * 3 + 2]};
    float p[3];
    mli_sample_point(c, r, dists[ray * ld_d + i], 0, 0, 0.0f, p);
    const float nrm = sqrtf(grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2]);
    const float den = fmaxf(nrm, 1e-12f);  // F.normalize eps
    float row[48];
    row[0] = p[0]; row[1] = p[1]; row[2] = p[2];
    mli_sh16(r[0], r[1], r[2], row + 3);
    row[19] = grad[0] / den; row[20] = grad[1] / den; row[21] = grad[2] / den;
    mli_sh16(pts_light[ray * 3], pts_light[ray * 3 + 1], pts_light[ray * 3 + 2], row + 22);
#pragma unroll
    for (int k = 38; k < 48; ++k) row[k] = 0.0f;
    if (XH) {
      float* dst = XH + m * ldxh + xh_off;
#pragma unroll
      for (int k = 0; k < 38; ++k) dst[k] = row[k];
      for (int k = xh_off + 38; k < ldxh; ++k) XH[m * ldxh + k] = 0.0f;
    }
    if (xh_tcl) {  // bf16 TCL chunks [xh_chunk0, xh_chunk0 + 6) of the head-input matrix: 16 B per thread and chunk
      const int64_t tile = m >> 7;
      const int rl = (int)(m & 127);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(row[j * 8 + 2 * q], row[j * 8 + 2 * q + 1]);
          w[q] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(xh_tcl + ((tile * xh_chunks + xh_chunk0 + j) * 128 + rl) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) geometry_bwd_kernel(const float* __restrict__ gradients, int64_t M, GeoConst g,
                                                           const uint8_t* __restrict__ outside,
                                                           const float* __restrict__ d_grad_in,
                                                           const float* __restrict__ d_hess_in,
                                                           const float* __restrict__ dXH, int64_t ldxh, int xh_off,
                                                           const float* __restrict__ d_sdf_center_in,
                                                           float* __restrict__ d_sdf) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t ray = m / g.N;
  float dg[3] = {0.f, 0.f, 0.f}, dh[3] = {0.f, 0.f, 0.f};
  if (d_grad_in) { dg[0] = d_grad_in[m * 3]; dg[1] = d_grad_in[m * 3 + 1]; dg[2] = d_grad_in[m * 3 + 2]; }
  if (d_hess_in) { dh[0] = d_hess_in[m * 3]; dh[1] = d_hess_in[m * 3 + 1]; dh[2] = d_hess_in[m * 3 + 2]; }
  if (dXH) {  // normals = g / max(|g|, 1e-12)
    const float gx = gradients[m * 3], gy = gradients[m * 3 + 1], gz = gradients[m * 3 + 2];
    const float dn[3] = {dXH[m * ldxh + xh_off + 19], dXH[m * ldxh + xh_off + 20], dXH[m * ldxh + xh_off + 21]};
    const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
    if (nrm > 1e-12f) {
      const float inv = 1.0f / nrm;
      const float n[3] = {gx * inv, gy * inv, gz * inv};
      const float dot = n[0] * dn[0] + n[1] * dn[1] + n[2] * dn[2];
#pragma unroll
      for (int a = 0; a < 3; ++a) dg[a] += (dn[a] - n[a] * dot) * inv;
    } else {  // clamped denominator carries no gradient
#pragma unroll
      for (int a = 0; a < 3; ++a) dg[a] += dn[a] * 1e12f;
    }
  }
  float d0;
  if (g.taps == 4) {
    const float ig = 1.0f / g.div_grad;
    const float dhx = (dh[0] + dh[1] + dh[2]) / 3.0f / g.div_hess;  // d hxx
    const float a1 = dg[0] * ig, a2 = dg[1] * ig, a3 = dg[2] * ig;
    d_sdf[M + m] = a1 - a2 - a3 + 0.5f * dhx;      // k1 = ( 1,-1,-1)
    d_sdf[2 * M + m] = -a1 - a2 + a3 + 0.5f * dhx; // k2 = (-1,-1, 1)
    d_sdf[3 * M + m] = -a1 + a2 - a3 + 0.5f * dhx; // k3 = (-1, 1,-1)
    d_sdf[4 * M + m] = a1 + a2 + a3 + 0.5f * dhx;  // k4 = ( 1, 1, 1)
    d0 = -2.0f * dhx;
  } else {
    d0 = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gg = dg[a] / g.div_grad, hh = dh[a] / g.div_hess;
      d_sdf[(1 + 2 * a) * M + m] = gg + hh;
      d_sdf[(2 + 2 * a) * M + m] = -gg + hh;
      d0 -= 2.0f * hh;
    }
  }
  if (d_sdf_center_in) d0 += d_sdf_center_in[m];
  // the in-place overwrite sdfs[outside] = 1000 cuts the graph for those samples (model.py:343)
  d_sdf[m] = outside[ray] ? 0.0f : d0;
}

int make_const(GeoConst* g, int32_t N, int32_t taps, double tap_eps, float outside_val) {
  MLI_REQUIRE(taps == 4 || taps == 6, "Only support 4 or 6 taps.");
  MLI_REQUIRE(N >= 1 && tap_eps > 0.0, "geometry: bad N/tap_eps");
  g->N = N; g->taps = taps; g->outside_val = outside_val;
  g->div_grad = (float)((taps == 4 ? 4.0 : 2.0) * tap_eps);
  g->div_hess = (float)(tap_eps * tap_eps);
  return MLI_OK;
}

}  // namespace

extern "C" int mli_geometry_fwd(float* sdf, int64_t M, int32_t N, int32_t taps, double tap_eps, const uint8_t* outside,
                                float outside_val, const float* center, const float* ray_unit, const float* pts_light,
                                const float* dists, int64_t ld_d, float* gradients, float* hessians, float* XH,
                                int64_t ldxh, int32_t xh_off, int32_t sdf_is_delta, void* xh_tcl, int32_t xh_chunks,
                                int32_t xh_chunk0, void* stream) {
  MLI_ENTRY();
  GeoConst g;
  if (int e = make_const(&g, N, taps, tap_eps, outside_val)) return e;
  MLI_REQUIRE(M >= 0 && M % N == 0, "geometry: M must be a multiple of N");
  MLI_REQUIRE(XH == nullptr || ldxh >= xh_off + 38, "geometry: XH row too short");
  MLI_REQUIRE(xh_tcl == nullptr || (xh_chunk0 >= 0 && xh_chunk0 + 6 <= xh_chunks && M % 128 == 0), "geometry: bad TCL chunk range");
  if (M == 0) return MLI_OK;
  geometry_fwd_kernel<<<mli_cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(sdf, M, g, outside, center, ray_unit, pts_light,
                                                                         dists, ld_d, gradients, hessians, XH, ldxh, xh_off, sdf_is_delta,
                                                                         (__nv_bfloat16*)xh_tcl, xh_chunks, xh_chunk0);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_geometry_bwd(const float* gradients, int64_t M, int32_t N, int32_t taps, double tap_eps,
                                const uint8_t* outside, const float* d_grad_in, const float* d_hess_in, const float* dXH,
                                int64_t ldxh, int32_t xh_off, const float* d_sdf_center_in, float* d_sdf, void* stream) {
  MLI_ENTRY();
  GeoConst g;
  if (int e = make_const(&g, N, taps, tap_eps, 0.0f)) return e;
  MLI_REQUIRE(M >= 0 && M % N == 0, "geometry: M must be a multiple of N");
  if (M == 0) return MLI_OK;
  geometry_bwd_kernel<<<mli_cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(gradients, M, g, outside, d_grad_in, d_hess_in,
                                                                         dXH, ldxh, xh_off, d_sdf_center_in, d_sdf);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

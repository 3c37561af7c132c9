// visibility.cu -- light-visibility estimation by sphere tracing (the stage-a export
// `test_all_light` that produces the pseudo shading labels of stage b).
//
// Reference (relative to /root/reference/): projects/NeuralLumen/model.py:133-200 (get_light_visibility,
// get_dist_bounds_visibility), projects/neuralangelo/model.py:298-325 (sphere_tracing_intersection, borrowed from
// L-Tracing with the near/far masking fix).  The SDF queries of the 20 + 20 marching iterations run through the
// encode + SDF-trunk kernels of the sampling path; the kernels here are the per-ray state updates around them.
// Elementwise, one thread per ray.
#include "common.cuh"

namespace {

// one marching iteration: dist[mask] += sdf[mask]; mask[dist > far] = False; mask[dist < near] = False
// (neuralangelo/model.py:311-322); last != 0 additionally applies the final clamp(dist, near, far) (:323)
__global__ void sphere_trace_step_kernel(float* __restrict__ dist, uint8_t* __restrict__ mask, const float* __restrict__ sdf,
                                         const float* __restrict__ near, const float* __restrict__ far, int64_t R, int last) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float d = dist[r];
  uint8_t m = mask[r];
  if (m) d = mli_add(d, sdf[r]);
  if (d > far[r]) m = 0;
  if (d < near[r]) m = 0;
  if (last) d = fminf(fmaxf(d, near[r]), far[r]);  // torch.clamp(dist, near, far)
  dist[r] = d;
  mask[r] = m;
}

struct Bound { float radius; float aabb[6]; int use_box; };

// light ray through the camera-ray intersection + its marching interval (NeuralLumen/model.py:149-170,186-200)
__global__ void light_rays_kernel(const float* __restrict__ center, const float* __restrict__ ray_unit,
                                  const float* __restrict__ inter_dist, const float* __restrict__ pts_light, int64_t R,
                                  Bound b, float* __restrict__ light_unit, float* __restrict__ near_l,
                                  float* __restrict__ far_tracing, uint8_t* __restrict__ inside_bounding) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float lr[3], lo[3], u[3];
  const float d = inter_dist[r];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    lo[k] = pts_light[r * 3 + k];
    const float inter = mli_add(center[r * 3 + k], mli_mul(ray_unit[r * 3 + k], d));  // center + ray_unit * dist
    lr[k] = mli_sub(inter, lo[k]);
  }
  const float len = sqrtf(mli_add(mli_add(mli_mul(lr[0], lr[0]), mli_mul(lr[1], lr[1])), mli_mul(lr[2], lr[2])));
  const float den = fmaxf(len, 1e-12f);  // F.normalize eps
#pragma unroll
  for (int k = 0; k < 3; ++k) { u[k] = mli_div(lr[k], den); light_unit[r * 3 + k] = u[k]; }
  float n, f;
  uint8_t out;
  if (b.use_box) {
    mli_bounds_aabb(lo, u, b.aabb, &n, &f, &out);
  } else {
    // nerf_util.intersect_with_sphere(center, ray_unit, radius) + relu_ / isnan / dummy distances (model.py:192-197)
    const float ctc = mli_add(mli_add(mli_mul(lo[0], lo[0]), mli_mul(lo[1], lo[1])), mli_mul(lo[2], lo[2]));
    const float ctv = mli_add(mli_add(mli_mul(lo[0], u[0]), mli_mul(lo[1], u[1])), mli_mul(lo[2], u[2]));
    const float disc = mli_sub(mli_mul(ctv, ctv), mli_sub(ctc, mli_mul(b.radius, b.radius)));
    const float sq = sqrtf(disc);  // NaN when the light ray misses the sphere
    n = mli_sub(-ctv, sq);
    f = mli_add(-ctv, sq);
    n = (n != n) ? n : (n > 0.0f ? n : 0.0f);
    out = n != n;
    if (out) { n = 1.0f; f = 1.2f; }
  }
  const float ft = mli_sub(len, 1e-3f);  // tolerance for reaching the limit in sphere_tracing_intersection
  near_l[r] = n;
  far_tracing[r] = ft;
  inside_bounding[r] = (n < ft) && (ft < f) && !out;
}

// visibility = ~mask_light | ~inside_bounding; normal_x_light = relu(normalize(-gradient) . light_unit)  (model.py:173-182)
__global__ void light_finish_kernel(const uint8_t* __restrict__ mask_light, const uint8_t* __restrict__ inside_bounding,
                                    const float* __restrict__ gradient, const float* __restrict__ light_unit, int64_t R,
                                    uint8_t* __restrict__ visibility, float* __restrict__ nxl) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  visibility[r] = (!mask_light[r]) || (!inside_bounding[r]);
  const float g[3] = {-gradient[r * 3], -gradient[r * 3 + 1], -gradient[r * 3 + 2]};
  const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  const float den = fmaxf(nrm, 1e-12f);
  const float dot = (g[0] / den) * light_unit[r * 3] + (g[1] / den) * light_unit[r * 3 + 1] + (g[2] / den) * light_unit[r * 3 + 2];
  nxl[r] = dot > 0.0f ? dot : 0.0f;
}

}  // namespace

extern "C" int mli_sphere_trace_step(float* dist, uint8_t* mask, const float* sdf, const float* near, const float* far,
                                     int64_t R, int32_t last, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0, "sphere_trace_step: bad R");
  if (R == 0) return MLI_OK;
  sphere_trace_step_kernel<<<mli_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(dist, mask, sdf, near, far, R, last);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_light_rays(const float* center, const float* ray_unit, const float* inter_dist, const float* pts_light,
                              int64_t R, float radius, const float* host_aabb6, float* light_unit, float* near_light,
                              float* far_tracing, uint8_t* inside_bounding, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0, "light_rays: bad R");
  if (R == 0) return MLI_OK;
  Bound b;
  b.radius = radius;
  b.use_box = host_aabb6 != nullptr;
  for (int k = 0; k < 6; ++k) b.aabb[k] = host_aabb6 ? host_aabb6[k] : 0.0f;
  light_rays_kernel<<<mli_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(center, ray_unit, inter_dist, pts_light, R, b, light_unit,
                                                                      near_light, far_tracing, inside_bounding);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_light_finish(const uint8_t* mask_light, const uint8_t* inside_bounding, const float* gradient,
                                const float* light_unit, int64_t R, uint8_t* visibility, float* normal_x_light,
                                void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0, "light_finish: bad R");
  if (R == 0) return MLI_OK;
  light_finish_kernel<<<mli_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(mask_light, inside_bounding, gradient, light_unit, R,
                                                                        visibility, normal_x_light);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

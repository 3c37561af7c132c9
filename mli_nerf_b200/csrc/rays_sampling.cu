// rays_sampling.cu -- ray generation, near/far bounds, coarse + hierarchical importance sampling.
//
// Reference (relative to /root/reference/):
//   projects/NeuralLumen/model.py:120-131 (render_pixels_lumen), projects/nerf/utils/camera.py:283-311,
//   projects/neuralangelo/model.py:420-430 (get_dist_bounds), :449-465 (sample_dists_all),
//   :467-484 (sample_dists_hierarchical), projects/nerf/utils/nerf_util.py:20-38,41-68,199-205,
//   projects/NeuralLumen/utils/utils.py:61-79,86-123.
//
// Layout: per-ray arrays [R, ld] with ld = final sample count (128); a ray's samples are contiguous (512 B)
// so one warp reads a ray with a single coalesced 128-bit-per-lane request.  One warp per ray; the parts that
// torch evaluates sequentially with a float64 accumulator (cumprod/cumsum) keep their dependent chain on lane 0
// (one DMUL / DADD per element, operands prepared by all lanes) so the integer bin indices are bit-identical to
// the oracle; everything else is lane-parallel.
// Bound: HBM bandwidth (~1.4 KB/ray/round), in practice launch latency.
#include "common.cuh"

namespace {

constexpr int kMaxN = 256;  // max samples per ray handled by the warp-per-ray kernels
constexpr int kWarpsPerBlock = 4;

// ---------------------------------------------------------------------------------------------------------
__global__ void rays_from_pose_kernel(const float* __restrict__ pose, const float* __restrict__ intr,
                                      const float* __restrict__ pose_light, const int64_t* __restrict__ ray_idx,
                                      int64_t B, int64_t R, int W, float* __restrict__ center,
                                      float* __restrict__ ray_unit, float* __restrict__ ray_norm,
                                      float* __restrict__ pts_light) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * R) return;
  const int64_t b = t / R;
  const float* P = pose + b * 12;
  const float* K = intr + b * 9;
  const float* L = pose_light + b * 12;
  const int64_t pix = ray_idx ? ray_idx[t] : (t - b * R);
  const float px = (float)(pix % W) + 0.5f, py = (float)(pix / W) + 0.5f;
  // inverse intrinsics by adjugate
  const float a = K[0], bb = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
  const float A = e * i - f * h, Bc = -(d * i - f * g), C = d * h - e * g;
  const float det = a * A + bb * Bc + c * C, inv = 1.0f / det;
  const float i00 = A * inv, i01 = -(bb * i - c * h) * inv, i02 = (bb * f - c * e) * inv;
  const float i10 = Bc * inv, i11 = (a * i - c * g) * inv, i12 = -(a * f - c * d) * inv;
  const float i20 = C * inv, i21 = -(a * h - bb * g) * inv, i22 = (a * e - bb * d) * inv;
  const float cx = i00 * px + i01 * py + i02, cy = i10 * px + i11 * py + i12, cz = i20 * px + i21 * py + i22;
  // pose is world->camera [R|t]; inverse: R^T, -R^T t
  float ctr[3], dir[3], lc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ctr[k] = -(P[0 * 4 + k] * P[3] + P[1 * 4 + k] * P[7] + P[2 * 4 + k] * P[11]);
    lc[k] = -(L[0 * 4 + k] * L[3] + L[1 * 4 + k] * L[7] + L[2 * 4 + k] * L[11]);
    const float world = P[0 * 4 + k] * cx + P[1 * 4 + k] * cy + P[2 * 4 + k] * cz + ctr[k];
    dir[k] = world - ctr[k];
  }
  const float nrm = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  const float den = fmaxf(nrm, 1e-12f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    center[t * 3 + k] = ctr[k];
    ray_unit[t * 3 + k] = dir[k] / den;
    pts_light[t * 3 + k] = lc[k];
  }
  if (ray_norm) ray_norm[t] = nrm;
}

struct Aabb { float v[6]; int use; };

__global__ void dist_bounds_kernel(const float* __restrict__ center, const float* __restrict__ ray_unit, int64_t R,
                                   Aabb box, float* __restrict__ near, float* __restrict__ far,
                                   uint8_t* __restrict__ outside) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float c[3] = {center[r * 3], center[r * 3 + 1], center[r * 3 + 2]};
  const float v[3] = {ray_unit[r * 3], ray_unit[r * 3 + 1], ray_unit[r * 3 + 2]};
  float n, f;
  uint8_t o;
  if (box.use) mli_bounds_aabb(c, v, box.v, &n, &f, &o);
  else mli_bounds_sphere(c, v, &n, &f, &o);
  near[r] = n; far[r] = f; outside[r] = o;
}

__global__ void sample_coarse_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                     const float* __restrict__ rands, int64_t R, int n, float* __restrict__ dists,
                                     int64_t ld) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= R * n) return;
  const int64_t r = t / n;
  const int i = (int)(t - r * n);
  // rands += arange; dists = rands / intvs * (far - near) + near   (nerf_util.py:36-37)
  const float u = mli_add(rands ? rands[t] : 0.5f, (float)i);
  dists[r * ld + i] = mli_add(mli_mul(mli_div(u, (float)n), mli_sub(far[r], near[r])), near[r]);
}

// ---------------------------------------------------------------------------------------------------------
// one warp per ray: weights -> cdf -> bins -> fine samples
// ---------------------------------------------------------------------------------------------------------
struct FineSmem {
  double f64[kMaxN];
  float d[kMaxN], s[kMaxN], w[kMaxN], cdf[kMaxN + 1];
};

// The two scans torch evaluates sequentially with a float64 accumulator (cumprod of 1 - alpha, cumsum of the pdf) stay
// sequential on lane 0 -- their roundings depend on the order -- but everything that is not on the dependent chain
// (the float -> double factors, the IEEE divisions, the final float products) is done by all lanes first, so that lane 0
// runs one DMUL / DADD per element fed by independent shared-memory loads.  Bit for bit mli_alphas_to_weights() and
// mli_weights_to_cdf(); a ray's latency through these kernels was 127 x ~350 cycles of lane-0 work before.
__device__ __forceinline__ void warp_alphas_to_weights(FineSmem& sm, int n_w, int lane) {
  for (int i = lane; i < n_w; i += 32) sm.f64[i] = (double)mli_sub(1.0f, sm.w[i]);
  __syncwarp();
  if (lane == 0) {
    double T = 1.0;
#pragma unroll 8
    for (int i = 0; i < n_w; ++i) {
      sm.cdf[i] = (float)T;  // transmittance in front of interval i (cdf[] is free until the next scan)
      T *= sm.f64[i];
    }
  }
  __syncwarp();
  for (int i = lane; i < n_w; i += 32) sm.w[i] = mli_mul(sm.w[i], sm.cdf[i]);
  __syncwarp();
}

__device__ __forceinline__ void warp_weights_to_cdf(FineSmem& sm, int n_w, int lane) {
  float denom = 0.0f;
  if (lane == 0) {
#pragma unroll 8
    for (int i = 0; i < n_w; ++i) denom = mli_add(denom, fabsf(sm.w[i]));
    denom = fmaxf(denom, 1e-12f);
  }
  denom = __shfl_sync(0xffffffffu, denom, 0);
  for (int i = lane; i < n_w; i += 32) sm.f64[i] = (double)mli_div(sm.w[i], denom);
  __syncwarp();
  if (lane == 0) {
    double acc = 0.0;
    sm.cdf[0] = 0.0f;
#pragma unroll 8
    for (int i = 0; i < n_w; ++i) {
      acc += sm.f64[i];
      sm.cdf[i + 1] = (float)acc;
    }
  }
  __syncwarp();
}

// rank of element i of cat(A[0..n), B[0..n_b)) under a stable ascending sort.  Both runs are sorted in every call the
// renderer makes (coarse samples ascend, inverse-CDF samples of ascending u ascend), which turns the rank into a binary
// search in the OTHER run; the all-pairs count is kept for anything else (unsorted input, NaNs): same result either way.
__device__ __forceinline__ bool warp_runs_sorted(const float* v, int n, int tot, int lane) {
  bool ok = true;
  for (int i = lane; i + 1 < tot; i += 32)
    if (i != n - 1) ok = ok && (v[i] <= v[i + 1]);
  return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ int stable_rank(const float* v, int n, int tot, int i, bool sorted) {
  const float x = v[i];
  if (sorted) {
    int lo, hi;
    if (i < n) {  // elements of B strictly below x
      lo = n; hi = tot;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] < x) lo = mid + 1; else hi = mid; }
      return i + (lo - n);
    }
    lo = 0; hi = n;  // elements of A not above x
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (v[mid] <= x) lo = mid + 1; else hi = mid; }
    return (i - n) + lo;
  }
  int rank = 0;
  for (int j = 0; j < tot; ++j) {
    const float o = v[j];
    rank += (o < x) || (o == x && j < i) || (x != x && o == o) || (x != x && o != o && j < i);  // NaN last
  }
  return rank;
}

__device__ __forceinline__ void bins_from_weights(FineSmem& sm, int n_w, int n_bins_src, int n_fine, int lane,
                                                  float* fine_out, int32_t* idx_o, int32_t* low_o, int32_t* high_o,
                                                  float* cdf_o) {
  // n_w weights -> cdf[0..n_w]; searchsorted over the n_w+1 cdf entries
  warp_weights_to_cdf(sm, n_w, lane);
  const int n = n_w + 1;
  for (int j = lane; j < n_fine; j += 32) {
    int idx, low, high;
    const float u = mli_unif(j, n_fine);
    const float v = mli_sample_bin(sm.d, sm.cdf, n, u, &idx, &low, &high);
    if (fine_out) fine_out[j] = v;
    if (idx_o) { idx_o[j] = idx; low_o[j] = low; high_o[j] = high; }
  }
  if (cdf_o) for (int j = lane; j < n; j += 32) cdf_o[j] = sm.cdf[j];
  (void)n_bins_src;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) sample_fine_kernel(
    const float* __restrict__ dists, const float* __restrict__ sdfs, int64_t ld, int64_t R, int n, int n_fine,
    float inv_s, float* __restrict__ fine, int32_t* __restrict__ idx, int32_t* __restrict__ low,
    int32_t* __restrict__ high, float* __restrict__ cdf) {
  __shared__ FineSmem smem[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  FineSmem& sm = smem[warp];
  for (int i = lane; i < n; i += 32) { sm.d[i] = dists[r * ld + i]; sm.s[i] = sdfs[r * ld + i]; }
  __syncwarp();
  for (int i = lane; i < n - 1; i += 32) sm.w[i] = mli_hier_alpha(sm.d, sm.s, i, inv_s);  // 2 sigmoids each: parallel
  __syncwarp();
  warp_alphas_to_weights(sm, n - 1, lane);  // float64 running product: sequential, like torch
  bins_from_weights(sm, n - 1, n, n_fine, lane, fine + r * n_fine, idx ? idx + r * n_fine : nullptr,
                    low ? low + r * n_fine : nullptr, high ? high + r * n_fine : nullptr, cdf ? cdf + r * n : nullptr);
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) pdf_bins_kernel(
    const float* __restrict__ weights, int64_t ld_w, int64_t R, int n_w, int n_fine, int32_t* __restrict__ idx,
    int32_t* __restrict__ low, int32_t* __restrict__ high, float* __restrict__ cdf) {
  __shared__ FineSmem smem[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  FineSmem& sm = smem[warp];
  for (int i = lane; i < n_w; i += 32) sm.w[i] = weights[r * ld_w + i];
  for (int i = lane; i <= n_w; i += 32) sm.d[i] = (float)i;  // dummy bin positions
  __syncwarp();
  bins_from_weights(sm, n_w, n_w + 1, n_fine, lane, nullptr, idx + r * n_fine, low + r * n_fine, high + r * n_fine,
                    cdf ? cdf + r * (n_w + 1) : nullptr);
}

// cat + stable sort (+ gather of sdfs): rank every element among the n + n_fine candidates
__global__ void __launch_bounds__(32 * kWarpsPerBlock) sample_merge_kernel(
    float* __restrict__ dists, float* __restrict__ sdfs, int64_t ld, int64_t R, int n, const float* __restrict__ fine,
    const float* __restrict__ sdf_fine, int n_fine) {
  __shared__ float sd[kWarpsPerBlock][kMaxN], ss[kWarpsPerBlock][kMaxN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  const int tot = n + n_fine;
  for (int i = lane; i < tot; i += 32) {
    sd[warp][i] = i < n ? dists[r * ld + i] : fine[r * n_fine + (i - n)];
    if (sdfs) ss[warp][i] = i < n ? sdfs[r * ld + i] : sdf_fine[r * n_fine + (i - n)];
  }
  __syncwarp();
  const bool sorted = warp_runs_sorted(sd[warp], n, tot, lane);
  for (int i = lane; i < tot; i += 32) {
    const int rank = stable_rank(sd[warp], n, tot, i, sorted);
    dists[r * ld + rank] = sd[warp][i];
    if (sdfs) sdfs[r * ld + rank] = ss[warp][i];
  }
}

// merge (cat + stable sort + gather of sdfs) of round h fused with the importance sampling of round h + 1: the merged
// ray stays in shared memory between the two (saves a launch and the re-read on the latency-bound sampling path).  Same
// arithmetic, in the same order, as sample_merge_kernel followed by sample_fine_kernel.
__global__ void __launch_bounds__(32 * kWarpsPerBlock) sample_merge_fine_kernel(
    float* __restrict__ dists, float* __restrict__ sdfs, int64_t ld, int64_t R, int n, const float* __restrict__ fine_in,
    const float* __restrict__ sdf_fine, int n_fine, float inv_s_next, float* __restrict__ fine_out) {
  __shared__ FineSmem smem[kWarpsPerBlock];
  __shared__ float sd[kWarpsPerBlock][kMaxN], ss[kWarpsPerBlock][kMaxN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (r >= R) return;
  FineSmem& sm = smem[warp];
  const int tot = n + n_fine;
  for (int i = lane; i < tot; i += 32) {
    sd[warp][i] = i < n ? dists[r * ld + i] : fine_in[r * n_fine + (i - n)];
    ss[warp][i] = i < n ? sdfs[r * ld + i] : sdf_fine[r * n_fine + (i - n)];
  }
  __syncwarp();
  const bool sorted = warp_runs_sorted(sd[warp], n, tot, lane);
  for (int i = lane; i < tot; i += 32) {
    const int rank = stable_rank(sd[warp], n, tot, i, sorted);
    sm.d[rank] = sd[warp][i];
    sm.s[rank] = ss[warp][i];
  }
  __syncwarp();
  for (int i = lane; i < tot; i += 32) { dists[r * ld + i] = sm.d[i]; sdfs[r * ld + i] = sm.s[i]; }
  for (int i = lane; i < tot - 1; i += 32) sm.w[i] = mli_hier_alpha(sm.d, sm.s, i, inv_s_next);
  __syncwarp();
  warp_alphas_to_weights(sm, tot - 1, lane);
  bins_from_weights(sm, tot - 1, tot, n_fine, lane, fine_out + r * n_fine, nullptr, nullptr, nullptr, nullptr);
}

}  // namespace

extern "C" int mli_sample_merge_fine(float* dists, float* sdfs, int64_t ld, int64_t R, int32_t n, const float* fine_in,
                                     const float* sdf_fine, int32_t n_fine, float inv_s_next, float* fine_out,
                                     void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0 && n >= 1 && n_fine >= 1 && n + n_fine <= kMaxN && ld >= n + n_fine, "sample_merge_fine: bad shape");
  MLI_REQUIRE(dists && sdfs && fine_in && sdf_fine && fine_out && fine_out != fine_in, "sample_merge_fine: NULL / aliased argument");
  if (R == 0) return MLI_OK;
  sample_merge_fine_kernel<<<mli_cdiv(R, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
      dists, sdfs, ld, R, n, fine_in, sdf_fine, n_fine, inv_s_next, fine_out);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_rays_from_pose(const float* pose, const float* intr, const float* pose_light,
                                  const int64_t* ray_idx, int64_t B, int64_t R, int32_t W, float* center,
                                  float* ray_unit, float* ray_norm, float* pts_light, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(B >= 1 && R >= 0 && W >= 1, "rays_from_pose: bad B/R/W");
  if (R == 0) return MLI_OK;
  rays_from_pose_kernel<<<mli_cdiv(B * R, 256), 256, 0, (cudaStream_t)stream>>>(pose, intr, pose_light, ray_idx, B, R, W,
                                                                             center, ray_unit, ray_norm, pts_light);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_dist_bounds(const float* center, const float* ray_unit, int64_t R, const float* host_aabb6,
                               float* near, float* far, uint8_t* outside, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0, "dist_bounds: bad R");
  if (R == 0) return MLI_OK;
  Aabb box;
  box.use = host_aabb6 != nullptr;
  for (int k = 0; k < 6; ++k) box.v[k] = host_aabb6 ? host_aabb6[k] : 0.0f;
  dist_bounds_kernel<<<mli_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(center, ray_unit, R, box, near, far, outside);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_sample_coarse(const float* near, const float* far, const float* rands, int64_t R, int32_t n,
                                 float* dists, int64_t ld_d, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0 && n >= 1 && ld_d >= n, "sample_coarse: bad R/n/ld");
  if (R == 0) return MLI_OK;
  sample_coarse_kernel<<<mli_cdiv(R * n, 256), 256, 0, (cudaStream_t)stream>>>(near, far, rands, R, n, dists, ld_d);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_sample_fine(const float* dists, const float* sdfs, int64_t ld, int64_t R, int32_t n,
                               int32_t n_fine, float inv_s, float* fine, int32_t* idx, int32_t* low, int32_t* high,
                               float* cdf, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0 && n >= 2 && n <= kMaxN && ld >= n && n_fine >= 1, "sample_fine: bad R/n/ld (n <= %d)", kMaxN);
  MLI_REQUIRE((idx == nullptr) == (low == nullptr) && (idx == nullptr) == (high == nullptr),
              "sample_fine: idx/low/high must be given together");
  if (R == 0) return MLI_OK;
  sample_fine_kernel<<<mli_cdiv(R, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(
      dists, sdfs, ld, R, n, n_fine, inv_s, fine, idx, low, high, cdf);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_pdf_bins(const float* weights, int64_t ld_w, int64_t R, int32_t n_w, int32_t n_fine, int32_t* idx,
                            int32_t* low, int32_t* high, float* cdf, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0 && n_w >= 1 && n_w < kMaxN && ld_w >= n_w && n_fine >= 1, "pdf_bins: bad shape");
  MLI_REQUIRE(idx && low && high, "pdf_bins: idx/low/high required");
  if (R == 0) return MLI_OK;
  pdf_bins_kernel<<<mli_cdiv(R, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(weights, ld_w, R, n_w,
                                                                                               n_fine, idx, low, high, cdf);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_sample_merge(float* dists, float* sdfs, int64_t ld, int64_t R, int32_t n, const float* fine,
                                const float* sdf_fine, int32_t n_fine, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(R >= 0 && n >= 1 && n_fine >= 1 && n + n_fine <= kMaxN && ld >= n + n_fine, "sample_merge: bad shape");
  MLI_REQUIRE((sdfs == nullptr) == (sdf_fine == nullptr), "sample_merge: sdfs and sdf_fine go together");
  if (R == 0) return MLI_OK;
  sample_merge_kernel<<<mli_cdiv(R, kWarpsPerBlock), 32 * kWarpsPerBlock, 0, (cudaStream_t)stream>>>(dists, sdfs, ld, R, n,
                                                                                                   fine, sdf_fine, n_fine);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

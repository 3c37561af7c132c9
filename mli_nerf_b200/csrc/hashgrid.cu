// hashgrid.cu -- multi-resolution hash-grid encoding for sm_100a (replaces tcnn.Encoding HashGrid,
// /root/reference/projects/neuralangelo/utils/modules.py:42-50,84-86).
//
// Data layout in HBM: one flat fp32 table [n_entries][F] (the reference's state_dict tensor
// neural_sdf.tcnn_encoding.params, level-major).  With F = 8 one entry is exactly one 32-byte DRAM/L2 sector,
// so every corner fetch is one fully used sector (2 x LDG.128).  Thread mapping: one thread per
// (sample, level), gridDim.y = level, so a CTA touches a single level's slab (L2 locality for the dense
// levels 0-5, 57 MB) and the per-level constants are warp-uniform.  The ray variant loops over the 1+taps
// stencil planes inside the thread: the 5 points of Neuralangelo's 4-tap stencil are < 0.3 fine cells
// apart, so their corner sectors mostly coincide and are served by L1 instead of a second trip to L2/HBM.
// Bound: HBM/L2 gather bandwidth (16 levels * 8 corners * 32 B = 4 KB per query).
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <int F>
__device__ __forceinline__ void load_entry(const float* __restrict__ table, uint32_t row, float* v) {
  if constexpr (F == 8) {
    // one 256-bit load per corner (LDG.E.256, sm_100+): the whole 32-byte sector in a single L1 request instead of two
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
        : "l"(table + (size_t)row * 8));
  } else if constexpr (F == 4) {
    float4 a = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    float2 a = __ldg(reinterpret_cast<const float2*>(table + (size_t)row * 2));
    v[0] = a.x; v[1] = a.y;
  }
}

template <int F>
__device__ __forceinline__ void interp(const mli_level_t& lv, const float* __restrict__ table, float x, float y,
                                       float z, float* acc) {
  mli_cell_t cell = mli_grid_cell(lv, x, y, z);
  uint32_t rows[8];
  float wts[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) mli_corner(lv, cell, c, &rows[c], &wts[c]);
  float vals[8][F];
#pragma unroll
  for (int c = 0; c < 8; ++c) load_entry<F>(table, rows[c], vals[c]);  // 8 independent sector fetches in flight
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c)  // same corner order as tcnn's fma chain
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = fmaf(wts[c], vals[c][f], acc[f]);
}

template <int F>
__device__ __forceinline__ void store_feat(float* __restrict__ dst, const float* acc) {
  if constexpr (F == 8) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else if constexpr (F == 4) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else {
    reinterpret_cast<float2*>(dst)[0] = make_float2(acc[0], acc[1]);
  }
}

template <int F>
__device__ __forceinline__ void scatter_entry(float* __restrict__ grad, uint32_t row, float w, const float* d) {
  float* p = grad + (size_t)row * F;
  if constexpr (F == 8) {
    // vector reductions (RED.E.ADD.F32x4 on sm_90+): 2 per corner instead of 8 scalar atomics
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(w * d[0], w * d[1], w * d[2], w * d[3]));
    atomicAdd(reinterpret_cast<float4*>(p) + 1, make_float4(w * d[4], w * d[5], w * d[6], w * d[7]));
  } else if constexpr (F == 4) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(w * d[0], w * d[1], w * d[2], w * d[3]));
  } else {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(w * d[0], w * d[1]));
  }
}

// ---------------------------------------------------------------------------------------------------------
// tcnn-compatible entry: x01 [M,3] -> out [M, ld]
// ---------------------------------------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(kThreads) hashgrid_fwd_kernel(mli_grid_t grid, const float* __restrict__ table,
                                                                const float* __restrict__ x01, int64_t M,
                                                                float* __restrict__ out, int64_t ld) {
  const int level = blockIdx.y;
  const int64_t m = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (m >= M) return;
  float acc[F];
  if (level >= (int)grid.active_levels) {
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.0f;
  } else {
    interp<F>(grid.level[level], table, x01[m * 3 + 0], x01[m * 3 + 1], x01[m * 3 + 2], acc);
  }
  store_feat<F>(out + m * ld + level * F, acc);
}

template <int F>
__global__ void __launch_bounds__(kThreads) hashgrid_bwd_kernel(mli_grid_t grid, const float* __restrict__ x01,
                                                                int64_t M, const float* __restrict__ d_out,
                                                                int64_t ld, float* __restrict__ table_grad) {
  const int level = blockIdx.y;
  const int64_t m = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (m >= M || level >= (int)grid.active_levels) return;
  const mli_level_t& lv = grid.level[level];
  float d[F];
#pragma unroll
  for (int f = 0; f < F; ++f) d[f] = d_out[m * ld + level * F + f];
  mli_cell_t cell = mli_grid_cell(lv, x01[m * 3 + 0], x01[m * 3 + 1], x01[m * 3 + 2]);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t row;
    float w;
    mli_corner(lv, cell, c, &row, &w);
    scatter_entry<F>(table_grad, row, w, d);
  }
}

__global__ void hashgrid_corners_kernel(mli_grid_t grid, uint32_t level, const float* __restrict__ x01, int64_t M,
                                        uint32_t* __restrict__ idx) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const mli_level_t& lv = grid.level[level];
  mli_cell_t cell = mli_grid_cell(lv, x01[m * 3 + 0], x01[m * 3 + 1], x01[m * 3 + 2]);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t row;
    float w;
    mli_corner(lv, cell, c, &row, &w);
    idx[m * 8 + c] = row;
  }
}

// ---------------------------------------------------------------------------------------------------------
// ray variant: points from (center, ray_unit, dists), 1+taps stencil planes, rows [enc | xyz | 0-pad]
// ---------------------------------------------------------------------------------------------------------
struct RayArgs {
  const float* center;
  const float* ray_unit;
  const float* dists;
  int64_t ld_d, R;
  int32_t n, taps;
  float tap_eps, vol_min, vol_range;
  float inv_range;  // 1/vol_range when that is exact (power-of-two range, e.g. [-2, 2]): x/range == x*inv_range bit for bit
};

__device__ __forceinline__ void ray_point01(const RayArgs& a, int64_t ray, int i, int plane, float* p, float* x01) {
  float c[3] = {a.center[ray * 3], a.center[ray * 3 + 1], a.center[ray * 3 + 2]};
  float r[3] = {a.ray_unit[ray * 3], a.ray_unit[ray * 3 + 1], a.ray_unit[ray * 3 + 2]};
  mli_sample_point(c, r, a.dists[ray * a.ld_d + i], a.taps, plane, a.tap_eps, p);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
}

template <int F>
__global__ void __launch_bounds__(kThreads) encode_rays_kernel(mli_grid_t grid, const float* __restrict__ table,
                                                               RayArgs a, float* __restrict__ X, int64_t ldx) {
  const int level = blockIdx.y;
  const int64_t M = a.R * a.n;
  const int64_t m = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (m >= M) return;
  const int64_t ray = m / a.n;
  const int i = (int)(m - ray * a.n);
  const int planes = 1 + a.taps;
  const int enc = grid.n_levels * F;
  for (int pl = 0; pl < planes; ++pl) {
    float p[3], x01[3], acc[F];
    ray_point01(a, ray, i, pl, p, x01);
    float* row = X + ((int64_t)pl * M + m) * ldx;
    if (level >= (int)grid.active_levels) {
#pragma unroll
      for (int f = 0; f < F; ++f) acc[f] = 0.0f;
    } else {
      interp<F>(grid.level[level], table, x01[0], x01[1], x01[2], acc);
    }
    store_feat<F>(row + level * F, acc);
    if (level == 0) {  // xyz + zero padding, once per row
      row[enc + 0] = p[0]; row[enc + 1] = p[1]; row[enc + 2] = p[2];
      for (int k = enc + 3; k < ldx; ++k) row[k] = 0.0f;
    }
  }
}

template <int F>
__global__ void __launch_bounds__(kThreads) encode_rays_bwd_kernel(mli_grid_t grid, RayArgs a,
                                                                   const float* __restrict__ dX, int64_t ldx,
                                                                   float* __restrict__ table_grad, int delta_basis) {
  const int level = blockIdx.y;
  const int64_t M = a.R * a.n;
  const int64_t m = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (m >= M || level >= (int)grid.active_levels) return;
  const mli_level_t& lv = grid.level[level];
  const int64_t ray = m / a.n;
  const int i = (int)(m - ray * a.n);
  const int planes = 1 + a.taps;
  // Aggregate the stencil planes that fall into the centre's cell before touching memory: their 8 corner
  // rows are identical, so one vector reduction per corner carries all of them (up to 5x fewer atomics).
  //
  // delta_basis: the rows of dX are gradients w.r.t. the DELTA-basis inputs of the tensor-core SDF trunk
  // (plane 0: x_centre, plane i: x_tap_i - x_centre), i.e. plane 0 carries the SUM over all planes of the
  // per-plane gradients and plane i the tap's own.  Then  sum_p d_p w_p(c) = d_sum w_0(c) + sum_i d_i (w_i(c) - w_0(c)):
  // the large, mutually cancelling tap gradients only ever multiply weight DIFFERENCES, so their bf16 rounding
  // error is not amplified by the stencil's 1/eps.
  float p[3], x01[3];
  ray_point01(a, ray, i, 0, p, x01);
  const mli_cell_t cell0 = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
  float w0[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) w *= ((c >> k) & 1) ? cell0.w[k] : 1.0f - cell0.w[k];
    w0[c] = w;
  }
  float agg[8][F];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int f = 0; f < F; ++f) agg[c][f] = 0.0f;
  for (int pl = 0; pl < planes; ++pl) {
    mli_cell_t cell = cell0;
    if (pl) {
      ray_point01(a, ray, i, pl, p, x01);
      cell = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
    }
    float d[F];
    const float* src = dX + ((int64_t)pl * M + m) * ldx + level * F;
#pragma unroll
    for (int f = 0; f < F; ++f) d[f] = src[f];
    const bool same = cell.g[0] == cell0.g[0] && cell.g[1] == cell0.g[1] && cell.g[2] == cell0.g[2];
    const bool sub0 = delta_basis && pl > 0;  // tap plane in the delta basis: its gradient also leaves the centre
    if (same) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float w = 1.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) w *= ((c >> k) & 1) ? cell.w[k] : 1.0f - cell.w[k];
        if (sub0) w -= w0[c];
#pragma unroll
        for (int f = 0; f < F; ++f) agg[c][f] = fmaf(w, d[f], agg[c][f]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t row;
        float w;
        mli_corner(lv, cell, c, &row, &w);
        scatter_entry<F>(table_grad, row, w, d);
        if (sub0) {
#pragma unroll
          for (int f = 0; f < F; ++f) agg[c][f] = fmaf(-w0[c], d[f], agg[c][f]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t row;
    float w;
    mli_corner(lv, cell0, c, &row, &w);
    scatter_entry<F>(table_grad, row, 1.0f, agg[c]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// tensor-core variant of encode_rays: writes the SDF trunk's input directly as split-bf16 TCL in the delta basis
//   plane 0   : x0 = [enc(x) | xyz | 0]                      rows [0, M)
//   plane i>0 : x_i - x0 (formed in fp32, then split)          rows [i*M, (i+1)*M)
// TCL-128 with x_chunks chunks per tile row: chunk l (l < L) = level l's 8 features, chunk L = [xyz | 0], remaining
// chunks up to kc = 0; chunks [kc, 2 kc) hold the bf16 remainders x - bf16(x) (the "lo" half).
// One (sample, level) thread writes 16 contiguous bytes per plane and half; a warp writes 512 contiguous bytes.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split8(__nv_bfloat16* __restrict__ X, int x_chunks, int kc, int64_t grow, int chunk,
                                             const float* v) {
  const int64_t tile = grow >> 7;
  const int r = (int)(grow & 127);
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * f], v[2 * f + 1]);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * f] - hf.x, v[2 * f + 1] - hf.y);
    hi[f] = *reinterpret_cast<const uint32_t*>(&h);
    lo[f] = *reinterpret_cast<const uint32_t*>(&l);
  }
  *reinterpret_cast<uint4*>(X + ((tile * x_chunks + chunk) * 128 + r) * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(X + ((tile * x_chunks + kc + chunk) * 128 + r) * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Thread = (sample, level, plane): blockDim = (kS samples, planes).  A warp is 32 consecutive samples of ONE plane, so
// its stores are 512 contiguous bytes; the 1+taps warps that share a sample range run side by side, their corner
// fetches hit the same sectors (the tap points are < 0.15 finest cells from the centre) and merge in L1.  The centre
// row reaches the tap threads through shared memory (fp32) for the delta.  ~60 registers -> 40+ warps per SM: the
// previous thread-per-(sample, level) version looped over the planes at 16 warps per SM and was latency bound.
template <int KS>
__global__ void __launch_bounds__(KS * 7) encode_rays_tcl_kernel(mli_grid_t grid, const float* __restrict__ table,
                                                                 RayArgs a, __nv_bfloat16* __restrict__ X,
                                                                 int x_chunks, int kc) {
  constexpr int F = 8;
  __shared__ __align__(16) float s_acc[KS][F];
  __shared__ float s_p[KS][3];
  const int level = blockIdx.y;
  const int64_t M = a.R * a.n;
  const int sl = threadIdx.x, pl = threadIdx.y;
  const int64_t m = (int64_t)blockIdx.x * KS + sl;
  const bool valid = m < M;
  const int L = grid.n_levels;
  float p[3] = {0.f, 0.f, 0.f}, acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.0f;
  if (valid) {
    const int64_t ray = m / a.n;
    const int i = (int)(m - ray * a.n);
    float x01[3];
    ray_point01(a, ray, i, pl, p, x01);
    if (level < (int)grid.active_levels) interp<F>(grid.level[level], table, x01[0], x01[1], x01[2], acc);
  }
  if (pl == 0) {
#pragma unroll
    for (int f = 0; f < F; ++f) s_acc[sl][f] = acc[f];
    s_p[sl][0] = p[0]; s_p[sl][1] = p[1]; s_p[sl][2] = p[2];
  }
  __syncthreads();
  if (!valid) return;
  if (pl != 0) {
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] -= s_acc[sl][f];
    p[0] -= s_p[sl][0]; p[1] -= s_p[sl][1]; p[2] -= s_p[sl][2];
  }
  const int64_t grow = (int64_t)pl * M + m;
  store_split8(X, x_chunks, kc, grow, level, acc);
  if (level == 0) {  // xyz chunk + zero padding chunks, once per row
    float v[8] = {p[0], p[1], p[2], 0.f, 0.f, 0.f, 0.f, 0.f};
    store_split8(X, x_chunks, kc, grow, L, v);
    v[0] = v[1] = v[2] = 0.0f;
    for (int c = L + 1; c < kc; ++c) store_split8(X, x_chunks, kc, grow, c, v);
  }
}

// Corner-caching variant for the stencil launch (taps > 0): thread = (sample, level), all planes.  The tap points of
// Neuralangelo's stencil sit a fraction of a FINE cell from the centre, so on most levels they fall into the centre's
// cell and read the same eight table entries: those are fetched once (8 x LDG.256 in flight) and kept in registers;
// every same-cell tap is then pure arithmetic.  Taps that left the cell are deferred to a second, warp-compacted pass
// that re-uses the 64 value registers.  Per (sample, level) the kernel issues 8 + 8 x (taps outside the cell) sector requests instead of
// 8 x planes, which is what the thread-per-plane kernel above was bound by (L1TEX at 69 %: profiles/r02_summary.md section 7).  Results are bit
// for bit those of encode_rays_tcl_kernel: same weights, same fma chain, delta formed in fp32.
template <int PLANES>
__global__ void __launch_bounds__(128, 4) encode_rays_tcl_cached_kernel(mli_grid_t grid, const float* __restrict__ table,
                                                                        RayArgs a, __nv_bfloat16* __restrict__ X,
                                                                        int x_chunks, int kc) {
  constexpr int F = 8;
  __shared__ float s_acc0[128][F];
  __shared__ uint8_t s_items[4][32 * (PLANES - 1)];
  const int level = blockIdx.y;
  const int64_t M = a.R * a.n;
  const int64_t m = (int64_t)blockIdx.x * 128 + threadIdx.x;  // M is a multiple of 128 here (checked by the entry point)
  const int L = grid.n_levels;
  const bool active = level < (int)grid.active_levels;
  const mli_level_t& lv = grid.level[level];
  const int64_t ray = m / a.n;
  const int i = (int)(m - ray * a.n);
  const float rc[3] = {__ldg(a.center + ray * 3), __ldg(a.center + ray * 3 + 1), __ldg(a.center + ray * 3 + 2)};
  const float rr[3] = {__ldg(a.ray_unit + ray * 3), __ldg(a.ray_unit + ray * 3 + 1), __ldg(a.ray_unit + ray * 3 + 2)};
  const float rd = __ldg(a.dists + ray * a.ld_d + i);
  float p0[3], x01[3];
  mli_sample_point(rc, rr, rd, a.taps, 0, a.tap_eps, p0);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p0[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p0[k], a.vol_min), a.vol_range);
  const mli_cell_t cell0 = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
  float vals[8][F], acc0[F];
#pragma unroll
  for (int f = 0; f < F; ++f) acc0[f] = 0.0f;
  if (active) {
    uint32_t rows[8];
    float wts[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) mli_corner(lv, cell0, c, &rows[c], &wts[c]);
#pragma unroll
    for (int c = 0; c < 8; ++c) load_entry<F>(table, rows[c], vals[c]);
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int f = 0; f < F; ++f) acc0[f] = fmaf(wts[c], vals[c][f], acc0[f]);
  }
  store_split8(X, x_chunks, kc, m, level, acc0);
  if (level == 0) {
    float v[8] = {p0[0], p0[1], p0[2], 0.f, 0.f, 0.f, 0.f, 0.f};
    store_split8(X, x_chunks, kc, m, L, v);
    v[0] = v[1] = v[2] = 0.0f;
    for (int c = L + 1; c < kc; ++c) store_split8(X, x_chunks, kc, m, c, v);
  }
  uint32_t deferred = 0;
#pragma unroll 1
  for (int pl = 1; pl < PLANES; ++pl) {
    float p[3];
    mli_sample_point(rc, rr, rd, a.taps, pl, a.tap_eps, p);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
    const int64_t grow = (int64_t)pl * M + m;
    if (level == 0) {
      float v[8] = {p[0] - p0[0], p[1] - p0[1], p[2] - p0[2], 0.f, 0.f, 0.f, 0.f, 0.f};
      store_split8(X, x_chunks, kc, grow, L, v);
      v[0] = v[1] = v[2] = 0.0f;
      for (int c = L + 1; c < kc; ++c) store_split8(X, x_chunks, kc, grow, c, v);
    }
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.0f;
    if (active) {
      const mli_cell_t cell = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
      if (cell.g[0] != cell0.g[0] || cell.g[1] != cell0.g[1] || cell.g[2] != cell0.g[2]) {
        deferred |= 1u << pl;
        continue;
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float w = 1.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) w *= ((c >> d) & 1) ? cell.w[d] : 1.0f - cell.w[d];
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = fmaf(w, vals[c][f], acc[f]);
      }
#pragma unroll
      for (int f = 0; f < F; ++f) acc[f] -= acc0[f];
    }
    store_split8(X, x_chunks, kc, grow, level, acc);
  }
  // Second pass, compacted per warp: the (lane, plane) pairs that left the centre's cell are listed in shared memory and
  // handed out 32 at a time, so every round runs with full warps (the first version let each lane walk its own taps:
  // 30 % of all executed instructions ran with 6 of 32 lanes active: profiles/r02_summary.md section 7).  The owner's centre row comes
  // through shared memory; the item's ray is re-read (L1 hits).
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (__ballot_sync(0xffffffffu, deferred != 0) == 0) return;
#pragma unroll
  for (int f = 0; f < F; ++f) s_acc0[threadIdx.x][f] = acc0[f];
  int n_items = 0;
#pragma unroll 1
  for (int pl = 1; pl < PLANES; ++pl) {
    const bool mine = (deferred >> pl) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (mine) s_items[warp][n_items + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)(lane | ((unsigned)pl << 5));
    n_items += __popc(bal);
  }
  __syncwarp();
#pragma unroll 1
  for (int it = (int)lane; it < n_items; it += 32) {
    const unsigned item = s_items[warp][it];
    const unsigned src = item & 31u;
    const int pl = (int)(item >> 5);
    const int64_t m2 = m - (int64_t)lane + (int64_t)src;
    const int64_t ray2 = m2 / a.n;
    const int i2 = (int)(m2 - ray2 * a.n);
    const float c2[3] = {__ldg(a.center + ray2 * 3), __ldg(a.center + ray2 * 3 + 1), __ldg(a.center + ray2 * 3 + 2)};
    const float r2[3] = {__ldg(a.ray_unit + ray2 * 3), __ldg(a.ray_unit + ray2 * 3 + 1), __ldg(a.ray_unit + ray2 * 3 + 2)};
    const float d2 = __ldg(a.dists + ray2 * a.ld_d + i2);
    float p[3], acc[F];
    mli_sample_point(c2, r2, d2, a.taps, pl, a.tap_eps, p);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
    interp<F>(lv, table, x01[0], x01[1], x01[2], acc);
    const float* own = s_acc0[(warp << 5) + src];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] -= own[f];
    store_split8(X, x_chunks, kc, (int64_t)pl * M + m2, level, acc);
  }
}

// backward of encode_rays_tcl w.r.t. the table: dX is bf16 TCL-128 ([.., x_chunks, 128, 8], chunk l = level l) in the
// delta basis (see encode_rays_bwd_kernel).  All planes' 16-byte gradient slices are loaded up front (independent,
// fully coalesced: consecutive samples -> consecutive 16 B), then the same aggregation as the fp32 kernel.
constexpr int kBwdThreads = 128;

// Two threads per (sample, level), four features each: the per-thread state (corner accumulators 8x4, gradient slices
// PLANES x 8 B) halves, which lifts the kernel from 16 to 20 resident warps per SM without spills; every corner update is still one
// 16-byte RED per thread.  The cell / weight arithmetic is done by both threads of a pair (the kernel is latency bound,
// not issue bound).  Ray origin / direction / distance are loaded once, not once per plane.
template <int PLANES>
__global__ void __launch_bounds__(kBwdThreads, 5) encode_rays_bwd_tcl_kernel(mli_grid_t grid, RayArgs a,
                                                                             const __nv_bfloat16* __restrict__ dX, int x_chunks,
                                                                             float* __restrict__ table_grad, int level0) {
  constexpr int FH = 4;
  const int level = level0 + blockIdx.y;
  const int64_t M = a.R * a.n;
  const int64_t t = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x;
  const int64_t m = t >> 1;
  const int half = (int)(t & 1);
  if (m >= M || level >= (int)grid.active_levels) return;
  const mli_level_t& lv = grid.level[level];
  const int64_t ray = m / a.n;
  const int i = (int)(m - ray * a.n);
  uint2 draw[PLANES];
#pragma unroll
  for (int pl = 0; pl < PLANES; ++pl) {
    const int64_t grow = (int64_t)pl * M + m;
    draw[pl] = __ldg(reinterpret_cast<const uint2*>(dX + (((grow >> 7) * x_chunks + level) * 128 + (grow & 127)) * 8 + half * 4));
  }
  const float rc[3] = {__ldg(a.center + ray * 3), __ldg(a.center + ray * 3 + 1), __ldg(a.center + ray * 3 + 2)};
  const float rr[3] = {__ldg(a.ray_unit + ray * 3), __ldg(a.ray_unit + ray * 3 + 1), __ldg(a.ray_unit + ray * 3 + 2)};
  const float rd = __ldg(a.dists + ray * a.ld_d + i);
  float p[3], x01[3];
  mli_sample_point(rc, rr, rd, a.taps, 0, a.tap_eps, p);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
  const mli_cell_t cell0 = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
  float w0[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = 1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) w *= ((c >> k) & 1) ? cell0.w[k] : 1.0f - cell0.w[k];
    w0[c] = w;
  }
  float agg[8][FH];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int f = 0; f < FH; ++f) agg[c][f] = 0.0f;
  float* const tg = table_grad + half * FH;
  // Rolled plane loop (the unrolled one was 3.8k instructions and stalled on instruction fetch); the preloaded gradient
  // slices stay in registers and rotate down one slot per iteration so that draw[0] is always the current plane.
#pragma unroll 1
  for (int pl = 0; pl < PLANES; ++pl) {
    mli_cell_t cell = cell0;
    if (pl) {
      mli_sample_point(rc, rr, rd, a.taps, pl, a.tap_eps, p);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
      cell = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
    }
    float d[FH];
    {
      const uint2 cur = draw[0];
#pragma unroll
      for (int k = 0; k + 1 < PLANES; ++k) draw[k] = draw[k + 1];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur);
#pragma unroll
      for (int k = 0; k < 2; ++k) { const float2 f2 = __bfloat1622float2(h[k]); d[2 * k] = f2.x; d[2 * k + 1] = f2.y; }
    }
    const bool same = cell.g[0] == cell0.g[0] && cell.g[1] == cell0.g[1] && cell.g[2] == cell0.g[2];
    if (same) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float w = 1.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) w *= ((c >> k) & 1) ? cell.w[k] : 1.0f - cell.w[k];
        if (pl) w -= w0[c];
#pragma unroll
        for (int f = 0; f < FH; ++f) agg[c][f] = fmaf(w, d[f], agg[c][f]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t row;
        float w;
        mli_corner(lv, cell, c, &row, &w);
        atomicAdd(reinterpret_cast<float4*>(tg + (size_t)row * 8), make_float4(w * d[0], w * d[1], w * d[2], w * d[3]));
#pragma unroll
        for (int f = 0; f < FH; ++f) agg[c][f] = fmaf(-w0[c], d[f], agg[c][f]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t row;
    float w;
    mli_corner(lv, cell0, c, &row, &w);
    atomicAdd(reinterpret_cast<float4*>(tg + (size_t)row * 8), make_float4(agg[c][0], agg[c][1], agg[c][2], agg[c][3]));
  }
}

// Version 2 of the stencil scatter (PLANES > 1).  Two changes against encode_rays_bwd_tcl_kernel, both aimed at what ncu
// showed for it (profiles/r02_summary.md section 7: 24 % of the executed instructions ran in the "tap left the centre's cell" branch with 8 of 32
// lanes active, and that branch issued 8 reductions per tap):
//  * every tap -- inside the centre's cell or in a neighbouring one -- first adds what it owes to the lattice points it
//    SHARES with the centre's cell into the centre's eight register accumulators (a tap one cell over along one axis
//    shares four of its eight corners with the centre), branch-free: per axis the tap's weight on the centre's lower /
//    upper lattice plane is selected from {1-w, w, 0} by the cell offset;
//  * the remaining corners of the taps that left the cell are listed per warp in shared memory and scattered in dense
//    rounds of 32 (lane, plane) items.
template <int PLANES>
__global__ void __launch_bounds__(kBwdThreads, 5) encode_rays_bwd_tcl_v2_kernel(mli_grid_t grid, RayArgs a,
                                                                                const __nv_bfloat16* __restrict__ dX, int x_chunks,
                                                                                float* __restrict__ table_grad, int level0) {
  constexpr int FH = 4;
  static_assert(PLANES > 1, "the single-plane launch uses encode_rays_bwd_tcl_kernel<1>");
  __shared__ uint2 s_d[kBwdThreads][PLANES - 1];
  __shared__ uint32_t s_g0[kBwdThreads][3];
  __shared__ uint8_t s_items[kBwdThreads / 32][32 * (PLANES - 1)];
  const int level = level0 + blockIdx.y;
  const int64_t M = a.R * a.n;
  const int64_t t = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x;  // 2 M is a multiple of the block size (entry point)
  const int64_t m = t >> 1;
  const int half = (int)(t & 1);
  if (level >= (int)grid.active_levels) return;  // block-uniform
  const mli_level_t& lv = grid.level[level];
  const int64_t ray = m / a.n;
  const int i = (int)(m - ray * a.n);
  uint2 draw[PLANES];
#pragma unroll
  for (int pl = 0; pl < PLANES; ++pl) {
    const int64_t grow = (int64_t)pl * M + m;
    draw[pl] = __ldg(reinterpret_cast<const uint2*>(dX + (((grow >> 7) * x_chunks + level) * 128 + (grow & 127)) * 8 + half * 4));
  }
  const float rc[3] = {__ldg(a.center + ray * 3), __ldg(a.center + ray * 3 + 1), __ldg(a.center + ray * 3 + 2)};
  const float rr[3] = {__ldg(a.ray_unit + ray * 3), __ldg(a.ray_unit + ray * 3 + 1), __ldg(a.ray_unit + ray * 3 + 2)};
  const float rd = __ldg(a.dists + ray * a.ld_d + i);
#pragma unroll
  for (int pl = 1; pl < PLANES; ++pl) s_d[threadIdx.x][pl - 1] = draw[pl];
  float p[3], x01[3];
  mli_sample_point(rc, rr, rd, a.taps, 0, a.tap_eps, p);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
  const mli_cell_t cell0 = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) s_g0[threadIdx.x][k] = cell0.g[k];
  float w0[8], agg[8][FH];
  {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&draw[0]);
    const float2 d01 = __bfloat1622float2(h[0]), d23 = __bfloat1622float2(h[1]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float w = 1.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) w *= ((c >> k) & 1) ? cell0.w[k] : 1.0f - cell0.w[k];
      w0[c] = w;
      agg[c][0] = w * d01.x; agg[c][1] = w * d01.y; agg[c][2] = w * d23.x; agg[c][3] = w * d23.y;
    }
  }
  uint32_t deferred = 0;
#pragma unroll 1
  for (int pl = 1; pl < PLANES; ++pl) {
    mli_sample_point(rc, rr, rd, a.taps, pl, a.tap_eps, p);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
    const mli_cell_t cell = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
    float d[FH];
    {
      const uint2 cur = draw[1];  // draw[1] is always the current tap plane: the slices rotate down one slot per iteration
#pragma unroll
      for (int k = 1; k + 1 < PLANES; ++k) draw[k] = draw[k + 1];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur);
#pragma unroll
      for (int k = 0; k < 2; ++k) { const float2 f2 = __bfloat1622float2(h[k]); d[2 * k] = f2.x; d[2 * k + 1] = f2.y; }
    }
    // weight of the tap on the centre cell's lower / upper lattice plane, per axis
    float fl[3], fu[3];
    bool moved = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int off = (int)(cell.g[k] - cell0.g[k]);
      const float wk = cell.w[k], nk = 1.0f - wk;
      fl[k] = off == 0 ? nk : (off == -1 ? wk : 0.0f);
      fu[k] = off == 0 ? wk : (off == 1 ? nk : 0.0f);
      moved = moved || off != 0;
    }
    if (moved) deferred |= 1u << pl;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float w = ((c & 1) ? fu[0] : fl[0]) * ((c & 2) ? fu[1] : fl[1]) * ((c & 4) ? fu[2] : fl[2]) - w0[c];
#pragma unroll
      for (int f = 0; f < FH; ++f) agg[c][f] = fmaf(w, d[f], agg[c][f]);
    }
  }
  float* const tg = table_grad + half * FH;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t row;
    float w;
    mli_corner(lv, cell0, c, &row, &w);
    atomicAdd(reinterpret_cast<float4*>(tg + (size_t)row * 8), make_float4(agg[c][0], agg[c][1], agg[c][2], agg[c][3]));
  }
  // second pass: the corners of the moved taps that are NOT lattice points of the centre's cell, compacted per warp
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (__ballot_sync(0xffffffffu, deferred != 0) == 0) return;
  int n_items = 0;
#pragma unroll 1
  for (int pl = 1; pl < PLANES; ++pl) {
    const bool mine = (deferred >> pl) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (mine) s_items[warp][n_items + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)(lane | ((unsigned)pl << 5));
    n_items += __popc(bal);
  }
  __syncwarp();
#pragma unroll 1
  for (int it = (int)lane; it < n_items; it += 32) {
    const unsigned item = s_items[warp][it];
    const unsigned src = (warp << 5) + (item & 31u);
    const int pl = (int)(item >> 5);
    const int64_t t2 = t - (int64_t)lane + (int64_t)(item & 31u);
    const int64_t m2 = t2 >> 1;
    const int64_t ray2 = m2 / a.n;
    const int i2 = (int)(m2 - ray2 * a.n);
    const float c2[3] = {__ldg(a.center + ray2 * 3), __ldg(a.center + ray2 * 3 + 1), __ldg(a.center + ray2 * 3 + 2)};
    const float r2[3] = {__ldg(a.ray_unit + ray2 * 3), __ldg(a.ray_unit + ray2 * 3 + 1), __ldg(a.ray_unit + ray2 * 3 + 2)};
    const float d2 = __ldg(a.dists + ray2 * a.ld_d + i2);
    mli_sample_point(c2, r2, d2, a.taps, pl, a.tap_eps, p);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      x01[k] = a.inv_range != 0.0f ? mli_mul(mli_sub(p[k], a.vol_min), a.inv_range) : mli_div(mli_sub(p[k], a.vol_min), a.vol_range);
    const mli_cell_t cell = mli_grid_cell(lv, x01[0], x01[1], x01[2]);
    int off[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) off[k] = (int)(cell.g[k] - s_g0[src][k]);
    const uint2 cur = s_d[src][pl - 1];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&cur);
    const float2 d01 = __bfloat1622float2(h[0]), d23 = __bfloat1622float2(h[1]);
    float* const tg2 = table_grad + (t2 & 1) * FH;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      bool shared = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int q = off[k] + ((c >> k) & 1);
        shared = shared && (q == 0 || q == 1);
      }
      if (shared) continue;
      uint32_t row;
      float w;
      mli_corner(lv, cell, c, &row, &w);
      atomicAdd(reinterpret_cast<float4*>(tg2 + (size_t)row * 8), make_float4(w * d01.x, w * d01.y, w * d23.x, w * d23.y));
    }
  }
}

}  // namespace

#define DISPATCH_F(feat, CALL)                                      \
  switch (feat) {                                                   \
    case 8: { constexpr int F = 8; CALL; } break;                   \
    case 4: { constexpr int F = 4; CALL; } break;                   \
    case 2: { constexpr int F = 2; CALL; } break;                   \
    default: mli_set_error("unsupported n_features_per_level %u", feat); return MLI_EINVAL; \
  }

static int check_grid(const mli_grid_t* g) {
  MLI_REQUIRE(g != nullptr, "grid is NULL");
  MLI_REQUIRE(g->n_levels >= 1 && g->n_levels <= MLI_MAX_LEVELS, "bad n_levels %u", g->n_levels);
  MLI_REQUIRE(g->active_levels <= g->n_levels, "active_levels %u > n_levels %u", g->active_levels, g->n_levels);
  return MLI_OK;
}

extern "C" int mli_hashgrid_fwd(const mli_grid_t* grid, const float* table, const float* x01, int64_t M, float* out,
                                int64_t ld_out, void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  MLI_REQUIRE(M >= 0 && ld_out >= (int64_t)grid->n_levels * grid->feat, "bad M/ld_out");
  MLI_REQUIRE(ld_out % 4 == 0 || grid->feat == 2, "ld_out must keep rows 16-byte aligned");
  if (M == 0) return MLI_OK;
  dim3 g(mli_cdiv(M, kThreads), grid->n_levels);
  DISPATCH_F(grid->feat, (hashgrid_fwd_kernel<F><<<g, kThreads, 0, (cudaStream_t)stream>>>(*grid, table, x01, M, out,
                                                                                          ld_out)));
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_hashgrid_bwd(const mli_grid_t* grid, const float* x01, int64_t M, const float* d_out,
                                int64_t ld_dout, float* table_grad, void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  MLI_REQUIRE(M >= 0 && ld_dout >= (int64_t)grid->n_levels * grid->feat, "bad M/ld_dout");
  if (M == 0) return MLI_OK;
  dim3 g(mli_cdiv(M, kThreads), grid->n_levels);
  DISPATCH_F(grid->feat, (hashgrid_bwd_kernel<F><<<g, kThreads, 0, (cudaStream_t)stream>>>(*grid, x01, M, d_out,
                                                                                          ld_dout, table_grad)));
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_hashgrid_corners(const mli_grid_t* grid, uint32_t level, const float* x01, int64_t M,
                                    uint32_t* idx, void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  MLI_REQUIRE(level < grid->n_levels, "level %u out of range", level);
  if (M == 0) return MLI_OK;
  hashgrid_corners_kernel<<<mli_cdiv(M, 256), 256, 0, (cudaStream_t)stream>>>(*grid, level, x01, M, idx);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

static int make_ray_args(RayArgs* a, const float* center, const float* ray_unit, const float* dists, int64_t ld_d,
                         int64_t R, int32_t n, int32_t taps, float tap_eps, float vol_min, float vol_max) {
  MLI_REQUIRE(taps == 0 || taps == 4 || taps == 6, "Only support 4 or 6 taps.");
  MLI_REQUIRE(R >= 0 && n >= 1 && ld_d >= n, "bad R/n/ld_d");
  MLI_REQUIRE(vol_max > vol_min, "empty volume range");
  a->center = center; a->ray_unit = ray_unit; a->dists = dists; a->ld_d = ld_d; a->R = R; a->n = n;
  a->taps = taps; a->tap_eps = tap_eps; a->vol_min = vol_min; a->vol_range = vol_max - vol_min;
  int ex = 0;
  a->inv_range = frexpf(a->vol_range, &ex) == 0.5f ? 1.0f / a->vol_range : 0.0f;  // power of two -> exact reciprocal
  return MLI_OK;
}

extern "C" int mli_encode_rays(const mli_grid_t* grid, const float* table, const float* center,
                               const float* ray_unit, const float* dists, int64_t ld_d, int64_t R, int32_t n,
                               int32_t taps, float tap_eps, float vol_min, float vol_max, float* X, int64_t ldx,
                               void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  RayArgs a;
  if (int e = make_ray_args(&a, center, ray_unit, dists, ld_d, R, n, taps, tap_eps, vol_min, vol_max)) return e;
  MLI_REQUIRE(ldx >= (int64_t)grid->n_levels * grid->feat + 3 && ldx % 4 == 0, "ldx too small / unaligned");
  if (R == 0) return MLI_OK;
  dim3 g(mli_cdiv(R * n, kThreads), grid->n_levels);
  DISPATCH_F(grid->feat, (encode_rays_kernel<F><<<g, kThreads, 0, (cudaStream_t)stream>>>(*grid, table, a, X, ldx)));
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_encode_rays_bwd(const mli_grid_t* grid, const float* center, const float* ray_unit,
                                   const float* dists, int64_t ld_d, int64_t R, int32_t n, int32_t taps,
                                   float tap_eps, float vol_min, float vol_max, const float* dX, int64_t ldx,
                                   float* table_grad, int32_t delta_basis, void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  RayArgs a;
  if (int e = make_ray_args(&a, center, ray_unit, dists, ld_d, R, n, taps, tap_eps, vol_min, vol_max)) return e;
  if (R == 0) return MLI_OK;
  dim3 g(mli_cdiv(R * n, kThreads), grid->n_levels);
  DISPATCH_F(grid->feat,
             (encode_rays_bwd_kernel<F><<<g, kThreads, 0, (cudaStream_t)stream>>>(*grid, a, dX, ldx, table_grad,
                                                                                  delta_basis)));
  MLI_LAUNCH_OK();
  return MLI_OK;
}

static int encode_variant() {
  const char* e = getenv("MLI_ENCODE_VARIANT");  // read per call: tests and tools/bench_encode.py switch it at run time
  return e ? atoi(e) : 2;
}

extern "C" int mli_encode_rays_tcl(const mli_grid_t* grid, const float* table, const float* center,
                                   const float* ray_unit, const float* dists, int64_t ld_d, int64_t R, int32_t n,
                                   int32_t taps, float tap_eps, float vol_min, float vol_max, void* X, int32_t x_chunks,
                                   int32_t k_chunks, void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  RayArgs a;
  if (int e = make_ray_args(&a, center, ray_unit, dists, ld_d, R, n, taps, tap_eps, vol_min, vol_max)) return e;
  MLI_REQUIRE(grid->feat == 8, "encode_rays_tcl: n_features_per_level must be 8 (one level = one 8-column chunk)");
  MLI_REQUIRE(k_chunks >= (int32_t)grid->n_levels + 1 && k_chunks % 2 == 0 && x_chunks >= 2 * k_chunks,
              "encode_rays_tcl: need k_chunks >= levels+1 (even) and x_chunks >= 2*k_chunks");
  MLI_REQUIRE(taps == 0 || (R * n) % 128 == 0, "encode_rays_tcl: with taps, R*n must be a multiple of 128 (plane-major tiles)");
  if (R == 0) return MLI_OK;
  if (taps == 0) {
    constexpr int KS = 128;
    dim3 g(mli_cdiv(R * n, KS), grid->n_levels);
    encode_rays_tcl_kernel<KS><<<g, dim3(KS, 1), 0, (cudaStream_t)stream>>>(*grid, table, a, (__nv_bfloat16*)X, x_chunks, k_chunks);
  } else if (encode_variant() == 1) {  // thread-per-plane kernel (MLI_ENCODE_VARIANT=1): kept for A/B and as the bit-exactness pin
    constexpr int KS = 64;
    dim3 g(mli_cdiv(R * n, KS), grid->n_levels);
    encode_rays_tcl_kernel<KS><<<g, dim3(KS, 1 + taps), 0, (cudaStream_t)stream>>>(*grid, table, a, (__nv_bfloat16*)X, x_chunks, k_chunks);
  } else {
    dim3 g(mli_cdiv(R * n, 128), grid->n_levels);
    if (taps == 4) encode_rays_tcl_cached_kernel<5><<<g, 128, 0, (cudaStream_t)stream>>>(*grid, table, a, (__nv_bfloat16*)X, x_chunks, k_chunks);
    else encode_rays_tcl_cached_kernel<7><<<g, 128, 0, (cudaStream_t)stream>>>(*grid, table, a, (__nv_bfloat16*)X, x_chunks, k_chunks);
  }
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_encode_rays_bwd_tcl(const mli_grid_t* grid, const float* center, const float* ray_unit,
                                       const float* dists, int64_t ld_d, int64_t R, int32_t n, int32_t taps,
                                       float tap_eps, float vol_min, float vol_max, const void* dX, int32_t x_chunks,
                                       float* table_grad, int32_t level_begin, int32_t level_end, void* stream) {
  MLI_ENTRY();
  if (int e = check_grid(grid)) return e;
  RayArgs a;
  if (int e = make_ray_args(&a, center, ray_unit, dists, ld_d, R, n, taps, tap_eps, vol_min, vol_max)) return e;
  MLI_REQUIRE(grid->feat == 8 && x_chunks >= (int32_t)grid->n_levels, "encode_rays_bwd_tcl: needs F = 8 and one chunk per level");
  MLI_REQUIRE(taps == 0 || (R * n) % 128 == 0, "encode_rays_bwd_tcl: with taps, R*n must be a multiple of 128");
  MLI_REQUIRE(level_begin >= 0 && level_begin <= level_end && level_end <= (int32_t)grid->n_levels, "encode_rays_bwd_tcl: bad level range");
  if (R == 0 || level_begin == level_end) return MLI_OK;
  dim3 g(mli_cdiv(2 * R * n, kBwdThreads), level_end - level_begin);
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* d = (const __nv_bfloat16*)dX;
  const bool v1 = encode_variant() == 1;  // MLI_ENCODE_VARIANT=1: the first-generation kernels, kept for A/B and as the pin
  if (taps == 4 && v1) encode_rays_bwd_tcl_kernel<5><<<g, kBwdThreads, 0, st>>>(*grid, a, d, x_chunks, table_grad, level_begin);
  else if (taps == 6 && v1) encode_rays_bwd_tcl_kernel<7><<<g, kBwdThreads, 0, st>>>(*grid, a, d, x_chunks, table_grad, level_begin);
  else if (taps == 4) encode_rays_bwd_tcl_v2_kernel<5><<<g, kBwdThreads, 0, st>>>(*grid, a, d, x_chunks, table_grad, level_begin);
  else if (taps == 6) encode_rays_bwd_tcl_v2_kernel<7><<<g, kBwdThreads, 0, st>>>(*grid, a, d, x_chunks, table_grad, level_begin);
  else encode_rays_bwd_tcl_kernel<1><<<g, kBwdThreads, 0, st>>>(*grid, a, d, x_chunks, table_grad, level_begin);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

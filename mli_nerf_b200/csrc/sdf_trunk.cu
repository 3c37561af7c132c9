// sdf_trunk.cu -- backward of the SDF trunk's softplus + 256->1 SDF head in the delta basis of the tensor-core path
// (forward: tc_gemm_nt_persist_kernel<EPI_SDF_*> in gemm_tcgen05.cu).
//
// Reference (relative to /root/reference/): the autograd backward of projects/neuralangelo/utils/mlp.py:55-69
// (softplus(beta=100) hidden layer, linear_sdf fed from the layer's input) evaluated at the 1+taps stencil points of
// projects/neuralangelo/utils/modules.py:131-177.
//
// Basis.  The trunk's inputs are X_d = [x_0 ; x_i - x_0] (plane 0 = centre, plane i = tap i minus centre).  With the
// per-plane pre-activation gradients  e_0 = (g_0 w_sdf + dL/dh_0) s(z_0),  e_i = g_i w_sdf s(z_0 + dz_i)
// (s = sigmoid(100 .), g_p = dL/d sdf_p), the chain rule in that basis reads
//     dW0 = E_d^T X_d,   dX_d = E_d W0,   db0 = colsum(E),      E_d = [E = sum_p e_p ; e_i].
// The e_i are huge (g_i ~ +-1/(4 eps)) and cancel in E; forming E here in fp32 keeps that cancellation out of the
// bf16 GEMMs that follow: E is O(1), and the e_i only ever multiply the small deltas x_i - x_0 (or corner-weight
// differences in the hash-grid backward).
//
// Elementwise, HBM bound: per sample 256 x (4 B sigma0 + 2 B dH0 + 2 B h0 + taps x 2 B dz) in, (1+taps) x 512 B out.
// One thread = one row of a 128-row tile x one 8-column chunk (16-byte accesses, 2 KB contiguous per CTA).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kTile = 128;
constexpr int kChunks = 32;      // 256 hidden units
constexpr int kMaxTaps = 6;
constexpr int kMaxBlocksX = 74;  // 74 x 32 CTAs of 128 threads = 16 CTAs per SM, one wave

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}
// a < b ? x : y as a predicated select (both sides evaluated: keeps the 8 elements of a chunk in one basic block)
__device__ __forceinline__ float sel_lt(float a, float b, float x, float y) {
  float r;
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\tselp.f32 %0, %3, %4, p;\n\t}" : "=f"(r) : "f"(a), "f"(b), "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float fast_expm1(float t) {
  const float big = __expf(t) - 1.0f;
  const float small = t * (1.0f + t * (0.5f + t * (0.16666667f + t * 0.041666668f)));
  return sel_lt(fabsf(t), 0.03f, small, big);
}
__device__ __forceinline__ float fast_log1p(float q) {
  const float big = __logf(1.0f + q);
  const float small = q * (1.0f + q * (-0.5f + q * (0.33333334f - q * 0.25f)));
  return sel_lt(fabsf(q), 0.03f, small, big);
}

template <bool WITH_DW>
__global__ void __launch_bounds__(kTile) sdf_trunk_bwd_kernel(const float* __restrict__ g, int64_t M, int taps,
                                                              const float* __restrict__ sigma0,
                                                              const __nv_bfloat16* __restrict__ dz,
                                                              const __nv_bfloat16* __restrict__ dH0,
                                                              const __nv_bfloat16* __restrict__ h0,
                                                              const float* __restrict__ w_sdf,
                                                              __nv_bfloat16* __restrict__ Ed, float* __restrict__ part) {
  __shared__ float red[4][9];
  const int j = blockIdx.y;  // 8-column chunk
  const int r = threadIdx.x;
  const int n_tiles = (int)(M / kTile);
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = __ldg(w_sdf + j * 8 + i);
  float acc_w[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, acc_b = 0.0f;
  // Software pipeline: the loads of tile t + gridDim.x are issued before tile t is processed, so every thread always
  // has one tile's worth of 16-byte loads in flight while it does the MUFU-heavy math.
  struct In {
    float4 sa, sb;
    uint4 dh, h0v, dzr[kMaxTaps];
    float g[kMaxTaps + 1];
  };
  auto load = [&](int t, In& in) {
    const int64_t m = (int64_t)t * kTile + r;
    const int64_t e_idx = (((int64_t)t * kChunks + j) * kTile + r) * 8;
    const float* sp = sigma0 + (((int64_t)t * 64 + 2 * j) * kTile + r) * 4;
    in.sa = __ldg(reinterpret_cast<const float4*>(sp));
    in.sb = __ldg(reinterpret_cast<const float4*>(sp + kTile * 4));
    in.dh = dH0 ? __ldg(reinterpret_cast<const uint4*>(dH0 + e_idx)) : make_uint4(0u, 0u, 0u, 0u);
    if (WITH_DW) in.h0v = __ldg(reinterpret_cast<const uint4*>(h0 + e_idx));
#pragma unroll
    for (int p = 0; p < kMaxTaps; ++p)
      if (p < taps) in.dzr[p] = __ldg(reinterpret_cast<const uint4*>(dz + (int64_t)p * M * 256 + e_idx));
#pragma unroll
    for (int p = 0; p <= kMaxTaps; ++p)
      if (p <= taps) in.g[p] = __ldg(g + (int64_t)p * M + m);
  };
  In cur, nxt;
  if ((int)blockIdx.x < n_tiles) load(blockIdx.x, cur);
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const bool more = t + (int)gridDim.x < n_tiles;
    if (more) load(t + gridDim.x, nxt);
    const int64_t e_idx = (((int64_t)t * kChunks + j) * kTile + r) * 8;  // element offset inside a [M,256] TCL matrix
    const float s0[8] = {cur.sa.x, cur.sa.y, cur.sa.z, cur.sa.w, cur.sb.x, cur.sb.y, cur.sb.z, cur.sb.w};
    float dh0[8];
    unpack8(cur.dh, dh0);
    const float g0 = cur.g[0];
    float G = g0;
    float E[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) E[i] = (g0 * w[i] + dh0[i]) * s0[i];
    float hw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // sum_i g_i dh_i (for dw_sdf)
#pragma unroll
    for (int p = 0; p < kMaxTaps; ++p) {
      if (p < taps) {
        const float gp = cur.g[p + 1];
        G += gp;
        float d[8], e[8];
        unpack8(cur.dzr[p], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // sigmoid(100 (z0 + dz)) = et s0 / (1 + (et - 1) s0), et = exp(100 dz).  Plain MUFU exp / log suffice in the
          // backward: the absolute error of dh (~4e-9) times |g_i| (~1e3) stays ~1e-5 below the O(0.1..1) sums.
          const float t100 = fminf(fmaxf(100.0f * d[i], -80.0f), 80.0f);
          const float et = __expf(t100);
          const float num = et * s0[i];
          const float den = 1.0f - s0[i] + num;
          const float sp = __fdividef(num, den);
          e[i] = gp * w[i] * sp;
          E[i] += e[i];
          if (WITH_DW) hw[i] = fmaf(gp, __logf(den) * 0.01f, hw[i]);
        }
        *reinterpret_cast<uint4*>(Ed + (int64_t)(p + 1) * M * 256 + e_idx) = pack8(e);
      }
    }
    *reinterpret_cast<uint4*>(Ed + e_idx) = pack8(E);
    if (WITH_DW) {
      float h[8];
      unpack8(cur.h0v, h);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc_w[i] += fmaf(G, h[i], hw[i]);
      acc_b += G;
    }
    if (more) cur = nxt;
  }
  if (WITH_DW) {  // deterministic two-stage column reduction: CTA partials here, fixed-order sum in the second kernel
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc_w[i] += __shfl_xor_sync(0xffffffffu, acc_w[i], o);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc_b += __shfl_xor_sync(0xffffffffu, acc_b, o);
    if ((r & 31) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[r >> 5][i] = acc_w[i];
      red[r >> 5][8] = acc_b;
    }
    __syncthreads();
    if (r < 9) part[((size_t)blockIdx.x * kChunks + j) * 9 + r] = red[0][r] + red[1][r] + red[2][r] + red[3][r];
  }
}

__global__ void sdf_trunk_bwd_reduce_kernel(const float* __restrict__ part, int S, float* __restrict__ dw, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;  // 0..255 columns, 256 = bias
  if (c > 256) return;
  const int j = c < 256 ? c >> 3 : 0, i = c < 256 ? c & 7 : 8;
  float v = 0.0f;
  for (int s = 0; s < S; ++s) v += part[((size_t)s * kChunks + j) * 9 + i];
  if (c < 256) dw[c] = v;
  else if (db) db[0] = v;
}

int blocks_x(int64_t M) {
  const int64_t n_tiles = M / kTile;
  return (int)(n_tiles < kMaxBlocksX ? n_tiles : kMaxBlocksX);
}

}  // namespace

extern "C" int64_t mli_tc_sdf_trunk_bwd_ws_bytes(int64_t M) {
  (void)M;
  return (int64_t)kMaxBlocksX * kChunks * 9 * sizeof(float);
}

extern "C" int mli_tc_sdf_trunk_bwd(const float* g, int64_t M, int32_t taps, const float* sigma0, const void* dz,
                                    const void* dH0, const void* h0, const float* w_sdf, void* Ed, float* dw_sdf,
                                    float* db_sdf, void* ws, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(M >= kTile && M % kTile == 0, "tc_sdf_trunk_bwd: M must be a positive multiple of 128");
  MLI_REQUIRE(taps == 4 || taps == 6, "Only support 4 or 6 taps.");
  MLI_REQUIRE(g && sigma0 && dz && w_sdf && Ed, "tc_sdf_trunk_bwd: NULL argument");
  MLI_REQUIRE(dw_sdf == nullptr || (h0 != nullptr && ws != nullptr), "tc_sdf_trunk_bwd: dw_sdf needs h0 and a workspace");
  dim3 grid(blocks_x(M), kChunks);
  cudaStream_t st = (cudaStream_t)stream;
  if (dw_sdf) {
    sdf_trunk_bwd_kernel<true><<<grid, kTile, 0, st>>>(g, M, taps, sigma0, (const __nv_bfloat16*)dz, (const __nv_bfloat16*)dH0,
                                                       (const __nv_bfloat16*)h0, w_sdf, (__nv_bfloat16*)Ed, (float*)ws);
    MLI_LAUNCH_OK();
    sdf_trunk_bwd_reduce_kernel<<<2, 256, 0, st>>>((const float*)ws, (int)grid.x, dw_sdf, db_sdf);
  } else {
    sdf_trunk_bwd_kernel<false><<<grid, kTile, 0, st>>>(g, M, taps, sigma0, (const __nv_bfloat16*)dz, (const __nv_bfloat16*)dH0,
                                                        nullptr, w_sdf, (__nv_bfloat16*)Ed, nullptr);
  }
  MLI_LAUNCH_OK();
  return MLI_OK;
}

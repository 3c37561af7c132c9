// tcl_ops.cu -- narrow output layers on TCL (bf16 tile-chunk layout) activations, the companions of gemm_tcgen05.cu.
//
// Reference: /root/reference/projects/nerf/utils/nerf_util.py:186-196 (output layer of each head, 256 -> 3/3/1) and
// /root/reference/projects/neuralangelo/utils/mlp.py:50,66 (linear_sdf).  One thread per row of a 128-row tile: a
// chunk load is 16 B per thread and 2 KB contiguous per CTA (fully coalesced); dot products stay in registers, so no
// shuffles are needed.  HBM-bandwidth bound (one pass over the [M, 256*nh] bf16 activations).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kMaxJ = 12;  // r_s_re: 3 heads x 3 channels = 9 narrow outputs
constexpr int kTile = 128;

struct RdArgs {
  int32_t col_off[kMaxJ];
  int32_t J, K;
};

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

// out[m, j] = act_j(sum_k A[m, col_off[j]+k] w[j,k] + b[j])
__global__ void __launch_bounds__(kTile) rowdot_tcl_fwd_kernel(const __nv_bfloat16* __restrict__ A, int a_chunks, int64_t M,
                                                               const float* __restrict__ w, const float* __restrict__ b,
                                                               RdArgs a, int act, uint32_t act_mask,
                                                               float* __restrict__ out, int64_t ldo) {
  extern __shared__ float sw[];  // [J][K]
  for (int i = threadIdx.x; i < a.J * a.K; i += kTile) sw[i] = w[i];
  __syncthreads();
  const int64_t tile = blockIdx.x;
  const int64_t m = tile * kTile + threadIdx.x;
  float acc[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) acc[j] = 0.0f;
  int j0 = 0;
  while (j0 < a.J) {  // outputs that share a column offset (= one head) reuse each chunk load
    int j1 = j0 + 1;
    while (j1 < a.J && a.col_off[j1] == a.col_off[j0]) ++j1;
    const __nv_bfloat16* src = A + ((tile * a_chunks + a.col_off[j0] / 8) * kTile + threadIdx.x) * 8;
    for (int c = 0; c < a.K / 8; ++c) {
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (int64_t)c * kTile * 8)), x);
#pragma unroll
      for (int j = 0; j < kMaxJ; ++j) {
        if (j >= j0 && j < j1) {
          const float* ww = sw + j * a.K + c * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j] = fmaf(x[i], ww[i], acc[j]);
        }
      }
    }
    j0 = j1;
  }
  if (m < M) {
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j)
      if (j < a.J) out[m * ldo + j] = mli_act(acc[j] + (b ? b[j] : 0.0f), ((act_mask >> j) & 1u) ? act : MLI_ACT_NONE);
  }
}

// dZ[m, c] = (sum_{j covering c} dS[m,j] w[j, c - off_j]) * act'(A[m,c])   (bf16 TCL out, same chunk geometry as A)
// One thread = one row of a 128-row tile; it keeps the row's J upstream gradients in registers and walks 32 chunks,
// so every A load / dZ store is 16 B per thread, 2 KB contiguous per CTA, with 32 independent loads in flight.
constexpr int kBwdChunksPerCta = 32;

__global__ void __launch_bounds__(kTile) rowdot_tcl_bwd_kernel(const float* __restrict__ dS, int64_t lds,
                                                               const __nv_bfloat16* __restrict__ A, int a_chunks, int64_t M,
                                                               const float* __restrict__ w, RdArgs a, int act_prev,
                                                               __nv_bfloat16* __restrict__ dZ,
                                                               const uint32_t* __restrict__ relu_mask) {
  extern __shared__ float sw[];  // [J][K]
  for (int i = threadIdx.x; i < a.J * a.K; i += kTile) sw[i] = w[i];
  __syncthreads();
  const int64_t tile = blockIdx.x;
  const int64_t m = tile * kTile + threadIdx.x;
  const int c_begin = blockIdx.y * kBwdChunksPerCta;
  const int c_end = min(a_chunks, c_begin + kBwdChunksPerCta);
  float d[kMaxJ];
#pragma unroll
  for (int j = 0; j < kMaxJ; ++j) d[j] = (j < a.J && m < M) ? dS[m * lds + j] : 0.0f;
#pragma unroll 4
  for (int c = c_begin; c < c_end; ++c) {
    const int64_t idx = ((tile * a_chunks + c) * kTile + threadIdx.x) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < kMaxJ; ++j) {
      const int k = c * 8 - a.col_off[j];  // warp-uniform
      if (j < a.J && k >= 0 && k < a.K) {
        const float* ww = sw + j * a.K + k;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(d[j], ww[i], v[i]);
      }
    }
    if (act_prev == MLI_ACT_RELU && relu_mask) {  // sign bits instead of the bf16 activations
      const uint32_t bits = __ldg(relu_mask + ((tile * (a_chunks / 4) + (c >> 2)) * kTile + threadIdx.x)) >> ((c & 3) * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ((bits >> i) & 1u) ? v[i] : 0.0f;
    } else if (act_prev != MLI_ACT_NONE) {
      float y[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(A + idx)), y);
      if (act_prev == MLI_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = y[i] > 0.0f ? v[i] : 0.0f;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= mli_dact_from_out(y[i], act_prev);
      }
    }
    *reinterpret_cast<uint4*>(dZ + idx) = pack8(v);  // padding rows are written as zeros (d = 0)
  }
}

int make_args(RdArgs* a, const int32_t* col_off, int32_t J, int32_t K) {
  MLI_REQUIRE(J >= 1 && J <= kMaxJ, "rowdot_tcl: J must be in 1..12");
  MLI_REQUIRE(K >= 8 && K % 8 == 0 && K <= 1024, "rowdot_tcl: K must be a multiple of 8 (<= 1024)");
  a->J = J; a->K = K;
  for (int j = 0; j < kMaxJ; ++j) {
    a->col_off[j] = (col_off && j < J) ? col_off[j] : 0;
    MLI_REQUIRE(a->col_off[j] % 8 == 0 && a->col_off[j] >= 0, "rowdot_tcl: col_off must be non-negative multiples of 8");
    if (j > 0 && j < J) MLI_REQUIRE(a->col_off[j] >= a->col_off[j - 1], "rowdot_tcl: col_off must be non-decreasing");
  }
  return MLI_OK;
}

}  // namespace

extern "C" int mli_tc_rowdot_fwd(const void* A, int32_t a_chunks, int64_t M, const float* w, const float* b,
                                 const int32_t* host_col_off, int32_t J, int32_t K, int32_t act, uint32_t act_mask,
                                 float* out, int64_t ldo, void* stream) {
  MLI_ENTRY();
  RdArgs a;
  if (int e = make_args(&a, host_col_off, J, K)) return e;
  MLI_REQUIRE(ldo >= J, "rowdot_tcl: ldo < J");
  if (M <= 0) return MLI_OK;
  rowdot_tcl_fwd_kernel<<<mli_cdiv(M, kTile), kTile, (size_t)J * K * sizeof(float), (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)A, a_chunks, M, w, b, a, act, act_mask, out, ldo);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_tc_rowdot_bwd_data(const float* dS, int64_t lds, const void* A, int32_t a_chunks, int64_t M,
                                      const float* w, const int32_t* host_col_off, int32_t J, int32_t K, int32_t act_prev,
                                      void* dZ, const void* relu_mask, void* stream) {
  MLI_ENTRY();
  RdArgs a;
  if (int e = make_args(&a, host_col_off, J, K)) return e;
  MLI_REQUIRE(lds >= J, "rowdot_tcl: lds < J");
  MLI_REQUIRE(relu_mask == nullptr || a_chunks % 4 == 0, "rowdot_tcl: relu_mask needs a_chunks % 4 == 0");
  if (M <= 0) return MLI_OK;
  dim3 grid(mli_cdiv(M, kTile), mli_cdiv(a_chunks, kBwdChunksPerCta));
  rowdot_tcl_bwd_kernel<<<grid, kTile, (size_t)J * K * sizeof(float), (cudaStream_t)stream>>>(
      dS, lds, (const __nv_bfloat16*)A, a_chunks, M, w, a, act_prev, (__nv_bfloat16*)dZ, (const uint32_t*)relu_mask);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

// rowdot.cu -- narrow output layers (SDF head 256->1, colour/intrinsic output layers 256->3/3/1),
// weight_norm packing and the dense AdamW step.
//
// Reference: /root/reference/projects/neuralangelo/utils/mlp.py:42-50,66 (linear_sdf, weight_norm),
// /root/reference/projects/nerf/utils/nerf_util.py:177-178,191 (head output layer),
// /root/reference/imaginaire/trainers/utils/get_trainer.py:106-150 (AdamW).
// All of these are HBM-bandwidth bound (one pass over the [M,K] activations).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int kMaxJ = 8;

struct RowdotArgs {
  int32_t col_off[kMaxJ];
  int32_t J, K;
};

// one warp per row; lane covers 4 consecutive k per 128-wide chunk (coalesced LDG.128)
__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const float* __restrict__ A, int64_t lda, int64_t M,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         RowdotArgs a, int act, uint32_t act_mask,
                                                         float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  const float* row = A + m * lda;
  for (int j = 0; j < a.J; ++j) {
    float s = 0.0f;
    for (int k = lane * 4; k < a.K; k += 128) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(row + a.col_off[j] + k));
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w + j * a.K + k));
      s = fmaf(x.x, ww.x, s); s = fmaf(x.y, ww.y, s); s = fmaf(x.z, ww.z, s); s = fmaf(x.w, ww.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[m * ldo + j] = mli_act(s + (b ? b[j] : 0.0f), ((act_mask >> j) & 1u) ? act : MLI_ACT_NONE);
  }
}

// dA[m,c] = (accumulate ? dA : 0) + sum_{j covering c} dS[m,j] w[j,c-off_j]; then * act'(A[m,c])
__global__ void __launch_bounds__(256) rowdot_bwd_data_kernel(const float* __restrict__ dS, int64_t lds,
                                                              const float* __restrict__ A, int64_t lda, int64_t M,
                                                              const float* __restrict__ w, RowdotArgs a,
                                                              int act_prev, int accumulate, float* __restrict__ dA,
                                                              int64_t ldda, int n_cols) {
  const int cols4 = n_cols >> 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * cols4) return;
  const int64_t m = e / cols4;
  const int c = (int)(e % cols4) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (accumulate) {
    const float4 o = *reinterpret_cast<const float4*>(dA + m * ldda + c);
    v[0] = o.x; v[1] = o.y; v[2] = o.z; v[3] = o.w;
  }
  for (int j = 0; j < a.J; ++j) {
    const int k = c - a.col_off[j];
    if (k < 0 || k >= a.K) continue;
    const float d = dS[m * lds + j];
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w + j * a.K + k));
    v[0] = fmaf(d, ww.x, v[0]); v[1] = fmaf(d, ww.y, v[1]); v[2] = fmaf(d, ww.z, v[2]); v[3] = fmaf(d, ww.w, v[3]);
  }
  if (act_prev != MLI_ACT_NONE) {
    const float4 y = __ldg(reinterpret_cast<const float4*>(A + m * lda + c));
    v[0] *= mli_dact_from_out(y.x, act_prev); v[1] *= mli_dact_from_out(y.y, act_prev);
    v[2] *= mli_dact_from_out(y.z, act_prev); v[3] *= mli_dact_from_out(y.w, act_prev);
  }
  *reinterpret_cast<float4*>(dA + m * ldda + c) = make_float4(v[0], v[1], v[2], v[3]);
}

// partial dw[j,k] / db[j] over a contiguous chunk of rows; blockDim.x == K.  Rows are processed 8 at a time with all
// loads issued before the FMAs so that 8*J independent 1 KB row reads are in flight per CTA (the loop is otherwise one
// dependent global load per iteration).
template <int J>
__global__ void rowdot_bwd_weight_kernel(const float* __restrict__ dS, int64_t lds, const float* __restrict__ A,
                                         int64_t lda, int64_t M, RowdotArgs a, int64_t rows_per_block,
                                         float* __restrict__ part) {
  constexpr int U = 8;
  const int k = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[J], accb[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = accb[j] = 0.0f;
  int64_t m = m0;
  for (; m + U <= m1; m += U) {
    float d[U][J], x[U][J];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        d[u][j] = __ldg(dS + (m + u) * lds + j);
        x[u][j] = __ldg(A + (m + u) * lda + a.col_off[j] + k);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < J; ++j) { acc[j] = fmaf(d[u][j], x[u][j], acc[j]); accb[j] += d[u][j]; }
  }
  for (; m < m1; ++m)
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float d = __ldg(dS + m * lds + j);
      acc[j] = fmaf(d, __ldg(A + m * lda + a.col_off[j] + k), acc[j]);
      accb[j] += d;
    }
  float* dst = part + (size_t)blockIdx.x * J * (a.K + 1);
#pragma unroll
  for (int j = 0; j < J; ++j) {
    dst[j * (a.K + 1) + k] = acc[j];
    if (k == 0) dst[j * (a.K + 1) + a.K] = accb[j];
  }
}

__global__ void rowdot_bwd_reduce_kernel(const float* __restrict__ part, int n_blocks, int J, int K,
                                         float* __restrict__ dw, float* __restrict__ db) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= J * (K + 1)) return;
  float v = 0.0f;
  for (int s = 0; s < n_blocks; ++s) v += part[(size_t)s * J * (K + 1) + e];
  const int j = e / (K + 1), k = e % (K + 1);
  if (k < K) dw[j * K + k] = v;
  else if (db) db[j] = v;
}

constexpr int kRowdotBlocks = 8 * MLI_NUM_SMS;

// ---------------------------------------------------------------------------------------------------------
// weight_norm: one CTA per output row
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) weightnorm_pack_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              int N, int K, const int32_t* __restrict__ col_map,
                                                              float* __restrict__ Wp, int64_t ldw,
                                                              float* __restrict__ Wpt, int64_t ldwt, int row_off) {
  __shared__ float red[32];
  __shared__ float s_scale;
  const int n = blockIdx.x;
  float ss = 0.0f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { const float x = v[(int64_t)n * K + k]; ss = fmaf(x, x, ss); }
  ss = mli_block_sum(ss, red);
  if (threadIdx.x == 0) s_scale = g[n] / sqrtf(ss);
  __syncthreads();
  const float scale = s_scale;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = v[(int64_t)n * K + k] * scale;
    const int c = col_map ? col_map[k] : k;
    Wp[(int64_t)(row_off + n) * ldw + c] = wv;
    if (Wpt) Wpt[(int64_t)c * ldwt + row_off + n] = wv;
  }
}

__global__ void __launch_bounds__(128) weightnorm_grad_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              const float* __restrict__ dWp, int64_t ldw, int N, int K,
                                                              const int32_t* __restrict__ col_map, int row_off,
                                                              float* __restrict__ dv, float* __restrict__ dg) {
  __shared__ float red[32];
  __shared__ float s_ss, s_dot;
  const int n = blockIdx.x;
  float ss = 0.0f, dot = 0.0f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float x = v[(int64_t)n * K + k];
    const float d = dWp[(int64_t)(row_off + n) * ldw + (col_map ? col_map[k] : k)];
    ss = fmaf(x, x, ss);
    dot = fmaf(x, d, dot);
  }
  ss = mli_block_sum(ss, red);
  if (threadIdx.x == 0) s_ss = ss;
  dot = mli_block_sum(dot, red);
  if (threadIdx.x == 0) s_dot = dot;
  __syncthreads();
  const float norm = sqrtf(s_ss), gn = g[n];
  // W = g v/|v|:  dg = <dW, v>/|v| ;  dv = g/|v| dW - g <dW,v>/|v|^3 v
  if (threadIdx.x == 0) dg[n] = s_dot / norm;
  const float c1 = gn / norm, c2 = gn * s_dot / (norm * norm * norm);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float d = dWp[(int64_t)(row_off + n) * ldw + (col_map ? col_map[k] : k)];
    dv[(int64_t)n * K + k] = c1 * d - c2 * v[(int64_t)n * K + k];
  }
}

// ---------------------------------------------------------------------------------------------------------
// batched weight_norm: every matrix of the step in one launch (descriptors travel as kernel parameters)
// ---------------------------------------------------------------------------------------------------------
struct WnBatch {
  mli_wn_desc_t d[MLI_WN_MAX_DESCS];
  int n;
};

__device__ __forceinline__ int64_t tcl_index(int64_t R, int C, int tile, int chunks) {
  return (((R / tile) * chunks + (C >> 3)) * tile + (R % tile)) * 8 + (C & 7);
}

__global__ void __launch_bounds__(128) weightnorm_pack_batch_kernel(const __grid_constant__ WnBatch b) {
  __shared__ float red[32];
  __shared__ float s_scale;
  int di = 0;
  while (di + 1 < b.n && (int)blockIdx.x >= b.d[di + 1].row_begin) ++di;
  const mli_wn_desc_t& d = b.d[di];
  const int n = blockIdx.x - d.row_begin;
  const float* vr = d.v + (int64_t)n * d.K;
  float ss = 0.0f;
  for (int k = threadIdx.x; k < d.K; k += blockDim.x) { const float x = vr[k]; ss = fmaf(x, x, ss); }
  ss = mli_block_sum(ss, red);
  if (threadIdx.x == 0) s_scale = d.g[n] / sqrtf(ss);
  __syncthreads();
  const float scale = s_scale;
  const int64_t R = d.row_off + n;
  for (int k = threadIdx.x; k < d.K; k += blockDim.x) {
    const float wv = vr[k] * scale;
    const int c = d.col_map ? d.col_map[k] : k;
    if (d.Wp) d.Wp[R * d.ldw + c] = wv;
    const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
    if (d.tcl) {
      __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(d.tcl);
      t[tcl_index(R, c, d.tcl_tile, d.tcl_chunks)] = hi;
      if (d.tcl_lo >= 0) t[tcl_index(R, c + 8 * d.tcl_lo, d.tcl_tile, d.tcl_chunks)] = __float2bfloat16_rn(wv - __bfloat162float(hi));
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (d.tclt[q] && c >= d.tclt_c0[q] && c < d.tclt_c1[q])
        reinterpret_cast<__nv_bfloat16*>(d.tclt[q])[tcl_index(d.tclt_row_off[q] + c - d.tclt_c0[q], d.tclt_col_off[q] + n,
                                                              d.tclt_tile[q], d.tclt_chunks[q])] = hi;
    }
  }
}

__global__ void __launch_bounds__(128) weightnorm_grad_batch_kernel(const __grid_constant__ WnBatch b) {
  __shared__ float red[32];
  __shared__ float s_ss, s_dot;
  int di = 0;
  while (di + 1 < b.n && (int)blockIdx.x >= b.d[di + 1].row_begin) ++di;
  const mli_wn_desc_t& d = b.d[di];
  const int n = blockIdx.x - d.row_begin;
  const float* vr = d.v + (int64_t)n * d.K;
  const float* dr = d.dWp + (int64_t)(d.row_off + n) * d.ldw;
  float ss = 0.0f, dot = 0.0f;
  for (int k = threadIdx.x; k < d.K; k += blockDim.x) {
    const float x = vr[k];
    const float g = dr[d.col_map ? d.col_map[k] : k];
    ss = fmaf(x, x, ss);
    dot = fmaf(x, g, dot);
  }
  ss = mli_block_sum(ss, red);
  if (threadIdx.x == 0) s_ss = ss;
  dot = mli_block_sum(dot, red);
  if (threadIdx.x == 0) s_dot = dot;
  __syncthreads();
  const float norm = sqrtf(s_ss), gn = d.g[n];
  if (threadIdx.x == 0) d.dg[n] = s_dot / norm;
  const float c1 = gn / norm, c2 = gn * s_dot / (norm * norm * norm);
  for (int k = threadIdx.x; k < d.K; k += blockDim.x)
    d.dv[(int64_t)n * d.K + k] = c1 * dr[d.col_map ? d.col_map[k] : k] - c2 * vr[k];
}

static int make_wn_batch(WnBatch* b, const mli_wn_desc_t* descs, int32_t n, int* total_rows, bool backward) {
  MLI_REQUIRE(descs != nullptr && n >= 1 && n <= MLI_WN_MAX_DESCS, "weightnorm batch: 1..%d descriptors", MLI_WN_MAX_DESCS);
  int rows = 0;
  for (int i = 0; i < n; ++i) {
    b->d[i] = descs[i];
    MLI_REQUIRE(descs[i].v && descs[i].g && descs[i].N >= 1 && descs[i].K >= 1, "weightnorm batch: bad descriptor %d", i);
    if (backward) MLI_REQUIRE(descs[i].dWp && descs[i].dv && descs[i].dg, "weightnorm batch: descriptor %d has no gradient buffers", i);
    b->d[i].row_begin = rows;
    rows += descs[i].N;
  }
  b->n = n;
  *total_rows = rows;
  return MLI_OK;
}

// ---------------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW single-tensor semantics), gradient pre-scaled by grad_scale (1/world_size)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m1, float* __restrict__ m2, int64_t n4,
                                                    int64_t n, float lr, float b1, float b2, float eps, float wd,
                                                    float bc1, float bc2_sqrt, float gs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 a = reinterpret_cast<float4*>(m1)[i], v = reinterpret_cast<float4*>(m2)[i];
    float* P = &pp.x; const float* G = &gg.x; float* A = &a.x; float* V = &v.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float gr = G[c] * gs;
      P[c] *= 1.0f - lr * wd;
      A[c] = A[c] + (gr - A[c]) * (1.0f - b1);  // lerp form used by torch
      V[c] = V[c] * b2 + (1.0f - b2) * gr * gr;
      const float denom = sqrtf(V[c]) / bc2_sqrt + eps;
      P[c] -= (lr / bc1) * (A[c] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m1)[i] = a;
    reinterpret_cast<float4*>(m2)[i] = v;
  }
  // scalar tail
  const int64_t t = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    const float gr = g[t] * gs;
    float pv = p[t] * (1.0f - lr * wd);
    const float a = m1[t] + (gr - m1[t]) * (1.0f - b1);
    const float v = m2[t] * b2 + (1.0f - b2) * gr * gr;
    pv -= (lr / bc1) * (a / (sqrtf(v) / bc2_sqrt + eps));
    p[t] = pv; m1[t] = a; m2[t] = v;
  }
}

int make_args(RowdotArgs* a, const int32_t* col_off, int32_t J, int32_t K) {
  MLI_REQUIRE(J >= 1 && J <= kMaxJ, "rowdot: J must be in 1..8");
  MLI_REQUIRE(K >= 128 && K % 128 == 0 && K <= 1024, "rowdot: K must be a multiple of 128 (<= 1024)");
  a->J = J; a->K = K;
  for (int j = 0; j < kMaxJ; ++j) {
    a->col_off[j] = (col_off && j < J) ? col_off[j] : 0;
    MLI_REQUIRE(a->col_off[j] % 4 == 0 && a->col_off[j] >= 0, "rowdot: col_off must be non-negative multiples of 4");
  }
  return MLI_OK;
}

}  // namespace

extern "C" int mli_rowdot_fwd(const float* A, int64_t lda, int64_t M, const float* w, const float* b,
                              const int32_t* col_off, int32_t J, int32_t K, int32_t act, uint32_t act_mask, float* out,
                              int64_t ldo, void* stream) {
  MLI_ENTRY();
  RowdotArgs a;
  if (int e = make_args(&a, col_off, J, K)) return e;
  MLI_REQUIRE(lda % 4 == 0 && ldo >= J, "rowdot: bad leading dimensions");
  if (M <= 0) return MLI_OK;
  rowdot_fwd_kernel<<<mli_cdiv(M, 8), 256, 0, (cudaStream_t)stream>>>(A, lda, M, w, b, a, act, act_mask, out, ldo);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int64_t mli_rowdot_bwd_ws_bytes(int64_t M, int32_t J, int32_t K) {
  (void)M;
  return (int64_t)kRowdotBlocks * J * (K + 1) * sizeof(float);
}

extern "C" int mli_rowdot_bwd(const float* dS, int64_t lds, const float* A, int64_t lda, int64_t M, const float* w,
                              const int32_t* col_off, int32_t J, int32_t K, int32_t act_prev, float* dA,
                              int64_t ldda, int32_t n_cols_dA, int32_t accumulate, float* dw, float* db, void* ws,
                              void* stream) {
  MLI_ENTRY();
  RowdotArgs a;
  if (int e = make_args(&a, col_off, J, K)) return e;
  MLI_REQUIRE(lda % 4 == 0 && lds >= J, "rowdot: bad leading dimensions");
  if (M <= 0) return MLI_OK;
  const int n_cols = n_cols_dA;
  if (dA) {
    MLI_REQUIRE(n_cols % 4 == 0 && ldda % 4 == 0, "rowdot: dA columns must be a multiple of 4");
    rowdot_bwd_data_kernel<<<mli_cdiv(M * (n_cols / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        dS, lds, A, lda, M, w, a, act_prev, accumulate, dA, ldda, n_cols);
    MLI_LAUNCH_OK();
  }
  if (dw) {
    MLI_REQUIRE(ws != nullptr, "rowdot: workspace is NULL");
    int blocks = kRowdotBlocks;
    int64_t rows = (M + blocks - 1) / blocks;
    blocks = (int)((M + rows - 1) / rows);
#define LAUNCH_W(JJ) rowdot_bwd_weight_kernel<JJ><<<blocks, K, 0, (cudaStream_t)stream>>>(dS, lds, A, lda, M, a, rows, (float*)ws)
    switch (J) {
      case 1: LAUNCH_W(1); break; case 2: LAUNCH_W(2); break; case 3: LAUNCH_W(3); break; case 4: LAUNCH_W(4); break;
      case 5: LAUNCH_W(5); break; case 6: LAUNCH_W(6); break; case 7: LAUNCH_W(7); break; default: LAUNCH_W(8); break;
    }
#undef LAUNCH_W
    MLI_LAUNCH_OK();
    rowdot_bwd_reduce_kernel<<<mli_cdiv(J * (K + 1), 256), 256, 0, (cudaStream_t)stream>>>((const float*)ws, blocks, J,
                                                                                          K, dw, db);
    MLI_LAUNCH_OK();
  }
  return MLI_OK;
}

extern "C" int mli_weightnorm_pack(const float* v, const float* g, int32_t N, int32_t K, const int32_t* col_map,
                                   float* Wp, int64_t ldw, float* Wpt, int64_t ldwt, int32_t row_off, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(N >= 1 && K >= 1, "weightnorm: bad shape");
  weightnorm_pack_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(v, g, N, K, col_map, Wp, ldw, Wpt, ldwt, row_off);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_weightnorm_unpack_grad(const float* v, const float* g, const float* dWp, int64_t ldw, int32_t N,
                                          int32_t K, const int32_t* col_map, int32_t row_off, float* dv, float* dg,
                                          void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(N >= 1 && K >= 1, "weightnorm: bad shape");
  weightnorm_grad_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(v, g, dWp, ldw, N, K, col_map, row_off, dv, dg);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                              float grad_scale, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(n >= 0 && step >= 1, "adamw: bad n/step");
  MLI_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "adamw: buffers must be 16-byte aligned");
  if (n == 0) return MLI_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 16 * MLI_NUM_SMS) blocks = 16 * MLI_NUM_SMS;
  if (blocks < 1) blocks = 1;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4, n, lr, beta1,
                                                                  beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_weightnorm_pack_batch(const mli_wn_desc_t* descs_on_host, int32_t n_descs, void* stream) {
  MLI_ENTRY();
  WnBatch b;
  int rows = 0;
  if (int e = make_wn_batch(&b, descs_on_host, n_descs, &rows, false)) return e;
  weightnorm_pack_batch_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(b);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_weightnorm_unpack_grad_batch(const mli_wn_desc_t* descs_on_host, int32_t n_descs, void* stream) {
  MLI_ENTRY();
  WnBatch b;
  int rows = 0;
  if (int e = make_wn_batch(&b, descs_on_host, n_descs, &rows, true)) return e;
  weightnorm_grad_batch_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(b);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

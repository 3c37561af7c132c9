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

constexpr int kMaxJ = 12;  // r_s_re: 3 heads x 3 channels = 9 narrow outputs

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
    if (d.tcl2) reinterpret_cast<__nv_bfloat16*>(d.tcl2)[tcl_index(R, c, d.tcl2_tile, d.tcl2_chunks)] = hi;
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
struct AdamHyper {
  // every derived constant is formed in double on the host, as torch forms them from the Python floats, and rounded once:
  // omb1 = 1 - beta1, omb2 = 1 - beta2 (NOT 1 - float(beta2): 4.7e-5 relative apart for 0.999), decay = 1 - lr * wd,
  // step_size = lr / (1 - beta1^t), inv_bc2_sqrt = 1 / sqrt(1 - beta2^t)
  float b2, omb1, omb2, eps, decay, step_size, inv_bc2_sqrt, gs;
};

static AdamHyper make_adam_hyper(double lr, double b1, double b2, double eps, double wd, int step, double gs) {
  AdamHyper h;
  h.b2 = (float)b2; h.omb1 = (float)(1.0 - b1); h.omb2 = (float)(1.0 - b2); h.eps = (float)eps;
  h.decay = (float)(1.0 - lr * wd); h.gs = (float)gs;
  h.step_size = (float)(lr / (1.0 - pow(b1, (double)step)));
  h.inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow(b2, (double)step)));
  return h;
}

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, const AdamHyper& h) {
  const float gr = g * h.gs;
  p *= h.decay;
  m = m + (gr - m) * h.omb1;  // lerp form used by torch
  v = v * h.b2 + h.omb2 * gr * gr;
  // 84 % of the hash-table entries see a zero gradient in a step and 38 % still have v == 0: the IEEE sqrtf / division
  // take their special-case slow paths for those (measured: 2.75 ms instead of 1.5 ms for the table), so both use the
  // branch-free MUFU forms (<= 2 ulp each; denom >= eps is always in __fdividef's range)
  const float denom = fast_sqrt(v) * h.inv_bc2_sqrt + h.eps;
  p -= h.step_size * __fdividef(m, denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m1, float* __restrict__ m2, int64_t n4,
                                                    int64_t n, const AdamHyper h) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 a = reinterpret_cast<float4*>(m1)[i], v = reinterpret_cast<float4*>(m2)[i];
    adamw1(pp.x, gg.x, a.x, v.x, h);
    adamw1(pp.y, gg.y, a.y, v.y, h);
    adamw1(pp.z, gg.z, a.z, v.z, h);
    adamw1(pp.w, gg.w, a.w, v.w, h);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m1)[i] = a;
    reinterpret_cast<float4*>(m2)[i] = v;
  }
  const int64_t t = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // scalar tail
  if (t < n) {
    float pv = p[t], a = m1[t], v = m2[t];
    adamw1(pv, g[t], a, v, h);
    p[t] = pv; m1[t] = a; m2[t] = v;
  }
}

// every parameter tensor of an optimizer group in ONE launch: descriptors + block prefix travel as kernel parameters.
// One CTA = one 4096-element chunk of one tensor (4 float4 per thread and buffer -> 16 independent 128-bit loads in
// flight per thread before the first use).
struct AdamBatch {
  mli_adamw_desc_t d[MLI_ADAMW_MAX_TENSORS];
  uint32_t first_block[MLI_ADAMW_MAX_TENSORS + 1];
  int n;
};
constexpr int kAdamChunk = 4096;

__global__ void __launch_bounds__(256) adamw_batch_kernel(const __grid_constant__ AdamBatch b, const AdamHyper h) {
  int lo = 0, hi = b.n - 1;  // tensor of this block: last t with first_block[t] <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.first_block[mid] <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const mli_adamw_desc_t& d = b.d[lo];
  const int64_t e0 = (int64_t)(blockIdx.x - b.first_block[lo]) * kAdamChunk;
  const int64_t left = d.n - e0;
  float* __restrict__ P = d.param + e0;
  const float* __restrict__ G = d.grad + e0;
  float* __restrict__ M1 = d.exp_avg + e0;
  float* __restrict__ M2 = d.exp_avg_sq + e0;
  const bool aligned = (((uintptr_t)d.param | (uintptr_t)d.grad | (uintptr_t)d.exp_avg | (uintptr_t)d.exp_avg_sq) & 15) == 0;
  if (aligned && left >= kAdamChunk) {
    float4 pp[4], gg[4], aa[4], vv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = threadIdx.x + k * 256;
      pp[k] = reinterpret_cast<float4*>(P)[i];
      gg[k] = __ldg(reinterpret_cast<const float4*>(G) + i);
      aa[k] = reinterpret_cast<float4*>(M1)[i];
      vv[k] = reinterpret_cast<float4*>(M2)[i];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = threadIdx.x + k * 256;
      adamw1(pp[k].x, gg[k].x, aa[k].x, vv[k].x, h);
      adamw1(pp[k].y, gg[k].y, aa[k].y, vv[k].y, h);
      adamw1(pp[k].z, gg[k].z, aa[k].z, vv[k].z, h);
      adamw1(pp[k].w, gg[k].w, aa[k].w, vv[k].w, h);
      reinterpret_cast<float4*>(P)[i] = pp[k];
      reinterpret_cast<float4*>(M1)[i] = aa[k];
      reinterpret_cast<float4*>(M2)[i] = vv[k];
    }
  } else {  // ragged last chunk / unaligned small tensors (biases of width 1 or 3, s_var)
    const int cnt = (int)(left < kAdamChunk ? left : kAdamChunk);
    for (int i = threadIdx.x; i < cnt; i += 256) {
      float p = P[i], m = M1[i], v = M2[i];
      adamw1(p, G[i], m, v, h);
      P[i] = p; M1[i] = m; M2[i] = v;
    }
  }
}

int make_args(RowdotArgs* a, const int32_t* col_off, int32_t J, int32_t K) {
  MLI_REQUIRE(J >= 1 && J <= kMaxJ, "rowdot: J must be in 1..12");
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
      case 5: LAUNCH_W(5); break; case 6: LAUNCH_W(6); break; case 7: LAUNCH_W(7); break; case 8: LAUNCH_W(8); break;
      case 9: LAUNCH_W(9); break; case 10: LAUNCH_W(10); break; case 11: LAUNCH_W(11); break; default: LAUNCH_W(12); break;
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
                              double lr, double beta1, double beta2, double eps, double weight_decay, int32_t step,
                              double grad_scale, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(n >= 0 && step >= 1, "adamw: bad n/step");
  MLI_REQUIRE(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "adamw: buffers must be 16-byte aligned");
  if (n == 0) return MLI_OK;
  const AdamHyper h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 16 * MLI_NUM_SMS) blocks = 16 * MLI_NUM_SMS;
  if (blocks < 1) blocks = 1;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4, n, h);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_adamw_step_batch(const mli_adamw_desc_t* descs_on_host, int32_t n_descs, double lr, double beta1,
                                    double beta2, double eps, double weight_decay, int32_t step, double grad_scale,
                                    void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(descs_on_host != nullptr && n_descs >= 0 && n_descs <= MLI_ADAMW_MAX_TENSORS, "adamw batch: bad descriptors");
  MLI_REQUIRE(step >= 1, "adamw batch: step must be >= 1");
  AdamBatch b;
  uint64_t blocks = 0;
  b.n = 0;
  for (int i = 0; i < n_descs; ++i) {
    const mli_adamw_desc_t& d = descs_on_host[i];
    MLI_REQUIRE(d.n >= 0, "adamw batch: negative size");
    if (d.n == 0) continue;
    MLI_REQUIRE(d.param && d.grad && d.exp_avg && d.exp_avg_sq, "adamw batch: null buffer");
    b.d[b.n] = d;
    b.first_block[b.n] = (uint32_t)blocks;
    blocks += (uint64_t)((d.n + kAdamChunk - 1) / kAdamChunk);
    ++b.n;
  }
  if (b.n == 0) return MLI_OK;
  MLI_REQUIRE(blocks < (1ull << 31), "adamw batch: too many elements for one launch");
  b.first_block[b.n] = (uint32_t)blocks;
  const AdamHyper h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step, grad_scale);
  adamw_batch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(b, h);
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

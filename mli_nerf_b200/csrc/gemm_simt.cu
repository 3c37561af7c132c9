// gemm_simt.cu -- fp32 CUDA-core dense layers: the "fp32 validation mode" of the MLP path
// (rtol 1e-3 parity vs the oracle).  The fast path for the same entry points is gemm_tcgen05.cu (bf16 operands
// on the 5th-gen tensor cores); both sit behind mli_linear_{fwd,dgrad,wgrad}.
//
// Replaces torch.nn.Linear (+ activation) of /root/reference/projects/neuralangelo/utils/mlp.py:55-69 and
// /root/reference/projects/nerf/utils/nerf_util.py:186-196 and their autograd backward.
//
// NT kernel: C[M,N] = epi(A[M,K] . B[N,K]^T); 128x128x16 tiles, 256 threads, 8x8 register micro-tiles,
// double-buffered smem.  TN kernel (weight gradients): C[N,K] = sum_m dZ[m,N] X[m,K], split over M with a
// deterministic second-stage reduction (fixed summation order => reproducible gradients).
#include "common.cuh"

int mli_tc_linear_fwd(const float* X, int64_t ldx, int64_t sx, const float* W, int64_t ldw, int64_t sw,
                      const float* bias, int64_t sb, float* Y, int64_t ldy, int64_t sy, int64_t M, int32_t N,
                      int32_t K, int32_t act, int32_t batch, void* stream);

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

enum { EPI_BIAS_ACT = 0, EPI_MUL_DACT = 1 };

struct NTArgs {
  const float* A; int64_t lda, sa;
  const float* B; int64_t ldb, sb;
  float* C; int64_t ldc, sc;
  const float* bias; int64_t sbias;
  const float* aux; int64_t ldaux, saux;
  int64_t M;
  int N, K, act, accumulate;
};

template <int EPI>
__global__ void __launch_bounds__(NT, 2) gemm_nt_kernel(NTArgs p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int b = blockIdx.z;
  const float* A = p.A + b * p.sa;
  const float* B = p.B + b * p.sb;
  float* C = p.C + b * p.sc;

  const int lrow = tid & 127, kq = (tid >> 7) * 8;
  const bool a_ok = (m0 + lrow) < p.M, b_ok = (n0 + lrow) < p.N;
  const float4* a_src = reinterpret_cast<const float4*>(A + (m0 + lrow) * p.lda + kq);
  const float4* b_src = reinterpret_cast<const float4*>(B + (int64_t)(n0 + lrow) * p.ldb + kq);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float4 ra0, ra1, rb0, rb1;
  auto gload = [&](int kt) {
    ra0 = a_ok ? __ldg(a_src + kt * 4) : z4;
    ra1 = a_ok ? __ldg(a_src + kt * 4 + 1) : z4;
    rb0 = b_ok ? __ldg(b_src + kt * 4) : z4;
    rb1 = b_ok ? __ldg(b_src + kt * 4 + 1) : z4;
  };
  auto sstore = [&](int buf) {
    As[buf][kq + 0][lrow] = ra0.x; As[buf][kq + 1][lrow] = ra0.y; As[buf][kq + 2][lrow] = ra0.z; As[buf][kq + 3][lrow] = ra0.w;
    As[buf][kq + 4][lrow] = ra1.x; As[buf][kq + 5][lrow] = ra1.y; As[buf][kq + 6][lrow] = ra1.z; As[buf][kq + 7][lrow] = ra1.w;
    Bs[buf][kq + 0][lrow] = rb0.x; Bs[buf][kq + 1][lrow] = rb0.y; Bs[buf][kq + 2][lrow] = rb0.z; Bs[buf][kq + 3][lrow] = rb0.w;
    Bs[buf][kq + 4][lrow] = rb1.x; Bs[buf][kq + 5][lrow] = rb1.y; Bs[buf][kq + 6][lrow] = rb1.z; Bs[buf][kq + 7][lrow] = rb1.w;
  };

  const int KT = p.K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < KT) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  const float* bias = (EPI == EPI_BIAS_ACT && p.bias) ? p.bias + b * p.sbias : nullptr;
  const float* aux = (EPI == EPI_MUL_DACT && p.aux) ? p.aux + b * p.saux : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= p.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gn = n0 + h * 64 + tx * 4;
      if (gn >= p.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      if (EPI == EPI_BIAS_ACT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = mli_act(v[j] + (bias ? bias[gn + j] : 0.0f), p.act);
      } else {
        if (aux) {
          const float4 y = *reinterpret_cast<const float4*>(aux + gm * p.ldaux + gn);
          v[0] *= mli_dact_from_out(y.x, p.act); v[1] *= mli_dact_from_out(y.y, p.act);
          v[2] *= mli_dact_from_out(y.z, p.act); v[3] *= mli_dact_from_out(y.w, p.act);
        }
      }
      float4* dst = reinterpret_cast<float4*>(C + gm * p.ldc + gn);
      if (p.accumulate) {
        float4 o = *dst;
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      }
      *dst = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct TNArgs {
  const float* dZ; int64_t lddz, sdz;  // [M, N_out]
  const float* X; int64_t ldx, sx;     // [M, K_in]
  float* part;                         // [batch*S][N_out][K_in]
  float* part_db;                      // [batch*S][N_out]
  int64_t M, rows_per_split;
  int N_out, K_in, S;
};

__global__ void __launch_bounds__(NT, 2) gemm_tn_kernel(TNArgs p) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];  // dZ chunk: [m][n]
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];  // X chunk:  [m][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * BN, n0 = blockIdx.y * BM;
  const int b = blockIdx.z / p.S, s = blockIdx.z % p.S;
  const float* dZ = p.dZ + b * p.sdz;
  const float* X = p.X + b * p.sx;
  const int64_t m_begin = (int64_t)s * p.rows_per_split;
  const int64_t m_end = min(p.M, m_begin + p.rows_per_split);

  const int lm = tid >> 5, c4 = (tid & 31) * 4;  // rows lm and lm+8 of the 16-row chunk, 4 columns
  const bool n_ok = (n0 + c4) < p.N_out, k_ok = (k0 + c4) < p.K_in;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  float dbsum = 0.0f;

  float4 ra0, ra1, rb0, rb1;
  auto gload = [&](int64_t m) {
    const int64_t r0 = m + lm, r1 = m + lm + 8;
    ra0 = (n_ok && r0 < m_end) ? __ldg(reinterpret_cast<const float4*>(dZ + r0 * p.lddz + n0 + c4)) : z4;
    ra1 = (n_ok && r1 < m_end) ? __ldg(reinterpret_cast<const float4*>(dZ + r1 * p.lddz + n0 + c4)) : z4;
    rb0 = (k_ok && r0 < m_end) ? __ldg(reinterpret_cast<const float4*>(X + r0 * p.ldx + k0 + c4)) : z4;
    rb1 = (k_ok && r1 < m_end) ? __ldg(reinterpret_cast<const float4*>(X + r1 * p.ldx + k0 + c4)) : z4;
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lm][c4]) = ra0;
    *reinterpret_cast<float4*>(&As[buf][lm + 8][c4]) = ra1;
    *reinterpret_cast<float4*>(&Bs[buf][lm][c4]) = rb0;
    *reinterpret_cast<float4*>(&Bs[buf][lm + 8][c4]) = rb1;
  };

  const int64_t n_chunks = (m_end - m_begin + BK - 1) / BK;
  if (n_chunks > 0) {
    gload(m_begin);
    sstore(0);
  }
  __syncthreads();
  for (int64_t ct = 0; ct < n_chunks; ++ct) {
    const int buf = (int)(ct & 1);
    if (ct + 1 < n_chunks) gload(m_begin + (ct + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (blockIdx.x == 0 && tid < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) dbsum += As[buf][k][tid];
    }
    if (ct + 1 < n_chunks) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  float* part = p.part + (size_t)blockIdx.z * p.N_out * p.K_in;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gn = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gn >= p.N_out) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gk = k0 + h * 64 + tx * 4;
      if (gk >= p.K_in) continue;
      *reinterpret_cast<float4*>(part + (size_t)gn * p.K_in + gk) =
          make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
    }
  }
  if (blockIdx.x == 0 && tid < BM && n0 + tid < p.N_out && p.part_db)
    p.part_db[(size_t)blockIdx.z * p.N_out + n0 + tid] = dbsum;
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_db, int S,
                                    int N_out, int K_in, float* __restrict__ dW, int64_t lddw, int64_t sdw,
                                    float* __restrict__ db, int64_t sdb) {
  const int b = blockIdx.y;
  const int64_t total = (int64_t)N_out * K_in;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < total) {
    const int n = (int)(e / K_in), k = (int)(e % K_in);
    float v = 0.0f;
    for (int s = 0; s < S; ++s) v += part[((size_t)(b * S + s)) * total + e];  // fixed order: deterministic
    dW[b * sdw + (int64_t)n * lddw + k] = v;
  }
  if (db && e < N_out) {
    float v = 0.0f;
    for (int s = 0; s < S; ++s) v += part_db[(size_t)(b * S + s) * N_out + e];
    db[b * sdb + e] = v;
  }
}

int wgrad_splits(int64_t M, int N_out, int K_in, int batch) {
  int64_t tiles = (int64_t)mli_cdiv(K_in, BN) * mli_cdiv(N_out, BM) * batch;
  int64_t s = (2 * MLI_NUM_SMS + tiles - 1) / tiles;
  int64_t max_s = (M + 1023) / 1024;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace

extern "C" int mli_linear_fwd(const float* X, int64_t ldx, int64_t sx, const float* W, int64_t ldw, int64_t sw,
                              const float* bias, int64_t sb, float* Y, int64_t ldy, int64_t sy, int64_t M, int32_t N,
                              int32_t K, int32_t act, int32_t batch, int32_t prec, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(M >= 0 && N > 0 && K > 0 && batch >= 1, "bad GEMM shape");
  MLI_REQUIRE(K % BK == 0 && N % 4 == 0, "K must be a multiple of 16 and N of 4 (K=%d N=%d)", K, N);
  MLI_REQUIRE(ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0, "leading dimensions must be multiples of 4");
  MLI_REQUIRE(act >= MLI_ACT_NONE && act <= MLI_ACT_SIGMOID, "unknown activation %d", act);
  if (M == 0) return MLI_OK;
  if (prec == MLI_PREC_BF16)
    return mli_tc_linear_fwd(X, ldx, sx, W, ldw, sw, bias, sb, Y, ldy, sy, M, N, K, act, batch, stream);
  MLI_REQUIRE(prec == MLI_PREC_FP32, "unknown precision mode %d", prec);
  NTArgs p{X, ldx, sx, W, ldw, sw, Y, ldy, sy, bias, sb, nullptr, 0, 0, M, N, K, act, 0};
  dim3 g(mli_cdiv(N, BN), mli_cdiv(M, BM), batch);
  gemm_nt_kernel<EPI_BIAS_ACT><<<g, NT, 0, (cudaStream_t)stream>>>(p);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int mli_linear_dgrad(const float* dZ, int64_t lddz, int64_t sdz, const float* Wt, int64_t ldwt,
                                int64_t swt, const float* Yprev, int64_t ldyp, int64_t syp, float* dX, int64_t lddx,
                                int64_t sdx, int64_t M, int32_t N_out, int32_t K_in, int32_t act_prev,
                                int32_t accumulate, int32_t batch, int32_t prec, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(M >= 0 && N_out > 0 && K_in > 0 && batch >= 1, "bad GEMM shape");
  MLI_REQUIRE(N_out % BK == 0 && K_in % 4 == 0, "N_out must be a multiple of 16 and K_in of 4");
  MLI_REQUIRE(lddz % 4 == 0 && ldwt % 4 == 0 && lddx % 4 == 0 && (Yprev == nullptr || ldyp % 4 == 0),
              "leading dimensions must be multiples of 4");
  MLI_REQUIRE(prec == MLI_PREC_FP32, "dgrad: precision mode %d not available", prec);
  if (M == 0) return MLI_OK;
  // dX[M,K_in] = dZ[M,N_out] . Wt[K_in,N_out]^T  -> NT GEMM with contraction length N_out
  NTArgs p{dZ, lddz, sdz, Wt, ldwt, swt, dX, lddx, sdx, nullptr, 0, Yprev, ldyp, syp, M, K_in, N_out,
           Yprev ? act_prev : MLI_ACT_NONE, accumulate};
  dim3 g(mli_cdiv(K_in, BN), mli_cdiv(M, BM), batch);
  gemm_nt_kernel<EPI_MUL_DACT><<<g, NT, 0, (cudaStream_t)stream>>>(p);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

extern "C" int64_t mli_linear_wgrad_ws_bytes(int64_t M, int32_t N_out, int32_t K_in, int32_t batch) {
  int S = wgrad_splits(M, N_out, K_in, batch);
  return (int64_t)batch * S * ((int64_t)N_out * K_in + N_out) * sizeof(float);
}

extern "C" int mli_linear_wgrad(const float* dZ, int64_t lddz, int64_t sdz, const float* X, int64_t ldx, int64_t sx,
                                float* dW, int64_t lddw, int64_t sdw, float* db, int64_t sdb, int64_t M,
                                int32_t N_out, int32_t K_in, int32_t batch, int32_t prec, void* ws, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(M >= 1 && N_out > 0 && K_in > 0 && batch >= 1, "bad GEMM shape");
  MLI_REQUIRE(N_out % 4 == 0 && K_in % 4 == 0 && lddz % 4 == 0 && ldx % 4 == 0, "dims must be multiples of 4");
  MLI_REQUIRE(prec == MLI_PREC_FP32, "wgrad: precision mode %d not available", prec);
  MLI_REQUIRE(ws != nullptr, "wgrad workspace is NULL");
  const int S = wgrad_splits(M, N_out, K_in, batch);
  int64_t rows = (M + S - 1) / S;
  rows = (rows + BK - 1) / BK * BK;
  float* part = reinterpret_cast<float*>(ws);
  float* part_db = part + (size_t)batch * S * N_out * K_in;
  TNArgs p{dZ, lddz, sdz, X, ldx, sx, part, part_db, M, rows, N_out, K_in, S};
  dim3 g(mli_cdiv(K_in, BN), mli_cdiv(N_out, BM), S * batch);
  gemm_tn_kernel<<<g, NT, 0, (cudaStream_t)stream>>>(p);
  MLI_LAUNCH_OK();
  dim3 g2(mli_cdiv((int64_t)N_out * K_in, 256), batch);
  wgrad_reduce_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(part, part_db, S, N_out, K_in, dW, lddw, sdw, db, sdb);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

/*
 * mli_b200.h -- C ABI of libmli_b200.so: the B200 (sm_100a) implementation of MLI-NeRF's per-ray render
 * hot path (projects/NeuralLumen, reference citations relative to /root/reference/).
 *
 * The reference has no FFI of its own: its only native boundary is
 *   tinycudann.Encoding(3, {"otype":"HashGrid",...})      projects/neuralangelo/utils/modules.py:42-50,84-86
 * and everything else on the path is eager torch.  This header is the boundary a maintainer binds instead
 * (ctypes stub shown in INTEGRATION.md); each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types.  Every pointer is a DEVICE pointer unless named host_*.
 *   - the library never allocates or frees device memory and never takes ownership; the caller sizes
 *     workspaces with the *_bytes() helpers.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); stream-ordered, re-entrant
 *     per process, not thread-safe across host threads that share a device.
 *   - return 0 on success, negative MLI_E* on failure; mli_last_error() gives the thread-local message.
 *   - row-major everywhere.  "M" = number of sample points, "R" = number of rays, "N" = samples per ray.
 *   - there is NO CPU fallback: every entry point fails with MLI_ENODEV if no sm_100 device is current.
 */
#ifndef MLI_B200_H
#define MLI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLI_ABI_VERSION 4

enum {
  MLI_OK = 0,
  MLI_EINVAL = -1,   /* bad argument (shape, alignment, unknown mode) -> ValueError / NotImplementedError */
  MLI_ECUDA = -2,    /* CUDA runtime error (message has cudaGetErrorString) */
  MLI_ENODEV = -3,   /* no sm_100-class device current */
  MLI_ENOTSUP = -4   /* valid in the reference but not built (e.g. analytical gradient mode) */
};

const char* mli_last_error(void);
int mli_abi_version(void);
/* 1 if the current device is compute capability 10.x, else 0 (never errors) */
int mli_device_ok(void);
/* Number of SMs the persistent tensor-core kernels (one CTA per SM, all of its shared memory) may occupy, default 148.
 * A multi-GPU step lowers it so that the NCCL kernels of the overlapped gradient all-reduce find free SMs instead of
 * queueing behind every persistent CTA.  The workspace size helpers follow the current limit. */
int mli_set_sm_limit(int32_t n_sms);

/* ------------------------------------------------------------------------------------------------------
 * Hash grid (replaces tcnn.Encoding HashGrid; modules.py:42-50, 84-86)
 * ---------------------------------------------------------------------------------------------------- */
#define MLI_MAX_LEVELS 32

typedef struct {
  float scale;       /* exp2f(l*log2f(per_level_scale))*base_resolution - 1 */
  uint32_t res;      /* ceilf(scale)+1 */
  uint32_t size;     /* entries in this level */
  uint32_t offset;   /* first entry of this level in the flat table */
  uint32_t hashed;   /* 1: coherent-prime hash, 0: dense index */
} mli_level_t;

typedef struct {
  uint32_t n_levels;            /* L  (16) */
  uint32_t feat;                /* F  (8; must be 8, 4 or 2) */
  uint32_t active_levels;       /* coarse-to-fine mask (modules.py:110-113): levels >= this output 0 */
  uint32_t n_entries;           /* sum of level sizes */
  mli_level_t level[MLI_MAX_LEVELS];
} mli_grid_t;

/* host-only: fills `grid` the way tcnn's GridEncodingTemplated constructor does. */
int mli_grid_init(mli_grid_t* grid, uint32_t n_levels, uint32_t feat, uint32_t log2_hashmap_size,
                  uint32_t base_resolution, float per_level_scale);

/* tcnn.Encoding.forward: x01 [M,3] in [0,1] -> out [M, ld_out] (first L*F columns written, level-major). */
int mli_hashgrid_fwd(const mli_grid_t* grid, const float* table, const float* x01, int64_t M,
                     float* out, int64_t ld_out, void* stream);
/* tcnn.Encoding.backward w.r.t. params: table_grad[entry,f] += sum_m w * d_out[m, l*F+f]. */
int mli_hashgrid_bwd(const mli_grid_t* grid, const float* x01, int64_t M, const float* d_out, int64_t ld_dout,
                     float* table_grad, void* stream);
/* KAT helper: the 8 corner rows (with level offset) of every point for one level -> idx [M,8] uint32. */
int mli_hashgrid_corners(const mli_grid_t* grid, uint32_t level, const float* x01, int64_t M, uint32_t* idx,
                         void* stream);

/* Encode sample points along rays (fuses camera.get_3D_points_from_dist, camera.py:314-320; the [0,1]
 * normalisation + xyz concat of NeuralSDF.encode, modules.py:82-94; and the tap offsets of
 * compute_gradients, modules.py:133-166).  Writes rows [enc(L*F) | xyz(3) | 1 | 0-pad] of width ldx for
 * `planes` = 1 (centre only) or 1+taps stencil planes: row = plane*R*n + ray*n + i.
 *   dists [R, ld_d] (first n used), tap_eps = per-axis tap offset (eps/sqrt3 for 4 taps, eps for 6 taps). */
int mli_encode_rays(const mli_grid_t* grid, const float* table, const float* center, const float* ray_unit,
                    const float* dists, int64_t ld_d, int64_t R, int32_t n, int32_t taps, float tap_eps,
                    float vol_min, float vol_max, float* X, int64_t ldx, void* stream);
/* backward of mli_encode_rays w.r.t. the table: dX rows as above, enc columns only.
 * delta_basis != 0: dX is the gradient w.r.t. the delta-basis rows of mli_encode_rays_tcl (plane 0 = SUM over the
 * stencil planes of the per-plane gradients, plane i = tap i's own gradient). */
int mli_encode_rays_bwd(const mli_grid_t* grid, const float* center, const float* ray_unit, const float* dists,
                        int64_t ld_d, int64_t R, int32_t n, int32_t taps, float tap_eps, float vol_min,
                        float vol_max, const float* dX, int64_t ldx, float* table_grad, int32_t delta_basis,
                        void* stream);
/* Tensor-core variant of mli_encode_rays (bf16 MLP-tile mode): writes the SDF trunk's input as split-bf16 TCL in the
 * DELTA basis -- plane 0 rows = [enc | xyz | 0] of the centre point, plane i rows = row(tap i) - row(centre), the
 * difference formed in fp32 BEFORE the down-cast (the 4-tap stencil does not survive bf16
 * values, it does survive bf16 deltas).  X is TCL-128 with x_chunks >= 2*k_chunks chunks per tile row: chunks
 * [0, k_chunks) = bf16(x), chunks [k_chunks, 2 k_chunks) = bf16(x - bf16(x)); chunk l < L is level l (F = 8),
 * chunk L = [xyz | 0], the rest zero.  With taps, R*n must be a multiple of 128. */
int mli_encode_rays_tcl(const mli_grid_t* grid, const float* table, const float* center, const float* ray_unit,
                        const float* dists, int64_t ld_d, int64_t R, int32_t n, int32_t taps, float tap_eps,
                        float vol_min, float vol_max, void* X, int32_t x_chunks, int32_t k_chunks, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Dense layers (replace torch.nn.Linear + weight_norm + activation: mlp.py:55-69, nerf_util.py:186-196)
 * ---------------------------------------------------------------------------------------------------- */
enum { MLI_ACT_NONE = 0, MLI_ACT_RELU = 1, MLI_ACT_SOFTPLUS100 = 2, MLI_ACT_SIGMOID = 3 };
enum { MLI_PREC_FP32 = 0, MLI_PREC_BF16 = 1 };

/* Y[b] = act(X[b] W[b]^T + bias[b]) for b < batch.  X [M,K] ld ldx, W [N,K] ld ldw, Y [M,N] ld ldy.
 * strides are in elements between batch members (0 = shared).  K % 16 == 0, ld* % 4 == 0. */
int mli_linear_fwd(const float* X, int64_t ldx, int64_t sx, const float* W, int64_t ldw, int64_t sw,
                   const float* bias, int64_t sb, float* Y, int64_t ldy, int64_t sy, int64_t M, int32_t N,
                   int32_t K, int32_t act, int32_t batch, int32_t prec, void* stream);
/* dX[b] = (dZ[b] Wt[b]^T) * act'(Yprev[b])   -- Wt is W transposed: [K_in, N_out] ld ldwt.
 * act' is evaluated from the *output* Yprev of the previous layer (NULL / MLI_ACT_NONE: no factor).
 * accumulate != 0: dX += ... */
int mli_linear_dgrad(const float* dZ, int64_t lddz, int64_t sdz, const float* Wt, int64_t ldwt, int64_t swt,
                     const float* Yprev, int64_t ldyp, int64_t syp, float* dX, int64_t lddx, int64_t sdx,
                     int64_t M, int32_t N_out, int32_t K_in, int32_t act_prev, int32_t accumulate, int32_t batch,
                     int32_t prec, void* stream);
/* dW[b] = dZ[b]^T X[b]  ([N_out,K_in], ld lddw), db[b] = colsum(dZ[b]); split over M with a deterministic
 * second-stage reduction.  ws must hold mli_linear_wgrad_ws_bytes(...) bytes. */
int64_t mli_linear_wgrad_ws_bytes(int64_t M, int32_t N_out, int32_t K_in, int32_t batch);
int mli_linear_wgrad(const float* dZ, int64_t lddz, int64_t sdz, const float* X, int64_t ldx, int64_t sx,
                     float* dW, int64_t lddw, int64_t sdw, float* db, int64_t sdb, int64_t M, int32_t N_out,
                     int32_t K_in, int32_t batch, int32_t prec, void* ws, void* stream);

/* ---- bf16 tensor-core path (tcgen05 / TMEM / TMA bulk copies), see csrc/gemm_tcgen05.cu ---------------------------
 * Activations and their gradients live in HBM as bf16 in "TCL" (tile-chunk layout): a matrix X[M, C] (C % 8 == 0) is
 * stored as [ceil(M/128)][C/8][128][8]; element (m, c) -> (((m/128)*(C/8) + c/8)*128 + m%128)*8 + c%8.  Buffers are
 * padded to whole 128-row tiles; padding rows must be zero wherever the matrix is later contracted over rows.
 * A "chunk" is 8 columns.  Weights use the same layout with the tile height equal to the N tile (BN) of the GEMM. */

/* fp32 row-major [M, cols] (ld) -> chunks [chunk0, chunk0+n_chunks) of a bf16 TCL matrix with `dst_chunks` chunks per
 * tile row and `tile_rows`-row tiles; rows >= M / columns >= cols are zero-filled. */
int mli_tc_to_tcl(const float* src, int64_t ld, int64_t M, int32_t cols, void* dst, int32_t tile_rows,
                  int32_t dst_chunks, int32_t chunk0, int32_t n_chunks, void* stream);
/* backward of mli_encode_rays_tcl w.r.t. the table: dX = bf16 TCL-128 gradient of the delta-basis rows (x_chunks chunks
 * per tile row, chunk l = level l; plane 0 = sum over the stencil planes, plane i = tap i), as the data-gradient GEMM
 * of the SDF trunk writes it.  Only levels [level_begin, level_end) are processed, so a caller can launch level groups
 * separately and start the gradient all-reduce of a finished level's slab while the next ones are still running. */
int mli_encode_rays_bwd_tcl(const mli_grid_t* grid, const float* center, const float* ray_unit, const float* dists,
                            int64_t ld_d, int64_t R, int32_t n, int32_t taps, float tap_eps, float vol_min,
                            float vol_max, const void* dX, int32_t x_chunks, float* table_grad, int32_t level_begin,
                            int32_t level_end, void* stream);
/* split-bf16 variant: chunks [chunk0, +n_chunks) = bf16(x), chunks [lo_chunk0, +n_chunks) = bf16(x - bf16(x)). */
int mli_tc_to_tcl_split(const float* src, int64_t ld, int64_t M, int32_t cols, void* dst, int32_t tile_rows,
                        int32_t dst_chunks, int32_t chunk0, int32_t lo_chunk0, int32_t n_chunks, void* stream);
int mli_tc_from_tcl(const void* src, int32_t src_chunks, int32_t chunk0, int32_t n_chunks, int64_t M, float* dst,
                    int64_t ld, void* stream);
/* Forward / data-gradient GEMM on TCL operands, fused epilogue:
 *   epi 0: out = act(A W^T + bias)           epi 1: out = (A W^T) * act'(aux)   (aux = layer output, TCL bf16, or NULL)
 * A: TCL-128 activations, columns [8*(a_chunk0 + b*a_batch_chunks), +K) for batch member b.
 * B: weights [N, K] in TCL with BN-row tiles ([N/BN][K/8][BN][8]), b_batch_elems elements between batch members.
 * out: bf16 TCL-128 (out_chunks per tile row, first chunk out_chunk0 + b*out_batch_chunks) or, if out_is_f32, fp32
 * row-major (ldo; first column out_col0 + b*out_batch_cols; rows >= M are not written).
 * relu_mask (may be NULL; act must be relu): sign bits of the layer output, one uint32 per row and 32-column chunk,
 * [ceil(M/128)][mask_chunks][128], first chunk mask_chunk0 + b*mask_batch_chunks.  epi 0 WRITES it next to `out`;
 * epi 1 READS it in place of `aux` (16x fewer bytes than the bf16 output for the same relu'). */
int mli_tc_linear(const void* A, int32_t a_chunks, int32_t a_chunk0, int32_t a_batch_chunks, const void* B,
                  int64_t b_batch_elems, int32_t K, int32_t N, int32_t BN, const float* bias, int32_t bias_batch,
                  const void* aux, int32_t aux_chunks, int32_t aux_chunk0, int32_t aux_batch_chunks, int32_t act,
                  void* out, int32_t out_is_f32, int32_t out_chunks, int32_t out_chunk0, int32_t out_batch_chunks,
                  int64_t ldo, int32_t out_col0, int32_t out_batch_cols, int64_t M, int32_t batch, int32_t epi,
                  void* relu_mask, int32_t mask_chunks, int32_t mask_chunk0, int32_t mask_batch_chunks, void* stream);
/* relu hidden layer (N = 256 per batch member, as mli_tc_linear epi 0) with the narrow output layer that follows it
 * (nerf_util.py:191, 256 -> 3/3/1) fused into the epilogue: S[m, j] = act_j(w_out[j] . out[m, batch(j) block] + b_out[j])
 * for j in [host_j0[b], host_j0[b] + host_nj[b]) of batch member b (nj <= 4, batch <= 4); the dot uses the fp32 layer
 * output before its bf16 rounding.  Saves one HBM pass over the [M, 256*batch] activations. */
int mli_tc_linear_dot(const void* A, int32_t a_chunks, int32_t a_chunk0, int32_t a_batch_chunks, const void* B,
                      int64_t b_batch_elems, int32_t K, const float* bias, int32_t bias_batch, void* out,
                      int32_t out_chunks, int32_t out_chunk0, int32_t out_batch_chunks, int64_t M, int32_t batch,
                      const float* w_out, const float* b_out, const int32_t* host_j0, const int32_t* host_nj,
                      int32_t act_out, uint32_t act_mask, float* S, int64_t lds, void* relu_mask, int32_t mask_chunks,
                      int32_t mask_chunk0, int32_t mask_batch_chunks, void* stream);

/* SDF trunk on the tensor cores (replaces MLPforNeuralSDF layer 0 + softplus + linear_sdf, mlp.py:55-69, for every
 * stencil plane; the numerical-gradient taps of modules.py:131-177 go through as fp32-formed deltas).
 *   X      split-bf16 TCL rows from mli_encode_rays_tcl (x_chunks chunks per tile row, K = 8*k_chunks padded inputs)
 *   W0s    layer-0 weights [256, K] as split-bf16 TCL with 256-row tiles (mli_tc_to_tcl_split), b0 [256]
 *   three tcgen05 products per k-step (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) ~ 16 mantissa bits
 * mode 0 (centre rows):  z0 = x0 W0^T + b0;  sigma0 = sigmoid(100 z0) -> fp32 "TCL32" [rows/128][64][128][4];
 *                        h0 = softplus100(z0) -> bf16 TCL (32 chunks);  vec_out[row] = w_sdf . h0 + b_sdf
 * mode 1 (tap rows, `rows` = taps * rows_per_plane):  dz = dx W0^T;  dh = log1p(expm1(100 dz) sigma0) / 100
 *                        (= softplus(z0 + dz) - softplus(z0));  vec_out[row] = w_sdf . dh  (= sdf_tap - sdf_centre);
 *                        dz -> bf16 TCL at h_or_dz unless NULL (only the backward pass needs it)
 * mode 2 (SDF only, e.g. sampling queries): vec_out[row] = w_sdf . softplus100(x W0^T + b0) + b_sdf */
int mli_tc_sdf_trunk_fwd(const void* X, int32_t x_chunks, int32_t K, const void* W0s, const float* b0,
                         const float* w_sdf, const float* b_sdf, int64_t rows, int32_t mode, int64_t rows_per_plane,
                         float* sigma0, void* h_or_dz, float* vec_out, void* stream);
/* The same in ONE launch for a training/eval forward: per 128-sample tile the centre rows and every tap plane run in
 * the same persistent CTA, sigma0 stays in TMEM between them (never re-read from HBM).  X holds the 1+taps planes
 * (plane-major, M rows each); sdf [(1+taps)*M] receives plane 0 = sdf, plane i = sdf_tap_i - sdf_centre; sigma0 and dz
 * may be NULL when no backward pass follows.  M % 128 == 0. */
int mli_tc_sdf_trunk_fused(const void* X, int32_t x_chunks, int32_t K, const void* W0s, const float* b0,
                           const float* w_sdf, const float* b_sdf, int64_t M, int32_t taps, float* sigma0, void* h0,
                           void* dz, float* sdf, void* stream);
/* Backward of the trunk's activation / SDF head in the delta basis.  g [(1+taps)*M] = dL/d sdf of every stencil plane
 * (mli_geometry_bwd), dH0 = dL/d h0 arriving through layer 1 (bf16 TCL, 32 chunks, may be NULL).  Writes Ed (bf16 TCL,
 * (1+taps)*M rows x 32 chunks): plane 0 = E = sum over planes of e_p, plane i = e_i, with
 *   e_0 = (g_0 w_sdf + dH0) sigma0,   e_i = g_i w_sdf sigmoid(100 (z0 + dz_i))
 * so that dW0 = Ed^T Xd, d Xd = Ed W0, db0 = colsum(E) in the basis of mli_encode_rays_tcl.
 * dw_sdf [256] / db_sdf [1] may be NULL (heads-only training). ws: mli_tc_sdf_trunk_bwd_ws_bytes(). M % 128 == 0. */
int64_t mli_tc_sdf_trunk_bwd_ws_bytes(int64_t M);
int mli_tc_sdf_trunk_bwd(const float* g, int64_t M, int32_t taps, const float* sigma0, const void* dz, const void* dH0,
                         const void* h0, const float* w_sdf, void* Ed, float* dw_sdf, float* db_sdf, void* ws,
                         void* stream);

/* Fused forward of the WHOLE head stack (LumenRGB.forward, NeuralLumen/utils/modules.py:106-174: per head the five
 * layers of MLPwithSkipConnection.forward, nerf_util.py:186-196) in one persistent tcgen05 kernel (csrc/heads_fused.cu):
 * the [256 x 256] hidden activations of a tile pair stay in shared memory between layers, weights stream from L2.
 * XH: TCL-128 input of head layer 0 (xh_chunks chunks per tile row, the first K0/8 are used); W0: layer-0 weights of
 * all nh heads in TCL with 256-row tiles, [nh][K0/8][256][8]; W1..W3: hidden-layer weights in TCL with 128-row tiles,
 * [nh][2][32][128][8] (mli_weightnorm_pack_batch, tcl / tcl2); bias0..3 [nh*256]; w_out [J,256] / b_out [J] fp32, head h owning outputs [host_j0[h], +host_nj[h]) (<= 4
 * each), act_out applied where bit j of act_mask is set.  store_activations != 0 (a backward pass follows): A0..A3
 * (bf16 TCL-128, nh*32 chunks per tile row) and the relu sign bits mask0..3 ([tiles][nh*8][128] uint32, may be NULL)
 * are written; otherwise nothing but S [M, lds] leaves the SM.  M % 128 == 0. */
int mli_tc_heads_fwd(const void* XH, int64_t M, int32_t nh, int32_t K0, int32_t store_activations, int32_t xh_chunks,
                     const void* W0, const void* W1, const void* W2, const void* W3, const float* bias0,
                     const float* bias1, const float* bias2, const float* bias3, const float* w_out, const float* b_out,
                     const int32_t* host_j0, const int32_t* host_nj, int32_t act_out, uint32_t act_mask, void* A0,
                     void* A1, void* A2, void* A3, void* mask0, void* mask1, void* mask2, void* mask3, float* S,
                     int64_t lds, void* stream);

/* Fused data-gradient chain of the head stack (backward of the same layers, csrc/heads_fused.cu): from dS [M, lds]
 * (gradient w.r.t. the PRE-activation head outputs, mli_composite_bwd) and the relu sign bits mask0..3 written by
 * mli_tc_heads_fwd to the pre-activation gradients of the four hidden layers,
 *     dZ3 = (dS W_out) * relu'(A3),   dZ_{l-1} = (dZ_l W_l) * relu'(A_{l-1})   (l = 3, 2, 1),
 * each written ONCE as bf16 TCL-128 (nh*32 chunks per tile row) -- the operands of the weight-gradient GEMMs and of the
 * layer-0 data gradient.  W3t, W2t, W1t: TRANSPOSED hidden-layer weights, TCL with 128-row tiles [nh][2][32][128][8]
 * (rows = input unit; mli_weightnorm_pack_batch, tclt).  M % 128 == 0. */
int mli_tc_heads_bwd(const float* dS, int64_t lds, int64_t M, int32_t nh, const void* W3t, const void* W2t,
                     const void* W1t, const float* w_out, const int32_t* host_j0, const int32_t* host_nj,
                     const void* mask0, const void* mask1, const void* mask2, const void* mask3, void* dZ0, void* dZ1,
                     void* dZ2, void* dZ3, void* stream);

/* Weight-gradient GEMM: out[b][r, c] = sum_m L[m, 8*(l_chunk0 + b*l_batch_chunks) + r] * R[m, 8*(r_chunk0 +
 * b*r_batch_chunks) + c]; r < rows_out (multiple of 128), c < cols_out (multiple of 16; < 256 or a multiple of 256).
 * L, R: TCL-128 (the same bytes serve as MN-major operands).  Split over row tiles, deterministic reduction.
 * transpose_out: write out[c, r] instead.  colsum_L (may be NULL): also colsum_L[b*colsum_batch_stride + r] =
 * sum_m L[m, ... + r] (the bias gradient when L = dZ), computed from the shared-memory stages at no extra HBM traffic. */
int64_t mli_tc_wgrad_ws_bytes(int64_t M, int32_t rows_out, int32_t cols_out, int32_t batch);
int mli_tc_wgrad(const void* L, int32_t l_chunks, int32_t l_chunk0, int32_t l_batch_chunks, const void* R,
                 int32_t r_chunks, int32_t r_chunk0, int32_t r_batch_chunks, int64_t M, int32_t rows_out,
                 int32_t cols_out, int32_t batch, float* out, int64_t ldo, int64_t out_batch_stride,
                 int32_t transpose_out, float* colsum_L, int64_t colsum_batch_stride, void* ws, void* stream);

/* The same GEMM with its split-K reduction DEFERRED: only the tensor-core kernel is launched and *job_on_host is filled
 * with what is left to do; mli_tc_wgrad_reduce_batch then reduces the partials of up to any number of such jobs (weight
 * and bias gradients of a whole backward pass: 13 small launches in the reference-shaped step) in ONE launch on the same
 * stream.  The workspaces must stay alive until that launch has run.  Results are bit-identical to mli_tc_wgrad. */
typedef struct {
  const float* part;      /* [batch][S][rows*cols] fp32 partials (inside ws) */
  const float* cs_part;   /* [batch][S][rows] column-sum partials, NULL without colsum_L */
  float* out;
  float* colsum;
  int64_t ldo, out_batch_stride, colsum_batch_stride;
  int32_t S, rows, cols, batch, transpose, reserved;
} mli_tn_reduce_job_t;
int mli_tc_wgrad_defer(const void* L, int32_t l_chunks, int32_t l_chunk0, int32_t l_batch_chunks, const void* R,
                       int32_t r_chunks, int32_t r_chunk0, int32_t r_batch_chunks, int64_t M, int32_t rows_out,
                       int32_t cols_out, int32_t batch, float* out, int64_t ldo, int64_t out_batch_stride,
                       int32_t transpose_out, float* colsum_L, int64_t colsum_batch_stride, void* ws,
                       mli_tn_reduce_job_t* job_on_host, void* stream);
int mli_tc_wgrad_reduce_batch(const mli_tn_reduce_job_t* jobs_on_host, int32_t n_jobs, void* stream);
/* Column sums (bias gradients) of chunks [chunk0, chunk0+n_chunks) of a TCL-128 matrix -> out[8*n_chunks]. */
int64_t mli_tc_colsum_ws_bytes(int64_t M, int32_t n_chunks);
int mli_tc_colsum(const void* src, int32_t src_chunks, int32_t chunk0, int32_t n_chunks, int64_t M, float* out, void* ws,
                  void* stream);

/* Narrow output layers on TCL activations (same maths as mli_rowdot_*): A is TCL-128 with a_chunks chunks per tile row;
 * col_off (host array, multiples of 8, non-decreasing) selects the K-wide column window of each output.
 * bwd_data writes dZ = (dS W) * act_prev'(A) as bf16 TCL with the same geometry as A (all a_chunks chunks); with
 * act_prev = relu and relu_mask != NULL ([tiles][a_chunks/4][128] uint32 sign bits) A itself is not read. */
int mli_tc_rowdot_fwd(const void* A, int32_t a_chunks, int64_t M, const float* w, const float* b,
                      const int32_t* host_col_off, int32_t J, int32_t K, int32_t act, uint32_t act_mask, float* out,
                      int64_t ldo, void* stream);
int mli_tc_rowdot_bwd_data(const float* dS, int64_t lds, const void* A, int32_t a_chunks, int64_t M, const float* w,
                           const int32_t* host_col_off, int32_t J, int32_t K, int32_t act_prev, void* dZ,
                           const void* relu_mask, void* stream);

/* Narrow output layers (SDF head 256->1, mlp.py:50,66; head output layers 256->3/3/1, nerf_util.py:191):
 * out[m, j] = act_j(sum_k A[m, col_off[j] + k] * w[j, k] + b[j]),  j < J <= 8, k < K; act_j = act if bit j of
 * act_mask is set, identity otherwise (network_mode r_s leaves o_s without a sigmoid, modules.py:119).
 * col_off is a HOST array of J column offsets. */
int mli_rowdot_fwd(const float* A, int64_t lda, int64_t M, const float* w, const float* b,
                   const int32_t* host_col_off, int32_t J, int32_t K, int32_t act, uint32_t act_mask, float* out,
                   int64_t ldo, void* stream);
/* dA[m,c] = ((accumulate ? dA[m,c] : 0) + sum_{j: col_off[j] <= c < col_off[j]+K} dS[m,j] w[j,c-col_off[j]])
 *           * act_prev'(A[m,c])   for c < n_cols_dA;   dw[j,k] = sum_m dS[m,j] A[m,col_off[j]+k];  db[j] = sum_m dS[m,j].
 * dA and/or dw may be NULL to skip that half. */
int64_t mli_rowdot_bwd_ws_bytes(int64_t M, int32_t J, int32_t K);
int mli_rowdot_bwd(const float* dS, int64_t lds, const float* A, int64_t lda, int64_t M, const float* w,
                   const int32_t* host_col_off, int32_t J, int32_t K, int32_t act_prev, float* dA, int64_t ldda,
                   int32_t n_cols_dA, int32_t accumulate, float* dw, float* db, void* ws, void* stream);

/* weight_norm reparameterisation W = g * v / ||v||_row (mlp.py:42-44), scattered into a packed/padded layout:
 * Wp[row_off + n, col_map[k]] = W[n,k]; Wpt is the transpose copy used by dgrad.  col_map NULL = identity. */
int mli_weightnorm_pack(const float* v, const float* g, int32_t N, int32_t K, const int32_t* col_map,
                        float* Wp, int64_t ldw, float* Wpt, int64_t ldwt, int32_t row_off, void* stream);
/* dv, dg from dWp (same layout as Wp). */
int mli_weightnorm_unpack_grad(const float* v, const float* g, const float* dWp, int64_t ldw, int32_t N, int32_t K,
                               const int32_t* col_map, int32_t row_off, float* dv, float* dg, void* stream);

/* All weight_norm reparameterisations of a step in ONE launch, written straight into the layouts the tensor-core
 * kernels read.  Per matrix (descriptor): W = g v/||v|| [N,K] is scattered (col_map, row_off as above) into any of
 *   Wp    fp32 row-major (ldw)                                                       (may be NULL)
 *   tcl   bf16 TCL with tcl_tile-row tiles and tcl_chunks chunks per tile row; tcl_lo >= 0 also writes the split-bf16
 *         remainder bf16(w - bf16(w)) tcl_lo chunks further                           (may be NULL)
 *   tcl2  a second bf16 TCL copy (tcl2_tile-row tiles, tcl2_chunks chunks per tile row)  (may be NULL)
 *   tclt  up to two bf16 TCL copies of column ranges [c0, c1) of the TRANSPOSE: element
 *         (tclt_row_off + c - c0, tclt_col_off + n)                                   (may be NULL)
 * Target buffers must be zero where no element is scattered (K padding).  The same descriptors drive the backward:
 * dWp (fp32, layout of Wp) -> dv [N,K], dg [N] (mli_weightnorm_unpack_grad_batch).  Descriptors live in HOST memory
 * (at most MLI_WN_MAX_DESCS; they travel as kernel parameters). */
#define MLI_WN_MAX_DESCS 24
typedef struct {
  const float* v; const float* g; const int32_t* col_map;
  float* Wp; void* tcl; void* tclt[2];
  const float* dWp; float* dv; float* dg;
  int64_t ldw;
  int32_t N, K, row_off;
  int32_t tcl_tile, tcl_chunks, tcl_lo;
  int32_t tclt_c0[2], tclt_c1[2], tclt_tile[2], tclt_chunks[2], tclt_row_off[2], tclt_col_off[2];
  int32_t row_begin;  /* filled by the library */
  void* tcl2; int32_t tcl2_tile, tcl2_chunks;  /* a second bf16 TCL copy with its own tile height (fused head kernel) */
} mli_wn_desc_t;
int mli_weightnorm_pack_batch(const mli_wn_desc_t* descs_on_host, int32_t n_descs, void* stream);
int mli_weightnorm_unpack_grad_batch(const mli_wn_desc_t* descs_on_host, int32_t n_descs, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Rays, bounds, sampling (projects/neuralangelo/model.py:420-484; nerf_util.py:20-68,199-205;
 * NeuralLumen/utils/utils.py:86-123)
 * ---------------------------------------------------------------------------------------------------- */
/* camera.get_center_and_ray + slice_by_ray_idx + F.normalize + get_center(pose_light)
 * (camera.py:283-311; nerf_util.py:127-131; NeuralLumen/model.py:120-131). pose/pose_light [B,3,4] = world->camera,
 * intr [B,3,3], ray_idx [B,R] pixel indices (NULL = 0..R-1), W = image width; outputs [B*R,3] (ray_norm [B*R]). */
int mli_rays_from_pose(const float* pose, const float* intr, const float* pose_light, const int64_t* ray_idx,
                       int64_t B, int64_t R, int32_t W, float* center, float* ray_unit, float* ray_norm,
                       float* pts_light, void* stream);
/* get_dist_bounds: aabb == NULL -> unit sphere.  outside is uint8 [R]. */
int mli_dist_bounds(const float* center, const float* ray_unit, int64_t R, const float* host_aabb6, float* near,
                    float* far, uint8_t* outside, void* stream);
/* nerf_util.sample_dists: dists[r,i] = (rand+i)/n*(far-near)+near; rands NULL -> 0.5. */
int mli_sample_coarse(const float* near, const float* far, const float* rands, int64_t R, int32_t n, float* dists,
                      int64_t ld_d, void* stream);
/* sample_dists_hierarchical + sample_dists_from_pdf: dists/sdfs [R, ld] (first n used) -> fine [R, n_fine].
 * Optional KAT outputs (may be NULL): idx/low/high int32 [R, n_fine], cdf float [R, n] (n = n-1 weights + leading 0). */
int mli_sample_fine(const float* dists, const float* sdfs, int64_t ld, int64_t R, int32_t n, int32_t n_fine,
                    float inv_s, float* fine, int32_t* idx, int32_t* low, int32_t* high, float* cdf, void* stream);
/* inverse-CDF binning alone, from given weights [R, ld_w] (n_w used): bit-exact KAT entry. */
int mli_pdf_bins(const float* weights, int64_t ld_w, int64_t R, int32_t n_w, int32_t n_fine, int32_t* idx,
                 int32_t* low, int32_t* high, float* cdf, void* stream);
/* mli_sample_merge(fine_in, sdf_fine) immediately followed by mli_sample_fine(inv_s_next) on the merged n + n_fine
 * samples, in one launch (the hierarchy loop of sample_dists_all, neuralangelo/model.py:455-465): dists/sdfs are merged
 * in place, fine_out [R, n_fine] receives the next round's samples.  Same results, bit for bit, as the two calls. */
int mli_sample_merge_fine(float* dists, float* sdfs, int64_t ld, int64_t R, int32_t n, const float* fine_in,
                          const float* sdf_fine, int32_t n_fine, float inv_s_next, float* fine_out, void* stream);
/* cat + sort(dim=2) (+ gather of sdfs): merges n sorted + n_fine new samples, stable. sdf pointers may be NULL. */
int mli_sample_merge(float* dists, float* sdfs, int64_t ld, int64_t R, int32_t n, const float* fine,
                     const float* sdf_fine, int32_t n_fine, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Per-sample geometry + head inputs (modules.py:157-175; NeuralLumen/model.py:343-349;
 * spherical_harmonics.py:47-70; NeuralLumen/utils/modules.py:106-109)
 * ---------------------------------------------------------------------------------------------------- */
/* sdf [planes*M] (plane 0 = centre; overwritten with outside_val for outside rays, model.py:343).
 * Writes gradients/hessians [M,3] (hessians NULL in eval) and XH[m, xh_off .. xh_off+38] =
 * [pts(3) | SH(view)(16) | normal(3) | SH(light position)(16)]; columns beyond that up to ldxh are zeroed.
 * tap_eps is the per-axis offset as a double (eps/sqrt(3) for 4 taps); the float32 divisors 4e, e^2 are derived
 * from it exactly as torch derives them from the Python float. */
/* sdf_is_delta != 0: planes >= 1 of `sdf` hold sdf_tap - sdf_centre (mli_tc_sdf_trunk_fwd mode 1) instead of
 * absolute values; they are left as deltas.  xh_tcl (may be NULL): the same 38 (+10 zero) columns also / instead as bf16
 * TCL chunks [xh_chunk0, xh_chunk0+6) of a matrix with xh_chunks chunks per tile row (the tensor-core head input). */
int mli_geometry_fwd(float* sdf, int64_t M, int32_t N, int32_t taps, double tap_eps, const uint8_t* outside,
                     float outside_val, const float* center, const float* ray_unit, const float* pts_light,
                     const float* dists, int64_t ld_d, float* gradients, float* hessians, float* XH, int64_t ldxh,
                     int32_t xh_off, int32_t sdf_is_delta, void* xh_tcl, int32_t xh_chunks, int32_t xh_chunk0,
                     void* stream);
/* d_sdf [planes*M] = backward of gradients/hessians/normals (+ d_sdf_center_in from the alpha path).
 * d_grad_in / d_hess_in [M,3] may be NULL; dXH supplies d normal at columns xh_off+19..21 (may be NULL). */
int mli_geometry_bwd(const float* gradients, int64_t M, int32_t N, int32_t taps, double tap_eps,
                     const uint8_t* outside, const float* d_grad_in, const float* d_hess_in, const float* dXH,
                     int64_t ldxh, int32_t xh_off, const float* d_sdf_center_in, float* d_sdf, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * NeuS alpha + alpha compositing (neuralangelo/model.py:492-515; render.py:87-112;
 * NeuralLumen/model.py:266-323, eval extras :365-368, :101-104)
 * ---------------------------------------------------------------------------------------------------- */
/* cfg.model.object.rgb.network_mode (NeuralLumen/utils/modules.py:16-55).  Per-sample channels S / per-ray out:
 *   RGB     S = rgb3                 out = rgb3
 *   RGB_R_S S = rgb3|o_r3|o_s1       out = rgb3|o_r3|o_s1|o_re3     (o_re = rgb - o_r*o_s)
 *   RGB_R   S = rgb3|o_r3            out = rgb3|o_r3|o_s3           (o_s = rgb / o_r)
 *   R_S     S = o_r3|o_s3            out = rgb3|o_r3|o_s3           (rgb = o_r*o_s)
 *   R_S_RE  S = o_r3|o_s3|o_re3      out = rgb3|o_r3|o_s3|o_re3     (rgb = o_r*o_s + o_re) */
enum { MLI_MODE_RGB = 0, MLI_MODE_RGB_R_S = 1, MLI_MODE_RGB_R = 2, MLI_MODE_R_S = 3, MLI_MODE_R_S_RE = 4 };

typedef struct {
  int32_t N;              /* samples per ray (<= 256) */
  int32_t mode;           /* MLI_MODE_* */
  int32_t white_bg;       /* add (1-opacity) to every composited channel */
  int32_t eval_extras;    /* also opacity, gradient, dist composites */
  float anneal_ratio;     /* min(progress/anneal_end, 1) */
  const float* anneal_dev; /* optional DEVICE float: when non-NULL the kernels read the anneal ratio from it instead of
                            * anneal_ratio, so a captured CUDA graph follows the schedule (neuralangelo/model.py:492-499
                            * recomputes it from `progress` every iteration) without being re-captured */
} mli_composite_cfg_t;

/* S [M, lds] per-sample head outputs; out [R, n_out(mode)]; weights [R,N]; alphas [R,N] or NULL;
 * extras [R,5] = opacity, gradient(3), dist (only with eval_extras). */
int mli_composite_fwd(const mli_composite_cfg_t* cfg, const float* s_var, const float* sdf_center,
                      const float* gradients, const float* ray_unit, const float* dists, int64_t ld_d,
                      const float* far, const float* S, int64_t lds, int64_t R, float* alphas, float* weights,
                      float* out, float* extras, void* stream);
/* d_out [R,n_out], d_weights [R,N] or NULL -> dS_pre [M,lds] (gradient w.r.t. the PRE-activation head outputs;
 * act_mask bit c set = channel c went through a sigmoid), d_sdf_center [M] (written), d_gradients [M,3]
 * (ACCUMULATED into: pre-load it with the eikonal seeds or zeros), d_s_var[0] (+)= sum over rays (deterministic
 * two-stage reduction; ws = R floats; NULL d_s_var skips it). */
int mli_composite_bwd(const mli_composite_cfg_t* cfg, const float* s_var, const float* sdf_center,
                      const float* gradients, const float* ray_unit, const float* dists, int64_t ld_d,
                      const float* far, const float* S, int64_t lds, int64_t R, const float* weights,
                      const float* d_out, const float* d_weights, uint32_t act_mask, float* dS_pre,
                      float* d_sdf_center, float* d_gradients, float* d_s_var, int32_t accumulate_s_var, void* ws,
                      void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Losses, reduced in-kernel (NeuralLumen/trainer.py:133-149; neuralangelo/utils/misc.py:74-89;
 * NeuralLumen/utils/utils.py:142-174; imaginaire/trainers/base.py:534-544)
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  float w_render, w_eikonal, w_curvature, w_intrinsic, w_regularize_re;
  float range_sha[2], range_vis[2], factor_ref, factor_sha;
  float factor_negative, factor_positive, exponent_positive;
  int32_t has_intrinsic;  /* o_r/o_s/o_re + pseudo labels present */
  const float* weights_dev; /* optional DEVICE float[5] = w_render, w_eikonal, w_curvature, w_intrinsic, w_regularize_re:
                             * when non-NULL it replaces the five by-value weights (the curvature weight is re-scheduled
                             * every iteration during warm-up, neuralangelo/trainer.py:56-63) */
} mli_loss_cfg_t;

enum { MLI_LOSS_TOTAL = 0, MLI_LOSS_RENDER, MLI_LOSS_EIKONAL, MLI_LOSS_CURVATURE, MLI_LOSS_INTRINSIC,
       MLI_LOSS_REG_RE, MLI_LOSS_MSE, MLI_LOSS_COUNT = 8 };

/* out [R,n_out(mode)] as produced by mli_composite_fwd; pseudo_* may be NULL when !has_intrinsic.
 * losses [MLI_LOSS_COUNT] floats.  Seeds (already multiplied by the loss weights): d_out [R,n_out] in the same
 * column layout as out, d_gradients [M,3], d_hessians [M,3] (hessians NULL = eval: no curvature term).
 * ws: mli_losses_ws_bytes(). */
int64_t mli_losses_ws_bytes(int64_t R, int64_t M);
int mli_losses_fwd_bwd(const mli_loss_cfg_t* cfg, int32_t mode, const float* out, const float* gradients,
                       const float* hessians, const uint8_t* outside, int64_t R, int32_t N, const float* image,
                       const float* pseudo_ref, const float* pseudo_sha, const float* pseudo_vis, float* losses,
                       float* d_out, float* d_gradients, float* d_hessians, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * "Next" rows
 * ---------------------------------------------------------------------------------------------------- */
/* Light visibility by sphere tracing (NeuralLumen/model.py:133-200; neuralangelo/model.py:298-325), the stage-a export.
 * The SDF values of each marching iteration come from the sampling-path kernels (mli_encode_rays[_tcl] with one
 * sample per ray at distance dist[r] + the SDF trunk); these entry points are the per-ray state around them.
 *   step:   dist[mask] += sdf[mask]; mask[dist > far] = 0; mask[dist < near] = 0; last != 0: dist = clamp(dist, near, far)
 *   rays:   intersection point center + ray_unit*inter_dist -> unit light ray from pts_light, its marching interval
 *           [near_light, far_tracing = |light ray| - 1e-3] inside the visibility bound (sphere of `radius`, or the
 *           box host_aabb6 when not NULL) and inside_bounding = near < far_tracing < far & ~outside
 *   finish: visibility = ~mask_light | ~inside_bounding; normal_x_light = relu(normalize(-gradient) . light_unit) */
int mli_sphere_trace_step(float* dist, uint8_t* mask, const float* sdf, const float* near, const float* far, int64_t R,
                          int32_t last, void* stream);
int mli_light_rays(const float* center, const float* ray_unit, const float* inter_dist, const float* pts_light, int64_t R,
                   float radius, const float* host_aabb6, float* light_unit, float* near_light, float* far_tracing,
                   uint8_t* inside_bounding, void* stream);
int mli_light_finish(const uint8_t* mask_light, const uint8_t* inside_bounding, const float* gradient,
                     const float* light_unit, int64_t R, uint8_t* visibility, float* normal_x_light, void* stream);

/* dense AdamW step (torch.optim.AdamW semantics, get_trainer.py:106-150) fused with gradient scaling (1/world).
 * Hyper-parameters arrive as doubles: 1 - beta, 1 - lr*wd and the bias corrections are formed in double and rounded once,
 * as torch forms them from the Python floats. */
int mli_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int32_t step, double grad_scale,
                   void* stream);

/* The same update for every tensor of an optimizer group in ONE launch (descriptors in HOST memory, at most
 * MLI_ADAMW_MAX_TENSORS; they travel as kernel parameters).  Tensors of any size / alignment. */
#define MLI_ADAMW_MAX_TENSORS 64
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64_t n;
} mli_adamw_desc_t;
int mli_adamw_step_batch(const mli_adamw_desc_t* descs_on_host, int32_t n_descs, double lr, double beta1,
                         double beta2, double eps, double weight_decay, int32_t step, double grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Multi-GPU exchange of the hash-table gradient over NVLink peer memory (mli_nerf_b200/dist.py): replaces the DDP
 * all-reduce of this gradient (imaginaire/trainers/utils/get_trainer.py:80-88).  Every rank owns 1/W of a slab: it pulls
 * that shard from every peer's buffer with the copy engines (mli_copy_async on IPC-mapped peer pointers), reduces
 * (mli_reduce_slots: dst = (dst + sum of the n_slots staged shards) * scale) and pushes the mean back.
 * ---------------------------------------------------------------------------------------------------- */
int mli_enable_peer_access(int32_t peer_device);
/* gradient buffer other ranks can map: cudaMalloc + its 64-byte CUDA IPC handle / map a peer's buffer into the current
 * device's address space (cudaIpcOpenMemHandle with lazy peer access) / unmap / free */
int mli_peer_alloc(int64_t bytes, void** host_out_ptr, void* host_out_handle64);
int mli_peer_open(const void* host_handle64, void** host_out_ptr);
int mli_peer_close(void* mapped_ptr);
int mli_peer_free(void* ptr);
int mli_copy_async(void* dst, const void* src, int64_t bytes, void* stream);
int mli_reduce_slots(float* dst, const float* slots, int32_t n_slots, int64_t slot_stride, int64_t n, float scale,
                     void* stream);

/* ------------------------------------------------------------------------------------------------------
 * L2 residency of the dense hash-grid levels (tcnn keeps no such control; north star (a)): access-policy window on
 * `stream` over [ptr, ptr+bytes) -- persisting with probability hit_ratio, everything else streaming; bytes == 0
 * removes it.  mli_l2_info reports the device's L2 size, maximum persisting carve-out and maximum window size.
 * ---------------------------------------------------------------------------------------------------- */
int mli_set_l2_window(const void* ptr, int64_t bytes, float hit_ratio, void* stream);
int mli_l2_info(int32_t* host_out_l2_bytes, int32_t* host_out_max_persist_bytes, int32_t* host_out_max_window_bytes);

/* ------------------------------------------------------------------------------------------------------
 * Background zero-fill of the hash-table gradient buffer (the reference gets its zeroed .grad from autograd /
 * optimizer.zero_grad, projects/NeuralLumen/trainer.py:167): a persistent grid of `n_ctas` CTAs (<= 0: 32) instead of
 * one that floods every SM, so that when it is issued on a side stream next to the latency-bound sampling rounds the
 * main stream's small kernels still find free SM slots (a full-grid fill delayed them by its whole duration).
 * ptr must be 16-byte aligned; any byte count.
 * ---------------------------------------------------------------------------------------------------- */
int mli_zero_fill_background(void* ptr, int64_t bytes, int32_t n_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MLI_B200_H */

// hostsim.cpp -- TEST-ONLY harness: compiles the __host__ __device__ math of mli_nerf_b200/csrc/mli_math.h with g++
// so the exact formulas the CUDA kernels execute can be checked against the oracle on a box without a GPU.
// It is never loaded by the product (mli_nerf_b200 has no CPU path); only tests/test_hostsim_math.py uses it.
#include <cstring>
#include "../../mli_nerf_b200/csrc/mli_math.h"

extern "C" {

void hs_corners(const mli_grid_t* grid, uint32_t level, const float* x01, int64_t M, uint32_t* idx, float* wts) {
  const mli_level_t& lv = grid->level[level];
  for (int64_t m = 0; m < M; ++m) {
    mli_cell_t c = mli_grid_cell(lv, x01[m * 3], x01[m * 3 + 1], x01[m * 3 + 2]);
    for (int k = 0; k < 8; ++k) mli_corner(lv, c, k, &idx[m * 8 + k], &wts[m * 8 + k]);
  }
}

void hs_bounds(const float* c, const float* r, int64_t R, const float* aabb, float* near, float* far, uint8_t* out) {
  for (int64_t i = 0; i < R; ++i) {
    if (aabb) mli_bounds_aabb(c + i * 3, r + i * 3, aabb, near + i, far + i, out + i);
    else mli_bounds_sphere(c + i * 3, r + i * 3, near + i, far + i, out + i);
  }
}

void hs_points(const float* c, const float* r, const float* d, int64_t n, int taps, int plane, float eps, float* p) {
  for (int64_t i = 0; i < n; ++i) mli_sample_point(c + i * 3, r + i * 3, d[i], taps, plane, eps, p + i * 3);
}

void hs_sample_fine(const float* dists, const float* sdfs, int64_t ld, int64_t R, int n, int n_fine, float inv_s,
                    float* fine, int32_t* idx, int32_t* low, int32_t* high, float* cdf_o, float* w_o) {
  float w[512], cdf[513];
  for (int64_t r = 0; r < R; ++r) {
    const float* d = dists + r * ld;
    const float* s = sdfs + r * ld;
    mli_hier_weights(d, s, n, inv_s, w);
    mli_weights_to_cdf(w, n - 1, cdf);
    for (int j = 0; j < n_fine; ++j) {
      int a, b, c;
      fine[r * n_fine + j] = mli_sample_bin(d, cdf, n, mli_unif(j, n_fine), &a, &b, &c);
      idx[r * n_fine + j] = a; low[r * n_fine + j] = b; high[r * n_fine + j] = c;
    }
    if (cdf_o) memcpy(cdf_o + r * n, cdf, sizeof(float) * n);
    if (w_o) memcpy(w_o + r * (n - 1), w, sizeof(float) * (n - 1));
  }
}

void hs_pdf_bins(const float* weights, int64_t ld_w, int64_t R, int n_w, int n_fine, int32_t* idx, int32_t* low,
                 int32_t* high, float* cdf_o) {
  float cdf[513], d[513];
  for (int i = 0; i <= n_w; ++i) d[i] = (float)i;
  for (int64_t r = 0; r < R; ++r) {
    mli_weights_to_cdf(weights + r * ld_w, n_w, cdf);
    for (int j = 0; j < n_fine; ++j) {
      int a, b, c;
      mli_sample_bin(d, cdf, n_w + 1, mli_unif(j, n_fine), &a, &b, &c);
      idx[r * n_fine + j] = a; low[r * n_fine + j] = b; high[r * n_fine + j] = c;
    }
    if (cdf_o) memcpy(cdf_o + r * (n_w + 1), cdf, sizeof(float) * (n_w + 1));
  }
}

void hs_unif(int n_fine, float* u) { for (int j = 0; j < n_fine; ++j) u[j] = mli_unif(j, n_fine); }

void hs_sh16(const float* d, int64_t n, float* out) { for (int64_t i = 0; i < n; ++i) mli_sh16(d[i * 3], d[i * 3 + 1], d[i * 3 + 2], out + i * 16); }

void hs_act(const float* x, int64_t n, int act, float* y, float* dy_from_out) {
  for (int64_t i = 0; i < n; ++i) { y[i] = mli_act(x[i], act); dy_from_out[i] = mli_dact_from_out(y[i], act); }
}

// alpha forward + backward for n independent samples (d_alpha = 1)
void hs_neus_alpha(const float* sdf, const float* g, const float* r, const float* intv, int64_t n, float inv_s,
                   float anneal, float* alpha, float* d_sdf, float* d_g, float* d_inv_s) {
  for (int64_t i = 0; i < n; ++i) {
    mli_alpha_t a = mli_neus_alpha(sdf[i], g + i * 3, r + i * 3, intv[i], inv_s, anneal);
    alpha[i] = a.alpha;
    float dg[3] = {0.f, 0.f, 0.f};
    d_inv_s[i] = mli_neus_alpha_bwd(a, sdf[i], r + i * 3, inv_s, anneal, 1.0f, d_sdf + i, dg);
    d_g[i * 3] = dg[0]; d_g[i * 3 + 1] = dg[1]; d_g[i * 3 + 2] = dg[2];
  }
}

}  // extern "C"

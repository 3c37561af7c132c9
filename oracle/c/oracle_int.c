/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the integer-deciding parts of the hot path, used as a second,
 * independent checker next to oracle/port.py (torch).  Never linked into or called by the product.
 *
 *   oc_grid_rows   tcnn HashGrid corner rows (grid.h: grid_index / grid_hash with the coherent primes; see
 *                  oracle/torch_hashgrid.py -- tcnn is not under /root/reference: "parity unpinned").
 *   oc_pdf_bins    nerf_util.sample_dists_from_pdf (/root/reference/projects/nerf/utils/nerf_util.py:41-68):
 *                  L1 normalise (sequential float32 sum), cumsum (float64 accumulator, float32 store),
 *                  searchsorted(right=True), low/high clamps, linear interpolation.
 */
#include <math.h>
#include <stdint.h>

typedef struct { float scale; uint32_t res, size, offset, hashed; } oc_level_t;

static uint32_t grid_index(const oc_level_t* lv, const uint32_t g[3]) {
  uint32_t index = 0;
  if (lv->hashed) {
    index = g[0] ^ (g[1] * 2654435761u) ^ (g[2] * 805459861u);
  } else {
    uint32_t stride = 1;
    for (int d = 0; d < 3 && stride <= lv->size; ++d) { index += g[d] * stride; stride *= lv->res; }
  }
  return index % lv->size;
}

void oc_grid_rows(const oc_level_t* lv, const float* x01, int64_t n, uint32_t* rows /* [n,8] */) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t cell[3];
    for (int d = 0; d < 3; ++d) {
      float pos = fmaf(lv->scale, x01[i * 3 + d], 0.5f);
      cell[d] = (uint32_t)(int)floorf(pos);
    }
    for (int c = 0; c < 8; ++c) {
      uint32_t g[3];
      for (int d = 0; d < 3; ++d) g[d] = cell[d] + ((c >> d) & 1u);
      rows[i * 8 + c] = lv->offset + grid_index(lv, g);
    }
  }
}

void oc_pdf_bins(const float* weights, const float* bins, int64_t rays, int n_w, int n_fine, float* cdf_out /* [rays,n_w+1] */,
                 int32_t* idx_out /* [rays,n_fine] */, float* dists_out /* [rays,n_fine] */) {
  for (int64_t r = 0; r < rays; ++r) {
    const float* w = weights + r * n_w;
    const float* b = bins + r * (n_w + 1);
    float* cdf = cdf_out + r * (n_w + 1);
    volatile float denom = 0.0f;
    for (int i = 0; i < n_w; ++i) denom = denom + fabsf(w[i]);
    float den = denom < 1e-12f ? 1e-12f : denom;
    double acc = 0.0;
    cdf[0] = 0.0f;
    for (int i = 0; i < n_w; ++i) { volatile float pdf = w[i] / den; acc += (double)pdf; cdf[i + 1] = (float)acc; }
    const int n = n_w + 1;
    for (int j = 0; j < n_fine; ++j) {
      /* unif = 0.5*(grid[j]+grid[j+1]), grid = linspace(0,1,n_fine+1); exact for n_fine = 16 */
      volatile float g0 = (float)j / (float)n_fine, g1 = (float)(j + 1) / (float)n_fine;
      volatile float u = 0.5f * (g0 + g1);
      int lo = 0, hi = n;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
      int idx = lo, low = idx - 1 < 0 ? 0 : idx - 1, high = idx > n - 1 ? n - 1 : idx;
      volatile float num = u - cdf[low];
      volatile float dc = cdf[high] - cdf[low];
      volatile float dcp = dc + 1e-8f;
      volatile float t = num / dcp;
      volatile float span = b[high] - b[low];
      volatile float ts = t * span;
      idx_out[r * n_fine + j] = idx;
      dists_out[r * n_fine + j] = b[low] + ts;
    }
  }
}

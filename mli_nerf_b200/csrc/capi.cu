// capi.cu -- error state, device check, host-side helpers of the C ABI (include/mli_b200.h)
#include <math.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void mli_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mli_last_error(void) { return g_err; }

static int g_sm_limit = MLI_NUM_SMS;
int mli_sm_limit() { return g_sm_limit; }
extern "C" int mli_set_sm_limit(int32_t n_sms) {
  MLI_REQUIRE(n_sms >= 8 && n_sms <= MLI_NUM_SMS, "sm limit must be in [8, %d]", MLI_NUM_SMS);
  g_sm_limit = n_sms;
  return MLI_OK;
}
extern "C" int mli_abi_version(void) { return MLI_ABI_VERSION; }

extern "C" int mli_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

int mli_check_device() {
  static thread_local int cached = -1;
  if (cached == 1) return MLI_OK;
  if (!mli_device_ok()) {
    mli_set_error("libmli_b200 needs a compute-capability 10.x (B200, sm_100a) device; there is no CPU fallback");
    return MLI_ENODEV;
  }
  cached = 1;
  return MLI_OK;
}

// tcnn GridEncodingTemplated constructor + grid_scale()/grid_resolution() (float32 arithmetic), see
// oracle/torch_hashgrid.py for the formulas this follows.
extern "C" int mli_grid_init(mli_grid_t* grid, uint32_t n_levels, uint32_t feat, uint32_t log2_hashmap_size,
                             uint32_t base_resolution, float per_level_scale) {
  MLI_REQUIRE(grid != nullptr, "grid is NULL");
  MLI_REQUIRE(n_levels >= 1 && n_levels <= MLI_MAX_LEVELS, "n_levels %u out of range", n_levels);
  MLI_REQUIRE(feat == 8 || feat == 4 || feat == 2, "n_features_per_level must be 2, 4 or 8 (got %u)", feat);
  MLI_REQUIRE(log2_hashmap_size >= 4 && log2_hashmap_size <= 28, "log2_hashmap_size %u out of range",
              log2_hashmap_size);
  memset(grid, 0, sizeof(*grid));
  grid->n_levels = n_levels;
  grid->feat = feat;
  grid->active_levels = n_levels;
  // tcnn evaluates exp2f(l * log2f(s)) * base - 1 in float32 on the device; its last-ulp rounding is platform
  // dependent, so oracle and product both use the float64 formula rounded once to float32 (bit-reproducible).
  const double log2_pls = log2((double)per_level_scale);
  uint64_t offset = 0;
  for (uint32_t l = 0; l < n_levels; ++l) {
    mli_level_t& lv = grid->level[l];
    lv.scale = (float)(exp2((double)l * log2_pls) * (double)base_resolution - 1.0);
    lv.res = (uint32_t)ceilf(lv.scale) + 1u;
    const uint32_t max_params = 0xFFFFFFFFu / 2u;
    double dense = (double)lv.res * (double)lv.res * (double)lv.res;
    uint32_t n = dense > (double)max_params ? max_params : (uint32_t)dense;
    n = (n + 7u) / 8u * 8u;
    uint32_t cap = 1u << log2_hashmap_size;
    lv.size = n < cap ? n : cap;
    uint64_t stride = 1;
    for (int d = 0; d < 3 && stride <= lv.size; ++d) stride *= lv.res;
    lv.hashed = lv.size < stride ? 1u : 0u;
    lv.offset = (uint32_t)offset;
    offset += lv.size;
    MLI_REQUIRE(offset < 0xFFFFFFFFull, "hash table too large for 32-bit row indices");
  }
  grid->n_entries = (uint32_t)offset;
  return MLI_OK;
}

// L2 residency control for the dense levels of the hash table (north star (a): "L2-resident tables"): an access-policy
// window on `stream` marks [ptr, ptr + bytes) as persisting in L2 with probability hit_ratio (the rest of the stream's
// traffic -- the 1.3 GB of hashed levels, activations -- is unaffected); bytes == 0 removes the window.  The persisting
// carve-out of L2 is raised to what the window needs (capped by the device limit).  Kernels captured into a CUDA graph
// from this stream inherit the window.
extern "C" int mli_set_l2_window(const void* ptr, int64_t bytes, float hit_ratio, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(bytes >= 0 && hit_ratio >= 0.0f && hit_ratio <= 1.0f, "l2_window: bad bytes / hit_ratio");
  int dev = 0;
  MLI_CUDA_OK(cudaGetDevice(&dev));
  int max_win = 0, max_persist = 0;
  MLI_CUDA_OK(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  MLI_CUDA_OK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (bytes > 0 && ptr != nullptr) {
    size_t win = (size_t)bytes < (size_t)max_win ? (size_t)bytes : (size_t)max_win;
    size_t carve = (size_t)((double)win * hit_ratio);
    if (carve > (size_t)max_persist) carve = (size_t)max_persist;
    MLI_CUDA_OK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
    attr.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = hit_ratio;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;
  }
  MLI_CUDA_OK(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  return MLI_OK;
}

extern "C" int mli_l2_info(int32_t* host_out_l2_bytes, int32_t* host_out_max_persist_bytes, int32_t* host_out_max_window_bytes) {
  MLI_ENTRY();
  int dev = 0, v = 0;
  MLI_CUDA_OK(cudaGetDevice(&dev));
  if (host_out_l2_bytes) { MLI_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev)); *host_out_l2_bytes = v; }
  if (host_out_max_persist_bytes) { MLI_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev)); *host_out_max_persist_bytes = v; }
  if (host_out_max_window_bytes) { MLI_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, dev)); *host_out_max_window_bytes = v; }
  return MLI_OK;
}

namespace {
__global__ void __launch_bounds__(256) zero_fill_kernel(uint4* __restrict__ dst, int64_t n16, uint8_t* __restrict__ tail, int n_tail) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 4
  for (; i < n16; i += stride) __stcs(dst + i, z);  // streaming stores: nothing of this buffer is worth keeping in L2
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] = 0;
}
}  // namespace

extern "C" int mli_zero_fill_background(void* ptr, int64_t bytes, int32_t n_ctas, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(bytes >= 0 && (ptr != nullptr || bytes == 0), "zero_fill: bad arguments");
  MLI_REQUIRE(((uintptr_t)ptr & 15u) == 0, "zero_fill: ptr must be 16-byte aligned");
  if (bytes == 0) return MLI_OK;
  const int64_t n16 = bytes / 16;
  const int n_tail = (int)(bytes - n16 * 16);
  int ctas = n_ctas > 0 ? n_ctas : 32;
  const int64_t need = (n16 + 255) / 256;
  if (need < ctas) ctas = need > 0 ? (int)need : 1;
  zero_fill_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>((uint4*)ptr, n16, (uint8_t*)ptr + n16 * 16, n_tail);
  MLI_LAUNCH_OK();
  return MLI_OK;
}


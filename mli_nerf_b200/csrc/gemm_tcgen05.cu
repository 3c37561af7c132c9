// gemm_tcgen05.cu -- bf16 tensor-core (tcgen05 / TMEM / TMA) path of the dense layers.  PLACEHOLDER: filled in
// once the fp32 path is parity-green on the GPU.
#include "common.cuh"

int mli_tc_linear_fwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, const float*, int64_t, float*,
                      int64_t, int64_t, int64_t, int32_t, int32_t, int32_t, int32_t, void*) {
  mli_set_error("bf16 tcgen05 path not built yet");
  return MLI_ENOTSUP;
}

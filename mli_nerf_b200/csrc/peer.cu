// peer.cu -- device side of the multi-GPU hash-table gradient exchange over NVLink peer memory (mli_nerf_b200/dist.py):
// every rank owns 1/W of each gradient slab, pulls that shard of every peer's buffer with the copy engines (no SMs, so
// the weight-gradient GEMMs of the same step keep the whole GPU), sums it here and pushes the mean back.
// Replaces the NCCL all-reduce DDP runs for this gradient in the reference
// (/root/reference/imaginaire/trainers/utils/get_trainer.py:80-88).
#include <string.h>

#include "common.cuh"

extern "C" int mli_enable_peer_access(int32_t peer_device) {
  int dev = 0, can = 0;
  MLI_CUDA_OK(cudaGetDevice(&dev));
  if (peer_device == dev) return MLI_OK;
  MLI_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  MLI_REQUIRE(can == 1, "device %d cannot access device %d over NVLink/PCIe peer-to-peer", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return MLI_OK;
  }
  MLI_CUDA_OK(e);
  return MLI_OK;
}

// A buffer other ranks of the node can map: plain cudaMalloc (its own allocation, so the IPC handle addresses it from
// offset 0) + the 64-byte CUDA IPC handle to hand to the peers.
extern "C" int mli_peer_alloc(int64_t bytes, void** host_out_ptr, void* host_out_handle64) {
  MLI_REQUIRE(bytes > 0 && host_out_ptr != nullptr && host_out_handle64 != nullptr, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  void* p = nullptr;
  MLI_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    MLI_CUDA_OK(e);
  }
  memcpy(host_out_handle64, &h, 64);
  *host_out_ptr = p;
  return MLI_OK;
}

// Map a peer rank's buffer into the CURRENT device's address space (peer access over NVLink enabled by the driver).
extern "C" int mli_peer_open(const void* host_handle64, void** host_out_ptr) {
  MLI_REQUIRE(host_handle64 != nullptr && host_out_ptr != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, host_handle64, 64);
  void* p = nullptr;
  MLI_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *host_out_ptr = p;
  return MLI_OK;
}

extern "C" int mli_peer_close(void* mapped_ptr) {
  if (mapped_ptr != nullptr) MLI_CUDA_OK(cudaIpcCloseMemHandle(mapped_ptr));
  return MLI_OK;
}

extern "C" int mli_peer_free(void* ptr) {
  if (ptr != nullptr) MLI_CUDA_OK(cudaFree(ptr));
  return MLI_OK;
}

extern "C" int mli_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  MLI_REQUIRE(bytes >= 0, "copy: negative size");
  if (bytes == 0) return MLI_OK;
  MLI_REQUIRE(dst != nullptr && src != nullptr, "copy: null pointer");
  MLI_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return MLI_OK;
}

// dst[i] = (dst[i] + sum_k slots[k * slot_stride + i]) * scale: 4 elements per thread and pass, all slot loads of a
// pass issued before the first add
template <int NS>
__global__ void __launch_bounds__(256) reduce_slots_kernel(float* __restrict__ dst, const float* __restrict__ slots,
                                                           int64_t slot_stride, int64_t n4, int64_t n, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = reinterpret_cast<float4*>(dst)[i];
    float4 v[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) v[k] = __ldg(reinterpret_cast<const float4*>(slots + (int64_t)k * slot_stride) + i);
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(dst)[i] = acc;
  }
  const int64_t t = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // scalar tail
  if (t < n) {
    float a = dst[t];
#pragma unroll
    for (int k = 0; k < NS; ++k) a += slots[(int64_t)k * slot_stride + t];
    dst[t] = a * scale;
  }
}

extern "C" int mli_reduce_slots(float* dst, const float* slots, int32_t n_slots, int64_t slot_stride, int64_t n,
                                float scale, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(n_slots >= 1 && n_slots <= 7, "reduce_slots: 1..7 slots (world size 2..8)");
  MLI_REQUIRE(n >= 0 && slot_stride >= n, "reduce_slots: bad sizes");
  if (n == 0) return MLI_OK;
  MLI_REQUIRE(((uintptr_t)dst | (uintptr_t)slots) % 16 == 0 && slot_stride % 4 == 0, "reduce_slots: 16-byte alignment");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 8 * MLI_NUM_SMS) blocks = 8 * MLI_NUM_SMS;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = (cudaStream_t)stream;
#define MLI_RS(NS) case NS: reduce_slots_kernel<NS><<<(unsigned)blocks, 256, 0, st>>>(dst, slots, slot_stride, n4, n, scale); break;
  switch (n_slots) {
    MLI_RS(1) MLI_RS(2) MLI_RS(3) MLI_RS(4) MLI_RS(5) MLI_RS(6) MLI_RS(7)
  }
#undef MLI_RS
  MLI_LAUNCH_OK();
  return MLI_OK;
}

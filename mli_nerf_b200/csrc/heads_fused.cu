// heads_fused.cu -- the whole head stack of LumenRGB on chip: one persistent tcgen05 kernel per 256-sample tile pair.
//
// Replaces, for the forward pass, the five torch.nn.Linear (+ReLU, +sigmoid) layers of every head
// (/root/reference/projects/nerf/utils/nerf_util.py:186-196 MLPwithSkipConnection.forward, called three times by
// /root/reference/projects/NeuralLumen/utils/modules.py:106-174 LumenRGB.forward) that the layer-by-layer path runs as
// four launches with every [M x 768] bf16 activation written to and re-read from HBM.
//
// Per CTA (one per SM, persistent over tile pairs), per head h:
//     L0 : A0 = relu(XH[256 x 304] Wh0[h]^T + b)      XH K-slices and weights streamed through the TMA ring
//     L1..L3 : A_l = relu(A_{l-1} W_l[h]^T + b)       A operand = the activation tile in shared memory (never leaves the SM
//                                                     between layers), weights streamed from L2 through the ring
//     out: S[:, j] = act(w_out[j] . A3 + b_out[j])    row dot fused into L3's epilogue
// Shared memory: two 64 KB activation tiles (rows 0..127 / 128..255 of the pair; UMMA K-major core-matrix layout =
// byte-identical to one TCL tile, so a finished layer leaves the SM as ONE 32 KB TMA bulk store per column half, and only
// when a backward pass will need it), an 8 x 8 KB ring (hidden-layer weights: [128 x 32] bf16 = one N-half x K32 slice;
// layer 0: [256 x 16] weight slices and [128 x 32] XH slices), biases and output-layer rows.  TMEM: four 128-column fp32
// accumulators (tile x N-half).
// Why two tiles per CTA: every 8 KB weight slice read from L2 feeds 4 MMAs (2 tiles x 2 k-steps) = 256 tensor cycles, i.e.
// 32 B/clk/SM of L2 traffic -- half of what one tile per CTA would need and under the ~42 B/clk/SM the L2 sustains with all
// 148 SMs pulling.  The MMA order inside a hidden layer is (K-half 0: N0, N1), (K-half 1: N0, N1), each for both tiles:
// the epilogue of column half N0 (which OVERWRITES activation columns 0..127 = K-half 0 of the next layer's operand)
// starts while the K-half-1 MMAs of N1 still run, and the next layer's K-half-0 MMAs start before N1's epilogue ends.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocator), 2..17 = four epilogue groups of four warps, one per
// (tile, column half) accumulator (a warp may only touch TMEM lane quarter warp % 4).
// Measured lesson (profiles/r02_heads_fused.md): the first version issued its MMAs from inside an `if (lane == 0)` region
// with modulo-9 ring arithmetic; ncu showed the tensor pipe 23 % active with the eight epilogue warps parked on the
// accumulator barrier and the issuing warp busy 100 % of the time (R2UR broadcast loops around every tcgen05 instruction).
// The issuer now runs warp-uniform code (operands stay in uniform registers, only the tcgen05 instructions themselves are
// predicated on lane 0), layer 0 uses N = 256 MMAs, and ring positions are carried incrementally.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kRows = 128;            // rows of one activation tile (UMMA M)
constexpr int kHid = 256;
constexpr int kSlotBytes = 8192;      // [128 x 32] bf16
constexpr int kSlots = 8;             // power of two: ring positions advance with a mask
constexpr int kEpiGroups = 4;         // (tile, column half)
constexpr int kThreads = 64 + 128 * kEpiGroups;   // 18 warps
constexpr int kMaxHeads = 3;
constexpr int kMaxJ = 12;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol error must surface as a device-side trap with the wait site, not as a hung GPU (a kernel that
// never returns costs the whole box).  try_wait suspends the thread for a hardware time slice per poll; 2^24 polls are
// several seconds, orders of magnitude beyond any legitimate wait in this kernel.
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_timeout(int site, uint32_t parity) {
  printf("tc_heads_kernel: wait timed out at site %d (parity %u) block %d thread %d\n", site, parity, (int)blockIdx.x,
         (int)threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int site) {
  for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
    if (spin > (1u << 24)) mbar_timeout(site, parity);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// TMA bulk store shared -> global (SASS: UBLKCP.G.S), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int eg) {  // the 128 threads of one epilogue group (named barriers 1..4)
  asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
}
__device__ __forceinline__ void tile_bar(int t) {    // both column-half groups of one tile (named barriers 5, 6)
  asm volatile("bar.sync %0, 256;" ::"r"(5 + t) : "memory");
}

// un-swizzled K-major shared-memory matrix descriptor (sm_100 version bits): LBO = bytes between 8-column chunks, SBO =
// bytes between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// the same descriptor with the start address left out: desc = desc_base(lbo, sbo) + (smem byte address >> 4)
__device__ __forceinline__ uint64_t desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = n, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// relu on a packed pair of bf16 (the pack rounds first: max(round(x), 0) == round(max(x, 0)) for round-to-nearest)
__device__ __forceinline__ uint32_t relu_pack2(float a, float b) {
  __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(a, b), __floats2bfloat162_rn(0.0f, 0.0f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}

struct HeadsArgs {
  // ---- forward ----
  const __nv_bfloat16* XH; int xh_chunks;         // TCL-128 [tiles][xh_chunks][128][8], K0 = 8 * k0_chunks columns used
  int k0_chunks;
  const __nv_bfloat16* W0;                        // [nh][k0_chunks][256][8] (TCL with 256-row tiles: N = 256 MMAs)
  const float* bias[4];                           // [nh * 256] per layer
  const float* bout;                              // [J]
  int act_out; uint32_t act_mask;
  float* S;                                       // [M, lds] head outputs (forward) ...
  // ---- backward ----
  const float* dS;                                // ... or [M, lds] gradient w.r.t. the pre-activation head outputs
  // ---- both ----
  const __nv_bfloat16* Wl[3];   // forward: W1..W3, backward: W3^T, W2^T, W1^T in use order; [nh][2 N-halves][32][128][8]
  const float* wout;            // [J][256] fp32 (output layers of all heads, head after head)
  int nh, J; int j0[kMaxHeads], nj[kMaxHeads];
  __nv_bfloat16* A[4];          // forward: activations A0..A3 (NULL: not stored); backward: dZ3, dZ2, dZ1, dZ0 in use order
  uint32_t* mask[4];            // relu sign bits [tiles][nh*8][128]: written by forward (may be NULL), read by backward
  int64_t lds;
  int64_t M; int n_tiles;
  int debug;   // timing experiments only (MLI_HF_DEBUG): bit 0 no wait for the bulk store, bit 1 no masks, bit 2 no bulk store
};

// Ring items the producer streams per head and tile pair (the MMA warp consumes them in the same order):
//   forward L0, per K32 step: XH slice of tile 0, XH slice of tile 1 ([128 x 32]), then one [256 x 16] W0 slice per K16 step
//   every 256x256 layer, per quarter (K-half, N-half) in the order (0,0) (0,1) (1,0) (1,1): four [128 x 32] weight slices
// Stages per head: forward  L0, L1, L2, L3 (+ output dot);  backward  P (dZ3 from dS, CUDA cores), L3', L2', L1'.
// `acc_full[tile][half]` completes once per stage that has MMAs.  `ready[tile][half]` ("accumulator drained, activation
// half written, next stage's bias stored") is arrived by the owning epilogue group
//   forward:  once before the first stage of the kernel (bias of L0 stored into the accumulator) and after every stage;
//   backward: after P, L3', L2' -- NOT after L1': the prologue of the next head needs nothing from the MMA warp, so an
//             arrival after L1' could be followed by the prologue's arrival before the MMA warp has looked at the
//             barrier, and a parity wait cannot tell a phase that completed twice from one that has not completed.
// The MMA warp waits for exactly one new arrival per (tile, half) before the first MMA of every stage.
// Forward bias: the accumulators are pre-loaded with the NEXT stage's bias by the epilogue (tcgen05.st) and every MMA
// accumulates -- 32 FADDs per 32-column chunk and row leave the epilogue, which is what limits this kernel.
template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1) tc_heads_kernel(const __grid_constant__ HeadsArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * kSlots + 8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = p.nh;
  const uint32_t sAct = smem_u32(smem);                       // 2 x 64 KB
  const uint32_t sRing = sAct + 2 * 65536;                    // kSlots x 8 KB
  float* s_bias = reinterpret_cast<float*>(smem + 2 * 65536 + kSlots * kSlotBytes);   // [4][nh][256] (forward only)
  float* s_wout = s_bias + (BWD ? 0 : 4 * nh * kHid);         // [J][256]
  float* s_bout = s_wout + p.J * kHid;                        // [16]
  float* s_dotp = s_bout + 16;                                // [2 tiles][4][128]
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kSlots]);
  const uint32_t acc_full0 = smem_u32(&bars[2 * kSlots]);     // [tile][N-half]
  const uint32_t ready0 = smem_u32(&bars[2 * kSlots + 4]);    // [tile][N-half]: accumulator drained + activation half written
  const int n_pairs = (p.n_tiles + 1) / 2;
  const int k0_steps = BWD ? 0 : (p.k0_chunks + 3) / 4;       // K32 steps of layer 0 (the last one may hold 2 chunks)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(acc_full0 + 8 * i, 1); mbar_init(ready0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!BWD)
    for (int i = threadIdx.x; i < 4 * nh * kHid; i += kThreads) s_bias[i] = p.bias[i / (nh * kHid)][i % (nh * kHid)];
  for (int i = threadIdx.x; i < p.J * kHid; i += kThreads) s_wout[i] = p.wout[i];
  if (threadIdx.x < 16) s_bout[threadIdx.x] = (!BWD && threadIdx.x < p.J && p.bout) ? p.bout[threadIdx.x] : 0.0f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer ------------------------------------
    if (lane == 0) {
      uint32_t slot = 0, par = 1;  // waiting for parity 1 on a fresh barrier passes at once: the first round finds the ring empty
      auto put = [&](const void* src, uint32_t bytes) {
        mbar_wait(empty0 + 8 * slot, par, 1);
        mbar_expect_tx(full0 + 8 * slot, bytes);
        bulk_g2s(sRing + slot * kSlotBytes, src, bytes, full0 + 8 * slot);
        slot = (slot + 1) & (kSlots - 1);
        par ^= (slot == 0);
      };
      for (int pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
        const int64_t t0 = 2 * (int64_t)pr, t1 = (t0 + 1 < p.n_tiles) ? t0 + 1 : t0;  // odd tile count: tile 0 twice
        for (int h = 0; h < nh; ++h) {
          for (int ks = 0; ks < k0_steps; ++ks) {
            const int nch = min(4, p.k0_chunks - ks * 4);
            put(p.XH + (t0 * p.xh_chunks + ks * 4) * 1024, nch * 2048);
            put(p.XH + (t1 * p.xh_chunks + ks * 4) * 1024, nch * 2048);
            for (int kk = 0; kk < nch / 2; ++kk)  // [256 rows x 16 columns] = 2 chunks of the 256-row-tile layout
              put(p.W0 + ((int64_t)h * p.k0_chunks + ks * 4 + kk * 2) * 2048, 8192);
          }
          for (int l = 0; l < 3; ++l)
            for (int q = 0; q < 4; ++q) {
              const int kh = q >> 1, nf = q & 1;
              for (int it = 0; it < 4; ++it)
                put(p.Wl[l] + ((int64_t)(h * 2 + nf) * 32 + kh * 16 + it * 4) * 1024, kSlotBytes);
            }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer --------------------------------------
    // Warp-uniform control flow: all 32 lanes run the loops (every operand stays in uniform registers), lane 0 alone
    // executes the tcgen05 instructions (tcgen05.commit tracks the MMAs of the thread that issues it).
    // elect.sync names the same leader for the same member mask every time, and tells the compiler that exactly one lane
    // runs the guarded instructions while their operands were computed warp-uniformly
    auto elected = []() {
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      return pred != 0;
    };
    const uint32_t idesc128 = make_idesc(128), idesc256 = make_idesc(256);
    const uint64_t dA = desc_base(2048, 128);      // [128 rows][16 B] chunks, 2048 B apart (activation tiles, 128-row slots)
    const uint64_t dB256 = desc_base(4096, 128);   // [256 rows][16 B] chunks, 4096 B apart (layer-0 weight slices)
    uint32_t slot = 0, par = 0, na = 0;            // ring position / `ready` arrivals consumed so far (per barrier)
    auto advance = [&]() { slot = (slot + 1) & (kSlots - 1); par ^= (slot == 0); };
    for (int pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
      for (int h = 0; h < nh; ++h) {
        if (!BWD) {
          // ---- layer 0: both operands from the ring, k-major, N = 256 (all four accumulators finish together);
          //      accumulators hold the bias already ----
          for (int i = 0; i < 4; ++i) mbar_wait(ready0 + 8 * i, na & 1, 3);
          ++na;
          tc_fence_after();
          for (int ks = 0; ks < k0_steps; ++ks) {
            const int nk = min(4, p.k0_chunks - ks * 4) / 2;  // K16 steps in this K32 step
            const uint32_t sa0 = sRing + slot * kSlotBytes, s0 = slot;
            mbar_wait(full0 + 8 * slot, par, 2); advance();
            const uint32_t sa1 = sRing + slot * kSlotBytes, s1 = slot;
            mbar_wait(full0 + 8 * slot, par, 2); advance();
            for (int kk = 0; kk < nk; ++kk) {
              const uint32_t sb = sRing + slot * kSlotBytes, sbs = slot;
              mbar_wait(full0 + 8 * slot, par, 2); advance();
              tc_fence_after();
              if (elected()) {
                umma(tmem_base, dA + ((sa0 + kk * 4096) >> 4), dB256 + (sb >> 4), idesc256, 1u);
                umma(tmem_base + 256, dA + ((sa1 + kk * 4096) >> 4), dB256 + (sb >> 4), idesc256, 1u);
                umma_commit(empty0 + 8 * sbs);
              }
            }
            if (elected()) { umma_commit(empty0 + 8 * s0); umma_commit(empty0 + 8 * s1); }
          }
          if (elected())
            for (int i = 0; i < 4; ++i) umma_commit(acc_full0 + 8 * i);
        }
        // ---- three 256x256 layers: A operand = the activation tiles in shared memory ----
        for (int l = 0; l < 3; ++l) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kh = q >> 1, nf = q & 1;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const uint32_t sb = sRing + slot * kSlotBytes, sbs = slot;
              mbar_wait(full0 + 8 * slot, par, 2); advance();
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                if (kh == 0 && it == 0) {
                  // first write of accumulator (t, nf) in this stage: its previous contents have been drained (forward:
                  // and replaced by this stage's bias), and activation column half nf of tile t (= K-half nf of this
                  // stage's A operand; K-half 0 is needed from quarter (0,0) on, K-half 1 from quarter (1,0) on, i.e.
                  // after the wait of quarter (0,1)) has been written
                  mbar_wait(ready0 + 8 * (t * 2 + nf), na & 1, 4);
                }
                tc_fence_after();
                if (elected()) {
                  const uint32_t sa = sAct + t * 65536 + (kh * 16 + it * 4) * 2048;
                  umma(tmem_base + (t * 2 + nf) * 128, dA + (sa >> 4), dA + (sb >> 4), idesc128, BWD ? (uint32_t)((kh | it) != 0) : 1u);
                  umma(tmem_base + (t * 2 + nf) * 128, dA + ((sa + 4096) >> 4), dA + ((sb + 4096) >> 4), idesc128, 1u);
                  if (kh == 1 && it == 3) umma_commit(acc_full0 + 8 * (t * 2 + nf));
                }
              }
              if (elected()) umma_commit(empty0 + 8 * sbs);
            }
          }
          ++na;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue ----------------------------------------
    const int eg = (warp - 2) >> 2;           // epilogue group = accumulator it owns
    const int t = eg >> 1, g = eg & 1;        // tile of the pair, column half
    const int q = warp & 3;                   // TMEM lane quarter
    const int r_local = q * 32 + lane;
    const bool leader = (threadIdx.x == 64 + eg * 128);
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const bool store = BWD || p.A[0] != nullptr;
    uint8_t* act = smem + t * 65536 + (size_t)(g * 16) * 2048 + r_local * 16;
    uint32_t n_acc = 0;  // stages with MMAs passed so far (`acc_full` flips once per such stage)
    // bias of the stage that comes after (h, l) in this CTA's stage sequence (forward); NULL after the very last one
    auto next_bias = [&](int pr, int h, int l) -> const float* {
      if (l < 3) return s_bias + ((l + 1) * nh + h) * kHid + g * 128;
      if (h + 1 < nh) return s_bias + (h + 1) * kHid + g * 128;
      return (pr + (int)gridDim.x < n_pairs) ? s_bias + g * 128 : nullptr;
    };
    auto store_bias = [&](const float* b) {  // accumulator (t, g) <- b broadcast over the rows
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b4 = *reinterpret_cast<const float4*>(b + c * 32 + i4 * 4);
          v[i4 * 4 + 0] = b4.x; v[i4 * 4 + 1] = b4.y; v[i4 * 4 + 2] = b4.z; v[i4 * 4 + 3] = b4.w;
        }
        tmem_st32(tmem_base + lane_addr + eg * 128 + c * 32, v);
      }
      tmem_st_wait();
    };
    if (!BWD && (int)blockIdx.x < n_pairs) {  // forward stage "-1": layer-0 bias of the first head into the accumulators
      store_bias(s_bias + g * 128);
      tc_fence_before();
      group_bar(eg);
      if (leader) mbar_arrive(ready0 + 8 * eg);
    }
    for (int pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
      const int64_t tile = 2 * (int64_t)pr + t;
      const bool live = tile < p.n_tiles;
      const int64_t row = tile * kRows + r_local;
      for (int h = 0; h < nh; ++h) {
        const float* wd = s_wout + p.j0[h] * kHid + g * 128;
        const int njh = p.nj[h];
        const int64_t mrow = (tile * (nh * 8) + h * 8 + g * 4) * kRows + r_local;  // + c * kRows per 32-column chunk
        for (int l = 0; l < 4; ++l) {
          const bool has_acc = !(BWD && l == 0);
          const float* nbias = BWD ? nullptr : next_bias(pr, h, l);
          // relu sign bits this stage reads (backward: of the activation whose pre-activation gradient it produces)
          const uint32_t* mask_in = BWD ? p.mask[3 - l] : nullptr;
          uint32_t* mask_out = BWD ? nullptr : p.mask[l];
          uint32_t mb0 = 0u, mb1 = 0u, mb2 = 0u, mb3 = 0u;
          float ds[4] = {0.f, 0.f, 0.f, 0.f};
          if (BWD && live) {
            mb0 = __ldg(mask_in + mrow); mb1 = __ldg(mask_in + mrow + kRows);
            mb2 = __ldg(mask_in + mrow + 2 * kRows); mb3 = __ldg(mask_in + mrow + 3 * kRows);
            if (l == 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < njh && row < p.M) ds[j] = __ldg(p.dS + row * p.lds + p.j0[h] + j);
            }
          }
          if (store) {
            // the bulk store that read this activation half (previous stage) must be done reading before it is rewritten
            if (leader && !(p.debug & 1)) bulk_wait_read0();
            group_bar(eg);
          }
          if (has_acc) {
            mbar_wait(acc_full0 + 8 * eg, n_acc & 1, 5);
            tc_fence_after();
            ++n_acc;
          }
          float dj[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            float v[32];
            if (has_acc) tmem_ld32(tmem_base + lane_addr + eg * 128 + c * 32, v);
            if (!BWD) {
              // v = pre-activation (the bias was in the accumulator).  The chunk has been read: hand it the next stage's bias.
              if (nbias != nullptr) {
                float nb[32];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(nbias + c * 32 + i4 * 4);
                  nb[i4 * 4 + 0] = b4.x; nb[i4 * 4 + 1] = b4.y; nb[i4 * 4 + 2] = b4.z; nb[i4 * 4 + 3] = b4.w;
                }
                tmem_st32(tmem_base + lane_addr + eg * 128 + c * 32, nb);
              }
              if (mask_out != nullptr && live && !(p.debug & 2)) {
                // relu'(x) as the complement of the sign bit: one funnel shift per element collects the signs (x >= +0 is
                // kept; the layer-by-layer kernels test x > 0 -- they differ for an exact +0.0 only, where the activation
                // itself is 0)
                uint32_t neg = 0;
#pragma unroll
                for (int i = 31; i >= 0; --i) neg = __funnelshift_l(__float_as_uint(v[i]), neg, 1);
                mask_out[mrow + c * kRows] = ~neg;
              }
              if (l == 3) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < njh) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) dj[j] = fmaf(v[i], wd[j * kHid + c * 32 + i], dj[j]);
                  }
              }
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                uint4 o;
                o.x = relu_pack2(v[gq * 8 + 0], v[gq * 8 + 1]); o.y = relu_pack2(v[gq * 8 + 2], v[gq * 8 + 3]);
                o.z = relu_pack2(v[gq * 8 + 4], v[gq * 8 + 5]); o.w = relu_pack2(v[gq * 8 + 6], v[gq * 8 + 7]);
                *reinterpret_cast<uint4*>(act + (size_t)(c * 4 + gq) * 2048) = o;
              }
              continue;
            } else {
              if (l == 0) {  // prologue: dA3 = dS . W_out (the narrow output layers' data gradient)
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < njh) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaf(ds[j], wd[j * kHid + c * 32 + i], v[i]);
                  }
              }
              const uint32_t mb = c == 0 ? mb0 : c == 1 ? mb1 : c == 2 ? mb2 : mb3;  // relu'(activation) from its sign bits
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = ((mb >> i) & 1u) ? v[i] : 0.0f;
            }
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) *reinterpret_cast<uint4*>(act + (size_t)(c * 4 + gq) * 2048) = pack8(v + gq * 8);
          }
          if (!BWD) tmem_st_wait();  // the next stage's bias is in the accumulator before `ready` is signalled
          if (has_acc) tc_fence_before();
          fence_async_smem();   // generic-proxy writes of the activation half -> visible to UMMA / the bulk store
          if (!BWD && l == 3) {  // output layer: column half 1 hands its partial row dots to column half 0
            float* dp = s_dotp + (size_t)t * 4 * kRows + r_local;
            if (g == 1) {
#pragma unroll
              for (int j = 0; j < 4; ++j) dp[j * kRows] = dj[j];
            }
            tile_bar(t);
            if (g == 0 && live && row < p.M) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < njh) {
                  const int jo = p.j0[h] + j;
                  const float r = dj[j] + dp[j * kRows] + s_bout[jo];
                  p.S[row * p.lds + jo] = ((p.act_mask >> jo) & 1u) ? mli_act(r, p.act_out) : r;
                }
            }
            tile_bar(t);  // s_dotp[t] is free again (next head) once every reader is past this point
          }
          group_bar(eg);
          if (leader) {
            if (!(BWD && l == 3)) mbar_arrive(ready0 + 8 * eg);
            if (store && live && !(p.debug & 4)) {
              bulk_s2g(p.A[l] + ((tile * nh + h) * 32 + g * 16) * 1024, sAct + t * 65536 + g * 32768, 32768);
              bulk_commit();
            }
          }
        }
      }
    }
    if (leader) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <bool BWD>
int launch_heads(const HeadsArgs& p, cudaStream_t st) {
  const size_t smem = (size_t)2 * 65536 + (size_t)kSlots * kSlotBytes +
                      (size_t)((BWD ? 0 : 4 * p.nh * kHid) + p.J * kHid + 16 + 2 * 4 * kRows) * sizeof(float);
  MLI_REQUIRE(smem + 1024 <= 232448, "tc_heads: shared memory budget exceeded");
  static size_t smem_set = 0;
  if (smem > smem_set) {
    MLI_CUDA_OK(cudaFuncSetAttribute(tc_heads_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int n_pairs = (p.n_tiles + 1) / 2;
  const int grid = n_pairs < mli_sm_limit() ? n_pairs : mli_sm_limit();
  HeadsArgs q = p;
  if (const char* e = getenv("MLI_HF_DEBUG")) q.debug = atoi(e);
  tc_heads_kernel<BWD><<<grid, kThreads, smem, st>>>(q);
  MLI_LAUNCH_OK();
  return MLI_OK;
}

int fill_heads(HeadsArgs* p, int32_t nh, const int32_t* host_j0, const int32_t* host_nj, int64_t lds, int64_t M) {
  MLI_REQUIRE(M >= kRows && M % kRows == 0, "tc_heads: M must be a positive multiple of 128");
  MLI_REQUIRE(nh >= 1 && nh <= kMaxHeads, "tc_heads: 1..3 heads");
  MLI_REQUIRE(host_j0 && host_nj, "tc_heads: NULL output ranges");
  p->nh = nh;
  int J = 0;
  for (int h = 0; h < nh; ++h) {
    p->j0[h] = host_j0[h]; p->nj[h] = host_nj[h];
    MLI_REQUIRE(p->nj[h] >= 1 && p->nj[h] <= 4 && p->j0[h] >= 0 && p->j0[h] + p->nj[h] <= lds, "tc_heads: bad output range");
    J = p->j0[h] + p->nj[h] > J ? p->j0[h] + p->nj[h] : J;
  }
  MLI_REQUIRE(J <= kMaxJ, "tc_heads: at most 12 outputs");
  p->J = J; p->lds = lds; p->M = M; p->n_tiles = (int)(M / kRows);
  return MLI_OK;
}

}  // namespace

// Fused forward of the whole head stack (see the header of this file).  XH: TCL-128 input of head layer 0 (K0 = 8 *
// k0_chunks <= 8 * xh_chunks columns); W0: TCL with 256-row tiles [nh][K0/8][256][8]; W1..W3: TCL with 128-row tiles
// [nh][2][32][128][8] (both written by mli_weightnorm_pack_batch); bias0..3 [nh * 256]; w_out [J, 256] / b_out [J] fp32 with head h owning outputs
// [host_j0[h], host_j0[h] + host_nj[h]); A0..A3 (bf16 TCL-128, nh * 32 chunks) and mask0..3 (relu sign bits) are written
// only when store_activations != 0 (a backward pass follows); S [M, lds] fp32 = act_j(w_out[j] . A3 + b_out[j]).
extern "C" int mli_tc_heads_fwd(const void* XH, int64_t M, int32_t nh, int32_t K0, int32_t store_activations,
                                int32_t xh_chunks, const void* W0, const void* W1, const void* W2, const void* W3,
                                const float* bias0, const float* bias1, const float* bias2, const float* bias3,
                                const float* w_out, const float* b_out, const int32_t* host_j0, const int32_t* host_nj,
                                int32_t act_out, uint32_t act_mask, void* A0, void* A1, void* A2, void* A3, void* mask0,
                                void* mask1, void* mask2, void* mask3, float* S, int64_t lds, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(K0 >= 16 && K0 % 16 == 0 && K0 / 8 <= xh_chunks, "tc_heads_fwd: K0 must be a multiple of 16 <= 8 * xh_chunks");
  MLI_REQUIRE(XH && W0 && W1 && W2 && W3 && bias0 && bias1 && bias2 && bias3 && w_out && S, "tc_heads_fwd: NULL argument");
  const bool store = store_activations != 0;
  MLI_REQUIRE(!store || (A0 && A1 && A2 && A3), "tc_heads_fwd: store_activations needs A0..A3");
  HeadsArgs p;
  memset(&p, 0, sizeof(p));
  if (int e = fill_heads(&p, nh, host_j0, host_nj, lds, M)) return e;
  p.XH = (const __nv_bfloat16*)XH; p.xh_chunks = xh_chunks; p.k0_chunks = K0 / 8;
  p.W0 = (const __nv_bfloat16*)W0;
  p.Wl[0] = (const __nv_bfloat16*)W1; p.Wl[1] = (const __nv_bfloat16*)W2; p.Wl[2] = (const __nv_bfloat16*)W3;
  p.bias[0] = bias0; p.bias[1] = bias1; p.bias[2] = bias2; p.bias[3] = bias3;
  p.wout = w_out; p.bout = b_out; p.act_mask = act_mask; p.act_out = act_out;
  if (store) {
    p.A[0] = (__nv_bfloat16*)A0; p.A[1] = (__nv_bfloat16*)A1; p.A[2] = (__nv_bfloat16*)A2; p.A[3] = (__nv_bfloat16*)A3;
    p.mask[0] = (uint32_t*)mask0; p.mask[1] = (uint32_t*)mask1; p.mask[2] = (uint32_t*)mask2; p.mask[3] = (uint32_t*)mask3;
  }
  p.S = S;
  return launch_heads<false>(p, (cudaStream_t)stream);
}

// Fused data-gradient chain of the head stack: from dS [M, lds] (gradient w.r.t. the PRE-activation head outputs) and the
// relu sign bits mask0..3 of the forward pass to the pre-activation gradients of all four hidden layers,
//     dZ3 = (dS W_out) * relu'(A3),   dZ_{l-1} = (dZ_l W_l) * relu'(A_{l-1})   (l = 3, 2, 1),
// each written once as bf16 TCL-128 (nh * 32 chunks per tile row: the operands of the weight-gradient GEMMs and of the
// layer-0 data gradient).  W3t, W2t, W1t: the TRANSPOSED hidden-layer weights [nh][2][32][128][8] (rows = input unit).
extern "C" int mli_tc_heads_bwd(const float* dS, int64_t lds, int64_t M, int32_t nh, const void* W3t, const void* W2t,
                                const void* W1t, const float* w_out, const int32_t* host_j0, const int32_t* host_nj,
                                const void* mask0, const void* mask1, const void* mask2, const void* mask3, void* dZ0,
                                void* dZ1, void* dZ2, void* dZ3, void* stream) {
  MLI_ENTRY();
  MLI_REQUIRE(dS && W3t && W2t && W1t && w_out && mask0 && mask1 && mask2 && mask3 && dZ0 && dZ1 && dZ2 && dZ3,
              "tc_heads_bwd: NULL argument");
  HeadsArgs p;
  memset(&p, 0, sizeof(p));
  if (int e = fill_heads(&p, nh, host_j0, host_nj, lds, M)) return e;
  p.dS = dS;
  p.Wl[0] = (const __nv_bfloat16*)W3t; p.Wl[1] = (const __nv_bfloat16*)W2t; p.Wl[2] = (const __nv_bfloat16*)W1t;
  p.wout = w_out;
  p.mask[0] = (uint32_t*)mask0; p.mask[1] = (uint32_t*)mask1; p.mask[2] = (uint32_t*)mask2; p.mask[3] = (uint32_t*)mask3;
  p.A[0] = (__nv_bfloat16*)dZ3; p.A[1] = (__nv_bfloat16*)dZ2; p.A[2] = (__nv_bfloat16*)dZ1; p.A[3] = (__nv_bfloat16*)dZ0;
  return launch_heads<true>(p, (cudaStream_t)stream);
}

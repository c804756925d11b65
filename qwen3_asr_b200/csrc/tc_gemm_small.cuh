// Small-M GEMM (round 2): D[M <= 128, N] = A[M, K] * B[N, K]^T for the calls the unbatched reference makes -- one WebSocket window
// or one SSE chunk per job (src/server.py:79-94): 13 .. 78 tokens.
//
// At such an M the persistent pair kernel (tc_gemm.cuh) is bound by DRAM LATENCY, not bandwidth: the weights do not fit L2
// across a forward (0.37 / 0.64 GB), a 128-column weight tile streams through ONE CTA pair's 6-stage ring (192 KB in flight
// against a ~1.5 us round trip), and with N / 128 tiles only 7 .. 28 pairs work at all: ~0.7 MB in flight where 10 MB would
// saturate the HBM.  This kernel turns the tiling around: N is cut into 32-column slices, one CTA per slice (28 .. 128 CTAs),
// and a CTA requests its WHOLE weight slice (32 x K bf16 <= 64 KB for K <= 1024, every k-block on its own mbarrier) up
// front -- BEFORE griddepcontrol.wait: the weights do not depend on the previous kernel, so under programmatic dependent
// launch the weight stream of GEMM i+1 overlaps the tail of kernel i.  Only the activation k-blocks (L2-resident) go through a
// ring.  The K loop is the same sequence of 128 x BN x 16 tcgen05.mma accumulating into one TMEM tile in ascending k, and the
// epilogue is the same code (epilogue_panel) with the same functors, so the result is BIT-IDENTICAL to the pair kernel's: a
// window alone still equals the window inside a batch (tests/test_gpu_graph.py::test_small_m_path_is_bit_identical).
#pragma once

#include "tc_gemm.cuh"

namespace qasr {
namespace tc {

constexpr int kSmallThreads = 256;      // warp 0 TMA, warp 1 MMA, warp 2 TMEM, warp 3 idle, warps 4..7 epilogue (one per lane quarter)
constexpr int kSmallAStages = 4;
constexpr int kSmallMaxKb = 16;         // K <= 1024

template <int BN>
struct SmallLayout {
  static constexpr int B_TILE_BYTES = BN * BLOCK_K_BYTES;              // one k-block of the slice: 2 / 4 KB (1024-byte multiple)
  static constexpr int A_OFFSET(int num_kb) { return num_kb * B_TILE_BYTES; }
  static constexpr int STAGING_OFFSET(int num_kb) { return A_OFFSET(num_kb) + kSmallAStages * A_STAGE_BYTES; }
  static constexpr int TAB_OFFSET(int num_kb) { return STAGING_OFFSET(num_kb) + 4 * kStagingBytes; }   // float bias[32], float colscale[32]
  static constexpr int BAR_OFFSET(int num_kb) { return TAB_OFFSET(num_kb) + 2 * 32 * 4; }
  static constexpr int NUM_BARS = kSmallMaxKb + 2 * kSmallAStages + 1;
  static constexpr int total(int num_kb) { return BAR_OFFSET(num_kb) + NUM_BARS * 8 + 16; }
  static_assert(B_TILE_BYTES % 1024 == 0, "weight k-block tiles keep the 1024-byte alignment of the 128B swizzle atom");
};

template <int BN, class Epi>
__global__ void __launch_bounds__(kSmallThreads, 1)
gemm_smallm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int num_kb, Epi epi) {
  using L = SmallLayout<BN>;
  static_assert(BN == 32, "one 32-column epilogue panel per CTA: the granularity at which the residual epilogues round their LayerNorm sums");
  static_assert(!Epi::kScaled, "bf16 only");
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("qasr tc_gemm_small: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* sb = smem;                                   // [num_kb][BN rows][128 B]
  uint8_t* sa = smem + L::A_OFFSET(num_kb);             // [kSmallAStages][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET(num_kb));
  uint64_t* full_b = bars;                              // [kSmallMaxKb]
  uint64_t* full_a = bars + kSmallMaxKb;                // [kSmallAStages]
  uint64_t* empty_a = full_a + kSmallAStages;           // [kSmallAStages]
  uint64_t* tmem_full = empty_a + kSmallAStages;        // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < num_kb; ++i) mbar_init(&full_b[i], 1);
    for (int i = 0; i < kSmallAStages; ++i) {
      mbar_init(&full_a[i], 1);
      mbar_init(&empty_a[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    // the whole weight slice, requested before the previous kernel of the chain has finished (weights do not depend on it)
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_arrive_expect_tx(&full_b[kb], L::B_TILE_BYTES);
      tma_load_2d(sb + kb * L::B_TILE_BYTES, &tmB, kb * BLOCK_K, n_blk * BN, &full_b[kb]);
    }
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // the activation (and the row statistics the epilogue reads) come from the previous kernel

  if (warp == 0) {
    // ===================== TMA producer: activation k-blocks through a 4-stage ring =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      wait_small_empty(&empty_a[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_a[stage], A_STAGE_BYTES);
        tma_load_2d(sa + stage * A_STAGE_BYTES, &tmA, kb * BLOCK_K, 0, &full_a[stage]);
      }
      __syncwarp();
      if (++stage == kSmallAStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN);
    const uint64_t desc_a0 = make_smem_desc_sw128(smem_u32(sa));
    const uint64_t desc_b0 = make_smem_desc_sw128(smem_u32(sb));
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      wait_small_full(&full_b[kb], 0);
      wait_small_full(&full_a[stage], phase);
      tc_fence_after();
      const uint64_t da = desc_a0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
      const uint64_t db = desc_b0 + static_cast<uint64_t>((kb * L::B_TILE_BYTES) >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BLOCK_K_BYTES / 32; ++k)
          umma<K_BF16>(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_a[stage]);
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == kSmallAStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: warp = TMEM lane quarter, ONE panel of BN columns =====================
    const int q = warp & 3;
    const int et = threadIdx.x - 4 * 32;
    float* bias_s = reinterpret_cast<float*>(smem + L::TAB_OFFSET(num_kb));
    float* scale_s = bias_s + 32;
    if (et < 32) {
      const float* bp = epi.bias_ptr();
      const int n = n_blk * BN + et;
      bias_s[et] = (et < BN && bp != nullptr && n < epi.n_cols()) ? __ldg(bp + n) : 0.f;
      if constexpr (Epi::kLnFold) scale_s[et] = (et < BN && n < epi.n_cols()) ? __ldg(epi.col_scale_ptr() + n) : 0.f;
    }
    named_bar_sync(1, 128);
    const int row0 = q * 32;
    const bool live = epi.row_live(row0 + lane);
    float2 ln = make_float2(0.f, 1.f);
    if constexpr (Epi::kLnFold) {
      static_assert(!Epi::kLnPart, "the small-M path reads finished row statistics");
      ln = epi.row_stats(row0 + lane);
      ln.x *= ln.y;
    }
    wait_small_full(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t stage_base = smem_u32(smem + L::STAGING_OFFSET(num_kb) + q * kStagingBytes);
    epilogue_panel(epi, taddr, BN, row0, n_blk * BN, live, 1.f, ln, bias_s, scale_s, stage_base, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

}  // namespace tc
}  // namespace qasr

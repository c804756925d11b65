// tcgen05 / TMEM / TMA GEMM for sm_100a: D[M,N] = A[M,K] * B[N,K]^T, bf16 in, fp32 accumulate in TMEM.
//
// One persistent, warp-specialised kernel serves every matrix product of the audio encoder:
//   * the Linear layers (K7, K9-K12 of SURVEY.md section 2.3): A is a row-major [M,K] activation,
//   * the 3x3 stride-2 convolutions as implicit GEMM (K3, K4): A is gathered by TMA straight from
//     the NHWC-like activation (a 3-D box with traversal stride 2 per filter tap, out-of-bounds
//     rows zero-filled by the TMA unit), K = 9 taps x 512 (480 channels padded),
//   * conv_out (K5): A is conv3's output viewed as [chunk*13, 16*480].
// B is always an nn.Linear-style [N,K] row-major (K-major) weight.
//
// Roles (640 threads, 1 CTA per SM, grid = the number of co-resident CTA pairs x 2):
//   warp 0 lane 0 : TMA producer   -- cp.async.bulk.tensor into a kStages-deep 128B-swizzled ring
//   warp 1 lane 0 : MMA issuer     -- tcgen05.mma (cta_group::2: M=256 over the pair, leader CTA only), N=BN, K=16 x4 per stage
//   warp 2        : TMEM allocator -- 512 columns = two accumulator stages
//   warps 2, 3    : (LnFoldPart epilogues only) reduce the LayerNorm partial sums of the tile's rows into a smem table
//   warps 4..19   : epilogue       -- four warps per TMEM lane quarter (a GELU epilogue on 8 warps took longer than the
//                                     K = 1024 main loop of its tile), 32-column panels: tcgen05.ld -> bias (from
//                                     smem) / activation -> bf16 -> swizzled smem staging -> coalesced 64-byte row
//                                     segments to global (+ residual / positional addend, prefetched)
// Pipelines: smem full/empty (TMA <-> MMA) and TMEM full/empty (MMA <-> epilogue) mbarriers, so the
// epilogue of tile i overlaps the MMAs of tile i+1.  The kernel takes part in programmatic dependent launch (common.cuh):
// its prologue runs while the previous kernel of the chain drains.
//
// CTA2 = true (the product configuration): the two CTAs of a 2-CTA cluster (one TPC) work on one 256 x BN tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 rows of A and only HALF of the B tile (BN / 2 weight rows); the
// tensor cores of both SMs read the two halves, so the shared-memory bytes written and read per FLOP drop by a third
// and a stage is 32 KB instead of 48 KB (6-deep ring).  Only the leader CTA (cluster rank 0) issues MMAs; every TMA of
// the pair signals the leader's full barrier, tcgen05.commit multicasts "stage free" / "accumulator ready" to both
// CTAs, and both CTAs' epilogue warps release the accumulator on the leader's TMEM-empty barrier.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace qasr {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16: 64 elements = 128 bytes = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int BLOCK_K_BYTES = 128;  // every operand kind stages 128-byte rows; one UMMA consumes 32 of them
enum Kind : int { K_BF16 = 0, K_E4M3 = 1 };  // kind::f16 (bf16 x bf16) / kind::f8f6f4 (e4m3 x e4m3), fp32 accumulate
template <int KIND>
constexpr int block_k_elems() { return KIND == K_E4M3 ? 128 : 64; }
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;  // 640
constexpr int kPanel = 32;            // epilogue panel: 32 bf16 columns = one 64-byte row segment
constexpr int kStagingBytes = 32 * kPanel * 2;  // per epilogue warp
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns per accumulator stage (2 stages)
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

enum AMode : int { A_LINEAR = 0, A_CONV = 1 };

struct GemmShape {
  int m_tiles;      // number of 128-row tiles
  int n_tiles;      // number of BN-column tiles
  int num_kb;       // K / 64
  // A_CONV only: K index kb -> (tap = kb / kb_per_tap, channel block = kb % kb_per_tap);
  // output tile m_blk covers gt global output columns x (128 / gt) output rows.
  int kb_per_tap;
  int gt;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> a CUDA error the host reports) instead of hanging the GPU.
// One copy per waiting role, each on its own source line, so profiler samples attribute to the role.
#define QASR_DEFINE_MBAR_WAIT(name)                                                                      \
  __device__ __forceinline__ void name(uint64_t* bar, uint32_t parity) {                                 \
    if (mbar_try_wait(bar, parity)) return;                                                              \
    const long long t0 = clock64();                                                                      \
    while (!mbar_try_wait(bar, parity)) {                                                                \
      if (clock64() - t0 > 8000000000LL) { /* ~4 s at 2 GHz */                                           \
        printf("qasr tc_gemm: " #name " timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);     \
        __trap();                                                                                        \
      }                                                                                                  \
    }                                                                                                    \
  }
QASR_DEFINE_MBAR_WAIT(wait_smem_empty)  // TMA producer: stage freed by the MMAs that read it
QASR_DEFINE_MBAR_WAIT(wait_smem_full)   // MMA issuer: TMA bytes of the stage have landed
QASR_DEFINE_MBAR_WAIT(wait_tmem_empty)  // MMA issuer: epilogue has drained the accumulator stage
QASR_DEFINE_MBAR_WAIT(wait_tmem_full)   // epilogue: the tile's last MMA has completed
QASR_DEFINE_MBAR_WAIT(wait_stats_full)  // epilogue: the tile's row statistics are in the smem table (LnFoldPart)
QASR_DEFINE_MBAR_WAIT(wait_stats_empty) // statistics warps: the table slot has been read
QASR_DEFINE_MBAR_WAIT(wait_small_empty) // small-M kernel, TMA producer: activation stage freed
QASR_DEFINE_MBAR_WAIT(wait_small_full)  // small-M kernel: weight k-block / activation stage / accumulator has landed
#undef QASR_DEFINE_MBAR_WAIT

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// ---- 2-CTA (cta_group::2) variants ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of a peer CTA.  Default semantics (as CUTLASS's ClusterBarrier::arrive): a .release.cluster
// arrive made every epilogue warp drain its global stores first (26 % of the GELU epilogue's stall samples); the only
// ordering this arrive has to carry -- TMEM reads before the next tile's MMAs -- comes from tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA (the pair's leader)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster_addr)
      : "memory");
}

// One lane of a converged warp (elect.sync): ptxas then knows the guarded region runs on exactly one lane and issues
// the uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR) directly instead of serialising over "active lanes".
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == K_E4M3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// cta_group::2: M = 256 over the CTA pair; issued by ONE thread of the leader CTA.
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (KIND == K_E4M3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive on the barrier at this smem offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major, dense.
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major=0 (K)  [16] b_major=0 (K)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// kind::f8f6f4 with both operands E4M3 (a_format = b_format = 0), fp32 accumulate, K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc_e4m3(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// Shared-memory matrix descriptor for a K-major tile stored as rows of 128 bytes with SWIZZLE_128B
// (exactly what a TMA box with inner extent 64 bf16 and CU_TENSOR_MAP_SWIZZLE_128B writes):
//   [0,14) start>>4  [16,30) LBO>>4 (=1, unused for swizzled K-major)  [32,46) SBO>>4 (8 rows*128B = 1024)
//   [46,48) version=1 (sm_100)  [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// SCALED (fp8 kinds): bias and column-scale tables are single-buffered (an extra epilogue barrier per tile keeps a fast
// warp from overwriting them) -- with 4 stages of 48 KB and 32 KB of staging there is no room for two copies of both.
template <int BN, int STAGES, bool SCALED = false, bool CTA2 = false, bool LNFOLD = false>
struct SmemLayout {
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;   // weight rows staged by this CTA (the pair's other CTA stages the rest)
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;               // [kEpiWarps][32 rows][128 B]
  static constexpr int TAB_BUFS = SCALED ? 1 : 2;
  static constexpr int BIAS_OFFSET = STAGING_OFFSET + kEpiWarps * kStagingBytes;  // float [TAB_BUFS][256]
  static constexpr int SCALE_OFFSET = BIAS_OFFSET + TAB_BUFS * 256 * 4;           // float [TAB_BUFS][256] column scales (fp8) / column sums (LnFold)
  static constexpr int STAT_OFFSET = SCALE_OFFSET + ((SCALED || LNFOLD) ? TAB_BUFS * 256 * 4 : 0);  // float2 [2][128] row statistics (LnFoldPart)
  static constexpr int BAR_OFFSET = STAT_OFFSET + (LNFOLD ? 2 * BLOCK_M * 8 : 0);
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 4;  // + stats_full[2], stats_empty[2] (LnFoldPart)
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");
  static_assert(B_STAGE_BYTES % 1024 == 0, "B stage must keep 1024-byte alignment");
};

// pair mode: six 32 KB stages at BN = 256; a 64-column tile stages 20 KB, and nine of them (the K = 3584 / 4096 fc2 of a single
// window runs on such tiles and is bound by the latency of its ring, see gemm.cu) still fit
template <int BN, bool CTA2 = false>
constexpr int default_stages() { return CTA2 ? (BN <= 64 ? 9 : 6) : (BN > 192 ? 4 : (BN > 128 ? 4 : 6)); }

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------------
// One 32-column panel of the epilogue, executed by ONE warp (lane = accumulator row in phase A): shared by the persistent
// pair kernel below and the small-M kernel (tc_gemm_small.cuh).
//   taddr_c0   TMEM address of the panel's first column in this warp's lane quarter
//   ncol0      global output column of the panel's first column;  ncols <= 32 valid columns
//   row0       global output row of lane 0;  bs_tab / ss_tab: the panel's bias / per-column table entries in shared memory
// ---------------------------------------------------------------------------------------------
template <class Epi>
__device__ __forceinline__ void epilogue_panel(const Epi& epi, uint32_t taddr_c0, int ncols, int row0, int ncol0, bool live, float row_scale,
                                               float2 ln, const float* bs_tab, const float* ss_tab, uint32_t stage_base, int lane) {
  const int sub = lane >> 2, ch = lane & 3;  // phase B: row within an 8-row group, 16-byte chunk of the 64-byte row segment
    // destinations of phase B + prefetch of the post-rounding addend (hidden behind phase A)
    long long offs[4];
    typename Epi::Prefetch pre[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = row0 + i * 8 + sub;
      const int n = ncol0 + ch * 8;
      offs[i] = ch * 8 < ncols ? epi.offset(m, n) : -1;
      if (offs[i] >= 0) pre[i] = epi.prefetch(m, n, offs[i]);
    }
    // phase A: TMEM -> registers -> bias / activation -> bf16 -> staging (row = lane, 64-byte rows whose 16-byte
    // chunks are XOR-swizzled with (row >> 1) & 3: conflict-free for the row-per-lane writes and the phase-B reads)
    uint32_t v[kPanel];
#pragma unroll
    for (int g = 0; g < kPanel / 16; ++g)
      if (g * 16 < ncols) tmem_ld16(taddr_c0 + g * 16, v + g * 16);
    tmem_ld_wait();
    const float* bs = bs_tab;
    const float* ss = ss_tab;
#pragma unroll
    for (int j = 0; j < kPanel / 8; ++j) {
      if (j * 8 < ncols) {
        const float4 b0 = *reinterpret_cast<const float4*>(bs + j * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bs + j * 8 + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = __uint_as_float(v[j * 8 + e]);
        if constexpr (Epi::kScaled) {  // dequantise: acc * (activation row scale * weight column scale)
          const float4 s0 = *reinterpret_cast<const float4*>(ss + j * 8);
          const float4 s1 = *reinterpret_cast<const float4*>(ss + j * 8 + 4);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] *= row_scale * sc[e];
        }
        float pre_act[8];  // the module output before its bf16 rounding
        if constexpr (Epi::kLnFold) {  // rstd * acc + (bias' - rstd * mean * colsum[n]): two FMAs per output
          const float4 s0 = *reinterpret_cast<const float4*>(ss + j * 8);
          const float4 s1 = *reinterpret_cast<const float4*>(ss + j * 8 + 4);
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) pre_act[e] = fmaf(acc[e], ln.y, fmaf(-ln.x, sc[e], bb[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) pre_act[e] = acc[e] + bb[e];
        }
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          const float2 r = bf16_round2(pre_act[e], pre_act[e + 1]);
          o[e] = live ? epi.act(r.x) : 0.f;
          o[e + 1] = live ? epi.act(r.y) : 0.f;
        }
        uint4 pk;
        pk.x = pack_bf16x2(o[0], o[1]);
        pk.y = pack_bf16x2(o[2], o[3]);
        pk.z = pack_bf16x2(o[4], o[5]);
        pk.w = pack_bf16x2(o[6], o[7]);
        st_shared_v4(stage_base + lane * (kPanel * 2) + ((j ^ ((lane >> 1) & 3)) << 4), pk);
      }
    }
    __syncwarp();
    // phase B: each instruction moves eight complete 64-byte row segments
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (!Epi::kRowStats) {
        if (offs[i] >= 0) {
          const int r = i * 8 + sub;
          const uint4 sv = ld_shared_v4(stage_base + r * (kPanel * 2) + ((ch ^ ((r >> 1) & 3)) << 4));
          epi.finish(offs[i], sv, pre[i]);
        }
      } else {
        // ... and leaves the row's partial LayerNorm statistics of this 32-column panel (sum, sum of squares of the stored bf16
        // values): the four lanes of a row add up through two shuffles, lane ch == 0 writes the fixed slot part[row][panel]
        const int r = i * 8 + sub;
        float s1 = 0.f, s2 = 0.f;
        if (offs[i] >= 0) {
          const uint4 sv = ld_shared_v4(stage_base + r * (kPanel * 2) + ((ch ^ ((r >> 1) & 3)) << 4));
          const uint4 fin = epi.finish(offs[i], sv, pre[i]);
          const float2 a = unpack_bf16x2(fin.x), b = unpack_bf16x2(fin.y), c = unpack_bf16x2(fin.z), d2 = unpack_bf16x2(fin.w);
          s1 = ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (d2.x + d2.y));
          s2 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(c.x, c.x, fmaf(c.y, c.y, fmaf(d2.x, d2.x, d2.y * d2.y)))))));
        }
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
        if (ch == 0) {
          const int srow = epi.stats_row(row0 + r);
          if constexpr (Epi::kRowAtomic) {
            if (srow >= 0) {
              atomicAdd(epi.acc + 2 * static_cast<long long>(srow), static_cast<unsigned long long>(__float2ll_rn(s1 * kStatSumScale)));
              atomicAdd(epi.acc + 2 * static_cast<long long>(srow) + 1, static_cast<unsigned long long>(__float2ll_rn(s2 * kStatSqScale)));
            }
          } else {
            if (srow >= 0) epi.part[static_cast<long long>(srow) * epi.n_panels + ((ncol0) >> 5)] = make_float2(s1, s2);
          }
        }
      }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// The kernel.  The epilogue functor (epilogues.cuh) supplies:
//   bias_ptr()            per-column fp32 vector added to the accumulator (or nullptr), n_cols() its length
//   row_live(m)           false -> the whole output row is written as zeros (conv padding columns)
//   act(v)                applied to bf16(acc + bias)
//   offset(m, n)          element offset of 8 consecutive output columns, or -1 to skip them
//   prefetch / finish     optional post-rounding addend (residual, positional table) and the final store
// ---------------------------------------------------------------------------------------------
template <int BN, int STAGES, int AMODE, int KIND, bool CTA2, class Epi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmShape shape, Epi epi) {
  using L = SmemLayout<BN, STAGES, Epi::kScaled, CTA2, Epi::kLnFold>;
  static_assert(!CTA2 || BN % 16 == 0, "UMMA N for M=256 must be a multiple of 16; each CTA stages BN / 2 rows (whole 8-row swizzle atoms)");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
  static_assert(BN <= kAccStride, "an accumulator stage is kAccStride TMEM columns");
  static_assert(KIND == K_BF16 || AMODE == A_LINEAR, "the implicit-GEMM convolutions stay bf16 (torchao quantises nn.Linear only)");
  constexpr int KB_ELEMS = block_k_elems<KIND>();

  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {  // SWIZZLE_128B tiles need 1024-byte aligned stage bases
    if (threadIdx.x == 0) printf("qasr tc_gemm: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* full_bar = bars;                      // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES] MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;    // [2] MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
  uint64_t* stats_full_bar = bars + 2 * STAGES + 4;  // [2] statistics warps -> epilogue (LnFoldPart)
  uint64_t* stats_empty_bar = bars + 2 * STAGES + 6; // [2] epilogue -> statistics warps
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + L::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA2: a "tile" is 256 rows (two m blocks, one per CTA of the pair) and the pair walks the tile list together
  const int rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int num_tiles = (CTA2 ? (shape.m_tiles + 1) / 2 : shape.m_tiles) * shape.n_tiles;
  const int tile0 = CTA2 ? blockIdx.x >> 1 : blockIdx.x;
  const int tile_step = CTA2 ? gridDim.x >> 1 : gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], (CTA2 ? 2 : 1) * kEpiWarps);  // one elected lane of each epilogue warp (of both CTAs)
      mbar_init(&stats_full_bar[i], 2);
      mbar_init(&stats_empty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) tmem_alloc_pair(tmem_ptr_smem, kTmemCols);
    else tmem_alloc(tmem_ptr_smem, kTmemCols);
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above overlapped the previous kernel's tail; its results are needed from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = CTA2 ? 2 * (tile / shape.n_tiles) + rank : tile / shape.n_tiles;
      const int n_blk = tile % shape.n_tiles;
      int tap = 0, cb = 0;
      for (int kb = 0; kb < shape.num_kb; ++kb) {
        wait_smem_empty(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if constexpr (CTA2) {
            // both CTAs' bytes are counted on the LEADER's full barrier (a row block past the end of A is zero-filled
            // by the TMA unit and still counts its full box)
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            if constexpr (AMODE == A_LINEAR) {
              tma_load_2d_pair(sa, &tmA, kb * KB_ELEMS, m_blk * BLOCK_M, bar);
            } else {
              const int kh = tap / 3, kw = tap - kh * 3;
              tma_load_3d_pair(sa, &tmA, cb * KB_ELEMS, kh - 1, 2 * (m_blk * shape.gt) + kw, bar);
            }
            tma_load_2d_pair(sb, &tmB, kb * KB_ELEMS, n_blk * BN + rank * L::B_ROWS, bar);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            if constexpr (AMODE == A_LINEAR) {
              tma_load_2d(sa, &tmA, kb * KB_ELEMS, m_blk * BLOCK_M, &full_bar[stage]);
            } else {
              const int kh = tap / 3, kw = tap - kh * 3;
              // input column = 2 * (global output column) + kw (left zero column is part of the layout),
              // input row    = 2 * (output row) + kh - 1    (row -1 is out of bounds -> TMA zero fill)
              tma_load_3d(sa, &tmA, cb * KB_ELEMS, kh - 1, 2 * (m_blk * shape.gt) + kw, &full_bar[stage]);
            }
            tma_load_2d(sb, &tmB, kb * KB_ELEMS, n_blk * BN, &full_bar[stage]);
          }
        }
        __syncwarp();
        if (++cb == shape.kb_per_tap) { cb = 0; ++tap; }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (CTA2: the leader CTA only) =====================
    // The whole warp walks the pipeline (converged waits); one elected lane issues.  Descriptor bases of the ring's
    // stages are loop invariants; the per-k advance (one UMMA = 32 bytes of K inside the 128B swizzle atom: 16 bf16 or
    // 32 e4m3) is +2 in the >>4 address field.
    constexpr int UM = CTA2 ? 2 * BLOCK_M : BLOCK_M;
    constexpr uint32_t idesc = KIND == K_E4M3 ? make_idesc_e4m3(UM, BN) : make_idesc_bf16(UM, BN);
    const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem));
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++iter) {
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      wait_tmem_empty(&tmem_empty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kAccStride);
      for (int kb = 0; kb < shape.num_kb; ++kb) {
        wait_smem_full(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t da = desc0 + static_cast<uint64_t>((stage * L::STAGE_BYTES) >> 4);
        const uint64_t db = da + static_cast<uint64_t>(A_STAGE_BYTES >> 4);
        // implicit-GEMM convolutions: a tap's 480 channels are 7.5 k-blocks; the last block of every tap holds 32 real channels and
        // 32 zero-padded ones -- its second half is skipped (two of four MMAs), 6.25 % of the convolution's tensor work
        const bool half_block = AMODE == A_CONV && (kb % shape.kb_per_tap) == shape.kb_per_tap - 1;
        if (elect_one()) {
          if constexpr (CTA2) {
#pragma unroll
            for (int k = 0; k < BLOCK_K_BYTES / 32; ++k)
              if (!(half_block && k >= BLOCK_K_BYTES / 64))
              umma_pair<KIND>(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage]);
            if (kb == shape.num_kb - 1) umma_commit_pair(&tmem_full_bar[as]);
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K_BYTES / 32; ++k)
              umma<KIND>(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == shape.num_kb - 1) umma_commit(&tmem_full_bar[as]);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===================== row statistics (warps 2, 3; LnFoldPart only) =====================
    if constexpr (Epi::kLnFold) {
      if constexpr (Epi::kLnPart) if (warp >= 2) {
        float2* stat_s = reinterpret_cast<float2*>(smem + L::STAT_OFFSET);
        const int sw = warp - 2;  // rows sw * 64 .. + 63 of the CTA's 128
        int iter = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++iter) {
          const int m_blk = CTA2 ? 2 * (tile / shape.n_tiles) + rank : tile / shape.n_tiles;
          const int as = iter & 1;
          wait_stats_empty(&stats_empty_bar[as], ((iter >> 1) & 1) ^ 1);
          // lane = row (two rows per lane): 2 x n_panels / 2 independent 16-byte loads per lane in flight, sums in registers, fixed
          // order.  (A warp-per-row butterfly is a chain of dependent L2 round trips -- it took longer than the tile's MMAs.)
          const int row0 = m_blk * BLOCK_M + sw * 64;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int row = row0 + hh * 32 + lane;
            float s1 = 0.f, s2 = 0.f;
            if (row < epi.m_valid) {
              const float4* p = reinterpret_cast<const float4*>(epi.part_p + static_cast<long long>(row) * epi.n_panels);
#pragma unroll 8
              for (int i = 0; i < epi.n_panels / 2; ++i) {
                const float4 v = __ldg(p + i);
                s1 += v.x; s2 += v.y;
                s1 += v.z; s2 += v.w;
              }
            }
            const float mean = s1 * epi.inv_d;
            const float var = fmaxf(fmaf(-mean, mean, s2 * epi.inv_d), 0.f);
            const float rstd = rsqrtf(var + epi.eps);
            stat_s[as * BLOCK_M + sw * 64 + hh * 32 + lane] = make_float2(mean * rstd, rstd);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&stats_full_bar[as]);  // release: the table writes above are visible to the waiting epilogue warps
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - kEpiWarp0;
    const int q = ew & 3;       // == warp % 4: the TMEM lane quarter this warp may read
    const int pw = ew >> 2;     // the four warps of a quarter take every fourth 32-column panel
    const int et = threadIdx.x - kEpiWarp0 * 32;
    const uint32_t stage_base = smem_u32(smem + L::STAGING_OFFSET + ew * kStagingBytes);
    float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
    float* scale_s = reinterpret_cast<float*>(smem + L::SCALE_OFFSET);
    constexpr int NP = (BN + kPanel - 1) / kPanel;
    int iter = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++iter) {
      const int m_blk = CTA2 ? 2 * (tile / shape.n_tiles) + rank : tile / shape.n_tiles;
      const int n_blk = tile % shape.n_tiles;
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      const int tab = as % L::TAB_BUFS;
      {  // this tile's bias columns -> smem (overlaps the tile's MMAs)
        const float* bp = epi.bias_ptr();
        const int n = n_blk * BN + et;
        float b = 0.f;
        if (et < BN && bp != nullptr && n < epi.n_cols()) b = __ldg(bp + n);
        if (et < 256) bias_s[tab * 256 + et] = b;
        if constexpr (Epi::kScaled || Epi::kLnFold) {  // per-column table: fp8 weight scales / LayerNorm-fold column sums
          float cs = 0.f;
          if (et < BN && n < epi.n_cols()) cs = __ldg(epi.col_scale_ptr() + n);
          if (et < 256) scale_s[tab * 256 + et] = cs;
        }
      }
      named_bar_sync(1, kEpiThreads);
      wait_tmem_full(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const int row0 = m_blk * BLOCK_M + q * 32;
      const bool live = epi.row_live(row0 + lane);
      float row_scale = 1.f;
      if constexpr (Epi::kScaled) row_scale = epi.row_scale(row0 + lane);
      float2 ln = make_float2(0.f, 1.f);  // (rstd * mean, rstd) of this thread's row
      if constexpr (Epi::kLnFold) {
        if constexpr (Epi::kLnPart) {
          wait_stats_full(&stats_full_bar[as], aphase);
          ln = reinterpret_cast<const float2*>(smem + L::STAT_OFFSET)[as * BLOCK_M + q * 32 + lane];
          __syncwarp();
          if (lane == 0) mbar_arrive(&stats_empty_bar[as]);  // the row's statistics are in registers: the table slot may be refilled
        } else {
          ln = epi.row_stats(row0 + lane);
          ln.x *= ln.y;
        }
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * kAccStride);
#pragma unroll 1
      for (int p = pw; p < NP; p += kEpiWarps / 4) {
        const int c0 = p * kPanel;
        const int ncols = BN - c0 < kPanel ? BN - c0 : kPanel;
        epilogue_panel(epi, taddr + c0, ncols, row0, n_blk * BN + c0, live, row_scale, ln, bias_s + tab * 256 + c0, scale_s + tab * 256 + c0, stage_base,
                       lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));  // the leader's MMA warp waits for both CTAs
        else mbar_arrive(&tmem_empty_bar[as]);
      }
      if constexpr (L::TAB_BUFS == 1) named_bar_sync(2, kEpiThreads);  // everyone is done with the single-buffered tables
    }
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // neither CTA may free TMEM or exit while the pair's MMAs / remote arrivals are in flight
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Debug-only SIMT evaluation of the same product with the same epilogue functor (bring-up aid:
// lets the tests tell a TMA/UMMA descriptor bug from an epilogue or layout bug).  Never used by
// the product path unless QASR_DEBUG_SIMT=1 is set.
// ---------------------------------------------------------------------------------------------
struct ALoadLinear {
  const __nv_bfloat16* a;
  long long lda;
  int m_valid;
  __device__ float operator()(int m, int k) const { return m < m_valid ? __bfloat162float(a[m * lda + k]) : 0.f; }
};
struct ALoadConv {
  const __nv_bfloat16* a;  // [G_in][H_in][C]
  int g_in, h_in, c, hc;   // hc = output rows per output column
  __device__ float operator()(int m, int k) const {
    const int tap = k / 512, ch = k % 512;
    if (ch >= c) return 0.f;
    const int kh = tap / 3, kw = tap % 3;
    const int g = m / hc, h = m % hc;
    const int col = 2 * g + kw, row = 2 * h + kh - 1;
    if (col < 0 || col >= g_in || row < 0 || row >= h_in) return 0.f;
    return __bfloat162float(a[(static_cast<long long>(col) * h_in + row) * c + ch]);
  }
};

template <class ALoad, class Epi>
__global__ void gemm_simt_kernel(ALoad aload, const __nv_bfloat16* __restrict__ b, long long ldb, int m_rows, int n, int k, Epi epi) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int n_groups = n / 8;
  const int m = static_cast<int>(idx / n_groups);
  const int n0 = static_cast<int>(idx % n_groups) * 8;
  if (m >= m_rows) return;
  const long long off = epi.offset(m, n0);
  if (off < 0) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int kk = 0; kk < k; ++kk) {
    const float av = aload(m, kk);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(av, __bfloat162float(b[(n0 + j) * ldb + kk]), acc[j]);
  }
  const float* bp = epi.bias_ptr();
  const bool live = epi.row_live(m);
  uint32_t pk[4];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    const float2 r = bf16_round2(acc[j] + (bp != nullptr ? bp[n0 + j] : 0.f), acc[j + 1] + (bp != nullptr ? bp[n0 + j + 1] : 0.f));
    const float y0 = live ? epi.act(r.x) : 0.f;
    const float y1 = live ? epi.act(r.y) : 0.f;
    pk[j / 2] = pack_bf16x2(y0, y1);
  }
  const typename Epi::Prefetch pre = epi.prefetch(m, n0, off);
  epi.finish(off, make_uint4(pk[0], pk[1], pk[2], pk[3]), pre);
}

}  // namespace tc
}  // namespace qasr

// Windowed (block-diagonal, non-causal) multi-head attention on the tcgen05 tensor cores.
//
// Restates Qwen3OmniMoeAudioAttention (transformers modeling_qwen3_omni_moe.py:496-565, eager definition :471-493) under the
// per-window cu_seqlens of :745-752 -- what the reference deployment gets from flash-attn varlen: for every window (<= 104
// tokens with n_window_infer = 800; up to 128 supported here) and head, softmax(q k^T / sqrt(64)) v with fp32 softmax and the
// un-normalised probabilities rounded to bf16 for the second product (flash-attn's rounding point).
//
// One persistent CTA per SM walks (window, head) items.  Per item everything stays on chip:
//   warp 0        TMA: Q, K, V tiles [128 tokens x 64 dims] of the head-major qkv activation (one contiguous 16 KB block each)
//                 -> 128B-swizzled smem (4-stage ring)
//   warp 1        MMA issuer: S = Q K^T (kind::f16, M = 128, N = keys rounded to 16, K = 64) into TMEM;
//                 O = P V with P read from TMEM (A operand) and V used as an MN-major B operand straight from its
//                 row-major tile -- no transposes, no smem round trip for S or P
//   warp 3        clip isolation: P V runs over keys rounded up to 16, and the V tile's rows past the window's end belong to the
//                 next window (another clip) or are stale; a NaN / Inf there would turn 0 * x into NaN.  This warp overwrites
//                 those (<= 15) rows of the landed V tile with zeros before P V may read it (v_ready barrier)
//   warp 2        TMEM allocation: 4 slots x 128 columns.  A slot holds S (<= 128 fp32 columns); P (bf16 pairs) overwrites its
//                 first 64 columns as the softmax proceeds, and O (64 fp32 columns) is accumulated into columns 64..127 --
//                 the upper half of S, dead once the whole row has been read -- so four items fit where two did
//   warps 4-11    two softmax warpgroups (thread = query row), item i served by warpgroup i % 2 in slot i % 4: ONE tcgen05.ld sweep
//                 of the S row into registers -> mask -> max -> exp2 -> row sum -> bf16 P back into the S columns (tcgen05.st) ->
//                 after the second MMA, O / sum -> bf16 -> HBM
// An item's softmax is a latency chain (TMEM round trips, 100+ dependent MUFU / FMNMX per row), so throughput comes from the
// number of items in flight: four warpgroups, QK^T issued two items ahead of PV, and a 5-stage TMA ring (tiles of 112 rows
// when every window fits, as with the checkpoints' 104-token windows).
#include <algorithm>

#include "kernels.h"
#include "tc_gemm.cuh"

namespace qasr {
namespace {

using namespace tc;

constexpr int AT_SLOTS = 4;                       // TMEM slots: Q K^T runs up to four items ahead
constexpr int AT_WGS = 2;                         // softmax warpgroups: few enough threads (384) that a whole S row fits in registers
constexpr int AT_THREADS = (4 + 4 * AT_WGS) * 32;
constexpr int AT_HD = 64;
constexpr int AT_ROWS = 128;                      // query rows of the MMA (TMEM lanes); window length <= 128
constexpr int AT_SLOT_COLS = 128;                 // TMEM: S in columns 0..127, P aliases 0..63, O aliases 64..127
constexpr int AT_O_COL = 64;
constexpr int AT_LOOKAHEAD = 2;                   // QK^T of item i + 2 is issued before PV of item i
// TR = token rows per TMA tile: 112 when every window fits (14 KB tiles, 5 stages), else 128 (16 KB tiles, 4 stages)
template <int TR>
struct AtCfg {
  static constexpr int TILE_BYTES = TR * AT_HD * 2;
  static constexpr int STAGE_BYTES = 3 * TILE_BYTES;   // Q, K, V
  static constexpr int STAGES = TR <= 112 ? 5 : 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 512;
  static_assert(STAGE_BYTES % 1024 == 0, "stages keep the 1024-byte alignment of SWIZZLE_128B tiles");
  static_assert(SMEM <= 232448, "exceeds the shared memory a CTA can opt in to");
};

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void at_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {
      printf("qasr attention: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// bf16 x bf16 -> fp32, A K-major (or TMEM), B K-major (b_mn = 0) or MN-major (b_mn = 1)
__device__ __forceinline__ uint32_t at_idesc(int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(AT_ROWS >> 4) << 24);
}

template <int TR>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const int2* __restrict__ win, int n_win, int heads, int d, int head_rows,
                    __nv_bfloat16* __restrict__ out, float scale_log2e) {
  constexpr int AT_STAGES = AtCfg<TR>::STAGES, AT_STAGE_BYTES = AtCfg<TR>::STAGE_BYTES, AT_TILE_BYTES = AtCfg<TR>::TILE_BYTES;
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_STAGES * AT_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES] TMA -> MMA: Q, K, V of the stage have landed
  uint64_t* empty_bar = bars + AT_STAGES;       // [STAGES] MMA -> TMA: both products of the item have read the stage
  uint64_t* s_full = bars + 2 * AT_STAGES;      // [SLOTS] MMA -> softmax: S is in TMEM
  uint64_t* p_full = s_full + AT_SLOTS;         // [SLOTS] softmax -> MMA: P is in TMEM
  uint64_t* o_full = s_full + 2 * AT_SLOTS;     // [SLOTS] MMA -> softmax: O is in TMEM
  uint64_t* o_empty = s_full + 3 * AT_SLOTS;    // [SLOTS] softmax -> MMA: the slot is drained (O aliases S: the next QK^T waits for it)
  uint64_t* v_ready = s_full + 4 * AT_SLOTS;    // [STAGES] warp 3 -> MMA: V rows past the window's end are zero
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(v_ready + AT_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_win * heads;

  if (warp == 0 && lane == 0) prefetch_tmap(&tm_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AT_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
      mbar_init(&v_ready[i], 1);
    }
    for (int i = 0; i < AT_SLOTS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);      // one elected lane per softmax warp
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // the prologue overlapped the QKV GEMM's tail; its output is read from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it % AT_STAGES;
      const uint32_t ph = (it / AT_STAGES) & 1;
      at_wait(&empty_bar[s], ph ^ 1);
      if (elect_one()) {
        const int w = item / heads, h = item - w * heads;
        const int row0 = __ldg(&win[w]).x;
        uint8_t* st = smem + s * AT_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], AT_STAGE_BYTES);
        // head-major qkv: plane (section, head) holds that head's [tokens, 64] rows back to back -> one contiguous 16 KB block
        tma_load_2d(st, &tm_qkv, 0, h * head_rows + row0, &full_bar[s]);
        tma_load_2d(st + AT_TILE_BYTES, &tm_qkv, 0, (heads + h) * head_rows + row0, &full_bar[s]);
        tma_load_2d(st + 2 * AT_TILE_BYTES, &tm_qkv, 0, (2 * heads + h) * head_rows + row0, &full_bar[s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // software-pipelined: S(i + AT_LOOKAHEAD) = Q K^T is issued before O(i) = P V, so several items are in softmax at once.
    // A slot is reused by item i + AT_SLOTS; since O lives in the upper half of the S columns, that item's Q K^T waits
    // until the softmax warpgroup has drained O(i) (o_empty).  P V(i) is issued AT_SLOTS - AT_LOOKAHEAD iterations before
    // that wait, so the chain o_full -> drain -> o_empty is never waiting on this warp: no deadlock.
    const uint64_t desc0 = make_smem_desc_sw128(smem_u32(smem));
    auto issue_qk = [&](int it, int item) {
      const int s = it % AT_STAGES, t = it % AT_SLOTS;
      const uint32_t ph = (it / AT_STAGES) & 1;
      const uint32_t pht = (it / AT_SLOTS) & 1;
      const int wl = __ldg(&win[item / heads]).y;
      const int n16 = (wl + 15) >> 4;
      at_wait(&full_bar[s], ph);
      at_wait(&o_empty[t], pht ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq = desc0 + static_cast<uint64_t>((s * AT_STAGE_BYTES) >> 4);
        const uint64_t dk = dq + static_cast<uint64_t>(AT_TILE_BYTES >> 4);
        const uint32_t idesc = at_idesc(n16 * 16, 0);
        const uint32_t t_s = tmem_base + static_cast<uint32_t>(t * AT_SLOT_COLS);
#pragma unroll
        for (int k = 0; k < AT_HD / 16; ++k) umma<K_BF16>(t_s, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int it, int item) {
      const int s = it % AT_STAGES, t = it % AT_SLOTS;
      const uint32_t pht = (it / AT_SLOTS) & 1;
      const int wl = __ldg(&win[item / heads]).y;
      const int n16 = (wl + 15) >> 4;
      at_wait(&p_full[t], pht);
      at_wait(&v_ready[s], (it / AT_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dv = desc0 + static_cast<uint64_t>((s * AT_STAGE_BYTES + 2 * AT_TILE_BYTES) >> 4);
        const uint32_t idesc = at_idesc(AT_HD, 1);
        const uint32_t t_p = tmem_base + static_cast<uint32_t>(t * AT_SLOT_COLS);
        const uint32_t t_o = t_p + AT_O_COL;
        for (int k = 0; k < n16; ++k)  // 16 keys per step: 8 packed-bf16 columns of P, 16 rows (2048 bytes) of V
          umma_bf16_ts(t_o, t_p + static_cast<uint32_t>(8 * k), dv + static_cast<uint64_t>(128 * k), idesc, k != 0 ? 1u : 0u);
        umma_commit(&o_full[t]);
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
    };
    static_assert(AT_LOOKAHEAD >= 1 && AT_LOOKAHEAD < AT_SLOTS, "P V(i) must be issued before Q K^T(i + AT_SLOTS) waits for its drain");
    const int step = gridDim.x;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += step, ++it) {
      issue_qk(it, item);
      if (it >= AT_LOOKAHEAD) issue_pv(it - AT_LOOKAHEAD, item - AT_LOOKAHEAD * step);
    }
    for (int j = it >= AT_LOOKAHEAD ? it - AT_LOOKAHEAD : 0; j < it; ++j) issue_pv(j, static_cast<int>(blockIdx.x) + j * step);
  } else if (warp == 3) {
    // ===================== V tail rows -> zero (clip isolation) =====================
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int s = it % AT_STAGES;
      const uint32_t ph = (it / AT_STAGES) & 1;
      const int wl = __ldg(&win[item / heads]).y;
      const int r1 = ((wl + 15) >> 4) << 4;
      at_wait(&full_bar[s], ph);
      if (r1 > wl) {
        // SWIZZLE_128B permutes 16-byte chunks inside a 128-byte row: row r is bytes [128 r, 128 r + 128) of the tile
        const uint32_t v0 = smem_u32(smem + s * AT_STAGE_BYTES + 2 * AT_TILE_BYTES) + static_cast<uint32_t>(wl) * 128u;
        for (int i = lane; i < (r1 - wl) * 8; i += 32) st_shared_v4(v0 + static_cast<uint32_t>(i) * 16u, make_uint4(0u, 0u, 0u, 0u));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core's reads
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_ready[s]);
    }
  } else if (warp >= 4) {
    // ===================== softmax + output warpgroups =====================
    const int wg = (warp - 4) >> 2;            // 0 .. AT_WGS-1: this warpgroup serves items it % AT_WGS == wg (slot it % AT_SLOTS)
    const int q = warp & 3;                    // TMEM lane quarter of this warp
    const int row = q * 32 + lane;             // query row == TMEM lane
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      if (it % AT_WGS != wg) continue;
      const int slot = it % AT_SLOTS;
      const uint32_t ph = (it / AT_SLOTS) & 1;
      const int w = item / heads, h = item - w * heads;
      const int2 wd = __ldg(&win[w]);
      const int wl = wd.y;
      const int n16 = (wl + 15) >> 4;
      const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(slot * AT_SLOT_COLS);

      at_wait(&s_full[slot], ph);
      tc_fence_after();
      // ONE sweep over the S row: all (<= 128) scores of the row into registers, then max, exp, sum from registers.  TMEM reads run
      // at 64 B/clk/SM and were the largest single cost of the two-pass version (S read twice: 146 KB per item -> 89 KB).
      // Only the last 16-key granule can hold keys past the window's end.
      const int tail = wl - (n16 - 1) * 16;  // live keys in the last granule, 1..16
      // (The granules are spelled out by macro: as loops over a 2-D register array the front end demoted the array to local memory.)
      constexpr int NG = TR / 16;  // key granules a window of this kernel instance can have: 7 (112-row tiles) or 8
      uint32_t v0[16], v1[16], v2[16], v3[16], v4[16], v5[16], v6[16], v7[16];
      tmem_ld16(t_s + 0, v0);   // unconditional: columns past the window's end are loaded and ignored
      tmem_ld16(t_s + 16, v1);
      tmem_ld16(t_s + 32, v2);
      tmem_ld16(t_s + 48, v3);
      tmem_ld16(t_s + 64, v4);
      tmem_ld16(t_s + 80, v5);
      tmem_ld16(t_s + 96, v6);
      if constexpr (NG > 7) tmem_ld16(t_s + 112, v7);
      tmem_ld_wait();
      float mx0 = -INFINITY, mx1 = -INFINITY;
#define AT_MAX_GRANULE(G, V)                                                                        \
  if ((G) < n16 - 1) {                                                                             \
    _Pragma("unroll") for (int j = 0; j < 16; j += 2) {                                            \
      mx0 = fmaxf(mx0, __uint_as_float(V[j]));                                                     \
      mx1 = fmaxf(mx1, __uint_as_float(V[j + 1]));                                                 \
    }                                                                                              \
  } else if ((G) == n16 - 1) {                                                                     \
    _Pragma("unroll") for (int j = 0; j < 16; ++j) mx0 = fmaxf(mx0, j < tail ? __uint_as_float(V[j]) : -INFINITY); \
  }
      AT_MAX_GRANULE(0, v0) AT_MAX_GRANULE(1, v1) AT_MAX_GRANULE(2, v2) AT_MAX_GRANULE(3, v3)
      AT_MAX_GRANULE(4, v4) AT_MAX_GRANULE(5, v5) AT_MAX_GRANULE(6, v6)
      if constexpr (NG > 7) { AT_MAX_GRANULE(7, v7) }
#undef AT_MAX_GRANULE
      // p = exp2((s - max) * scale), row sum in fp32, bf16 pairs back into the slot's first columns (every S column is in registers)
      const float mscaled = fmaxf(mx0, mx1) * scale_log2e;
      float sum0 = 0.f, sum1 = 0.f;
#define AT_EXP_GRANULE(G, V)                                                                        \
  if ((G) < n16) {                                                                                 \
    uint32_t pk[8];                                                                                \
    const int live = (G) < n16 - 1 ? 16 : tail; /* keys at or beyond `live` get p = 0 by select: NaN / Inf scores cannot leak */ \
    _Pragma("unroll") for (int j = 0; j < 16; j += 2) {                                            \
      const float e0 = ex2_approx(fmaf(__uint_as_float(V[j]), scale_log2e, -mscaled));             \
      const float e1 = ex2_approx(fmaf(__uint_as_float(V[j + 1]), scale_log2e, -mscaled));         \
      const float p0 = j < live ? e0 : 0.f, p1 = j + 1 < live ? e1 : 0.f;                          \
      sum0 += p0;                                                                                  \
      sum1 += p1;                                                                                  \
      pk[j >> 1] = pack_bf16x2(p0, p1);                                                            \
    }                                                                                              \
    tmem_st8(t_s + (G) * 8, pk);                                                                   \
  }
      AT_EXP_GRANULE(0, v0) AT_EXP_GRANULE(1, v1) AT_EXP_GRANULE(2, v2) AT_EXP_GRANULE(3, v3)
      AT_EXP_GRANULE(4, v4) AT_EXP_GRANULE(5, v5) AT_EXP_GRANULE(6, v6)
      if constexpr (NG > 7) { AT_EXP_GRANULE(7, v7) }
#undef AT_EXP_GRANULE
      const float sum = sum0 + sum1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[slot]);

      // O = P V is ready: normalise, round, store this row's 64 dims (128 contiguous bytes)
      at_wait(&o_full[slot], ph);
      tc_fence_after();
      const float inv = 1.0f / sum;
      __nv_bfloat16* orow = out + (static_cast<long long>(wd.x) + row) * d + h * AT_HD;
      uint4 packed[AT_HD / 8];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        tmem_ld16(t_s + AT_O_COL + half * 32, o);
        tmem_ld16(t_s + AT_O_COL + half * 32 + 16, o + 16);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 a;
          a.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          a.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          a.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          a.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          packed[half * 4 + g] = a;
        }
      }
      // the slot may be overwritten by the next item's Q K^T as soon as every warp has its O row in registers
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[slot]);
      if (row < wl) {
#pragma unroll
        for (int g = 0; g < AT_HD / 8; ++g) reinterpret_cast<uint4*>(orow)[g] = packed[g];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Token rows per TMA tile for a given longest window: the tensor map's box rows must be this value.
int attention_tc_tile_rows(int max_win_len) { return max_win_len <= 112 ? 112 : 128; }

// tm_qkv: make_tmap_rowmajor over the head-major qkv activation viewed as [3 * heads * head_rows, 64], box rows =
// attention_tc_tile_rows(longest window the handle can see)
cudaError_t launch_window_attention_tc(const CUtensorMap* tm_qkv, int tile_rows, __nv_bfloat16* out, const int2* win, int n_win, int max_win_len,
                                       int d, int heads, int head_rows, int num_sms, cudaStream_t stream) {
  if (n_win == 0) return cudaSuccess;
  if (d != heads * AT_HD || max_win_len > AT_ROWS || max_win_len > tile_rows || (tile_rows != 112 && tile_rows != 128)) return cudaErrorInvalidValue;
  const float scale_log2e = 0.125f * 1.44269504088896340736f;  // head_dim^-0.5 * log2(e)
  const int grid = std::min(n_win * heads, num_sms);
  cudaError_t e;
  if (tile_rows == 112) {
    e = cudaFuncSetAttribute(attention_tc_kernel<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<112>::SMEM);
    if (e != cudaSuccess) return e;
    return launch_pdl(attention_tc_kernel<112>, dim3(grid), dim3(AT_THREADS), AtCfg<112>::SMEM, stream, 1, *tm_qkv, win, n_win, heads, d, head_rows, out,
                      scale_log2e);
  } else {
    e = cudaFuncSetAttribute(attention_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtCfg<128>::SMEM);
    if (e != cudaSuccess) return e;
    return launch_pdl(attention_tc_kernel<128>, dim3(grid), dim3(AT_THREADS), AtCfg<128>::SMEM, stream, 1, *tm_qkv, win, n_win, heads, d, head_rows, out,
                      scale_log2e);
  }
  return cudaGetLastError();
}

}  // namespace qasr

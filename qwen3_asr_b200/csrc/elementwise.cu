// conv1 (1 -> C channels, 3x3 stride 2, +bias, exact GELU), LayerNorm and the windowed
// non-causal multi-head attention of the audio encoder.  Restates
// transformers/models/qwen3_omni_moe/modeling_qwen3_omni_moe.py: conv2d1 :649,730; LayerNorm :576,580,647;
// attention :496-565 / eager definition :471-493 with the per-window block-diagonal structure of :676-693.
#include <cuda_fp8.h>

#include <algorithm>

#include <cmath>
#include <cstring>
#include <algorithm>

#include "kernels.h"
#include "common.cuh"

namespace qasr {
namespace {

// ---------------------------------------------------------------------------------------------
// conv1.  Output layout act1[(chunk*52 + col)][h = 0..63][c], col 0 and 51 are the zero columns the
// implicit-GEMM conv2 reads as padding, col 1 + w holds output column w.  One CTA = one chunk x 13
// output-column slots; one thread = two adjacent channels (bf16x2 stores, 128 B per warp).
// ---------------------------------------------------------------------------------------------
constexpr int C1_COLS = 26;                    // column slots per CTA (52 / 2)
constexpr int C1_TILE_W = 2 * C1_COLS + 2;     // mel frames needed: 53, padded to an even count (LDS.128 alignment)
constexpr int C1_TILE_H = 130;                 // mel bins -1 .. 128
constexpr int C1_THREADS = 480;                // 240 channel pairs x 2 column halves: 15 full warps

template <typename MelT>
__device__ __forceinline__ float mel_load(const MelT* p);
template <>
__device__ __forceinline__ float mel_load<float>(const float* p) { return bf16_round(__ldg(p)); }  // the .to(bf16) cast
template <>
__device__ __forceinline__ float mel_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// Packed fp32 pairs (FFMA2 / FMUL2, sm_100): one issue slot per two lanes of work.  The kernel is issue-bound, so
// the channel pair a thread owns rides in one 64-bit register pair through the 9 taps and the GELU polynomial.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 upk(f32x2_t v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_erf (common.cuh) on a pair: identical operation order per lane, so the results are bit-identical to the scalar form.
__device__ __forceinline__ f32x2_t gelu_erf2(float xa, float xb) {
  const float aa = fabsf(xa), ab = fabsf(xb);
  const f32x2_t a = pk(aa, ab);
  const float2 u = upk(fma2(pk(0.231641888f, 0.231641888f), a, pk(1.0f, 1.0f)));
  const f32x2_t t = pk(rcp_approx(u.x), rcp_approx(u.y));
  f32x2_t p = fma2(pk(0.5307027145f, 0.5307027145f), t, pk(-0.7265760135f, -0.7265760135f));
  p = fma2(p, t, pk(0.7107068705f, 0.7107068705f));
  p = fma2(p, t, pk(-0.142248368f, -0.142248368f));
  p = fma2(p, t, pk(0.127414796f, 0.127414796f));
  const float2 s = upk(mul2(mul2(a, a), pk(-0.72134752044f, -0.72134752044f)));
  const f32x2_t e = pk(ex2_approx(s.x), ex2_approx(s.y));
  const f32x2_t q = mul2(mul2(p, t), e);
  return fma2(pk(-aa, -ab), q, pk(fmaxf(xa, 0.f), fmaxf(xb, 0.f)));
}

template <typename MelT>
__global__ void __launch_bounds__(C1_THREADS, 2) conv1_kernel(const MelT* __restrict__ mel, long long ld, const ChunkDesc* __restrict__ chunks,
                                                              const float* __restrict__ w, const float* __restrict__ bias, int channels,
                                                              __nv_bfloat16* __restrict__ act1) {
  // FFMA2 takes a 32-bit register as a broadcast operand ({x, x} costs no move), so the tile holds plain floats: 15 bytes of
  // shared-memory traffic per output (a {x, x} tile doubled that and made the kernel LDS-bandwidth bound)
  __shared__ __align__(16) float tile[C1_TILE_H][C1_TILE_W];
  const int chunk = blockIdx.x;
  const int slot0 = blockIdx.y * C1_COLS;  // first column slot (0..51) of this CTA
  const ChunkDesc cd = chunks[chunk];
  const int tid = threadIdx.x;

  // mel frames f = 2*(slot-1) - 1 + j for the slots of this CTA: f0 = 2*slot0 - 3
  const int f0 = 2 * slot0 - 3;
  for (int i = tid; i < C1_TILE_H * C1_TILE_W; i += C1_THREADS) {
    const int r = i / C1_TILE_W, cc = i - r * C1_TILE_W;
    const int bin = r - 1, f = f0 + cc;
    float v = 0.f;
    if (bin >= 0 && bin < 128 && f >= 0 && f < cd.valid) v = mel_load<MelT>(mel + bin * ld + cd.mel_col0 + f);
    tile[r][cc] = v;
  }
  __syncthreads();

  const int pair = tid % 240, half = tid / 240;
  const int c0 = 2 * pair;
  if (c0 >= channels) return;
  f32x2_t wv[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wv[t] = pk(w[c0 * 9 + t], w[(c0 + 1) * 9 + t]);
  const f32x2_t bv = pk(bias[c0], bias[c0 + 1]);

  for (int s = half * (C1_COLS / 2); s < (half + 1) * (C1_COLS / 2); ++s) {
    const int slot = slot0 + s;
    const int ow = slot - 1;  // output column of the chunk; slot 0 / 51 are padding
    const bool live = ow >= 0 && ow < cd.w1;
    __nv_bfloat16* dst = act1 + ((static_cast<long long>(chunk) * ACT1_PITCH + slot) * ACT1_H) * channels + c0;
    if (!live) {
      for (int h = 0; h < ACT1_H; ++h) *reinterpret_cast<uint32_t*>(dst + static_cast<long long>(h) * channels) = 0u;
      continue;
    }
    const int cc = 2 * s;  // tile column of frame 2*ow - 1 (even: the {x0,x0,x1,x1} quad is 16-byte aligned)
#pragma unroll 2
    for (int h = 0; h < ACT1_H; ++h) {
      f32x2_t acc = bv;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* row = &tile[2 * h + kh][cc];
        const float2 x01 = *reinterpret_cast<const float2*>(row);
        const float x2 = row[2];
        acc = fma2(wv[kh * 3 + 0], pk(x01.x, x01.x), acc);
        acc = fma2(wv[kh * 3 + 1], pk(x01.y, x01.y), acc);
        acc = fma2(wv[kh * 3 + 2], pk(x2, x2), acc);
      }
      const float2 a = upk(acc);
      const float2 xr = bf16_round2(a.x, a.y);
      const float2 y = upk(gelu_erf2(xr.x, xr.y));
      *reinterpret_cast<uint32_t*>(dst + static_cast<long long>(h) * channels) = pack_bf16x2(y.x, y.y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim.  A warp owns rows warp_global, warp_global + n_warps, ...; its lanes keep their
// gamma / beta columns in registers for all rows (re-loading them per row was 4x the traffic of x itself) and the
// next row's x is fetched while the current one is reduced.  fp32 statistics, two-pass in registers.
// ---------------------------------------------------------------------------------------------
constexpr int LN_WARPS = 8;

// FP8_OUT (QUANTIZE=fp8 variant, per-row scales): the bf16-rounded output row is quantised in the same pass -- row amax by
// warp shuffle, scale = max(amax, 1e-12) / 448, e4m3 to q_out, scale to row_scale -- instead of being written as bf16.
template <int NV, bool FP8_OUT>  // NV = ceil(d / 256): 8-element vectors per lane
__global__ void __launch_bounds__(LN_WARPS * 32, 2) layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                                  uint8_t* __restrict__ q_out, float* __restrict__ row_scale,
                                                                  int rows, int d, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * LN_WARPS;
  int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  float g[NV][8], b[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    if (col < d) {
      *reinterpret_cast<float4*>(&g[i][0]) = __ldg(reinterpret_cast<const float4*>(gamma + col));
      *reinterpret_cast<float4*>(&g[i][4]) = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
      *reinterpret_cast<float4*>(&b[i][0]) = __ldg(reinterpret_cast<const float4*>(beta + col));
      *reinterpret_cast<float4*>(&b[i][4]) = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    }
  }
  uint4 nxt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    nxt[i] = col < d ? *reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * d + col) : make_uint4(0, 0, 0, 0);
  }
  const float inv_d = 1.0f / d;
  for (; row < rows; row += n_warps) {
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float2 t;
      t = unpack_bf16x2(nxt[i].x); v[i][0] = t.x; v[i][1] = t.y;
      t = unpack_bf16x2(nxt[i].y); v[i][2] = t.x; v[i][3] = t.y;
      t = unpack_bf16x2(nxt[i].z); v[i][4] = t.x; v[i][5] = t.y;
      t = unpack_bf16x2(nxt[i].w); v[i][6] = t.x; v[i][7] = t.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
    const int next_row = row + n_warps;
    if (next_row < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 8;
        if (col < d) nxt[i] = *reinterpret_cast<const uint4*>(x + static_cast<long long>(next_row) * d + col);
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 8;
      if (col < d) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; sq = fmaf(t, t, sq); }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
    if constexpr (!FP8_OUT) {
      __nv_bfloat16* orow = out + static_cast<long long>(row) * d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 8;
        if (col < d) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = (v[i][j] - mean) * rstd * g[i][j] + b[i][j];
          uint4 u;
          u.x = pack_bf16x2(y[0], y[1]); u.y = pack_bf16x2(y[2], y[3]);
          u.z = pack_bf16x2(y[4], y[5]); u.w = pack_bf16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(orow + col) = u;
        }
      }
    } else {
      float amax = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 8;
        if (col < d) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 r = bf16_round2((v[i][j] - mean) * rstd * g[i][j] + b[i][j], (v[i][j + 1] - mean) * rstd * g[i][j + 1] + b[i][j + 1]);
            v[i][j] = r.x;  // the module output as the next op sees it: bf16
            v[i][j + 1] = r.y;
            amax = fmaxf(amax, fmaxf(fabsf(r.x), fabsf(r.y)));
          }
        }
      }
      const float scale = fmaxf(warp_max(amax), 1e-12f) / 448.0f;
      const FastDivisor sc(scale);
      uint8_t* qrow = q_out + static_cast<long long>(row) * d;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 8;
        if (col < d) {
          uint32_t w[2];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const unsigned short p0 = __nv_cvt_float2_to_fp8x2(make_float2(sc.div(v[i][4 * hh]), sc.div(v[i][4 * hh + 1])), __NV_SATFINITE, __NV_E4M3);
            const unsigned short p1 = __nv_cvt_float2_to_fp8x2(make_float2(sc.div(v[i][4 * hh + 2]), sc.div(v[i][4 * hh + 3])), __NV_SATFINITE, __NV_E4M3);
            w[hh] = static_cast<uint32_t>(p0) | (static_cast<uint32_t>(p1) << 16);
          }
          *reinterpret_cast<uint2*>(qrow + col) = make_uint2(w[0], w[1]);
        }
      }
      if (lane == 0) row_scale[row] = scale;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Windowed attention.  One CTA = one (window, head); K and V of the whole window live in shared
// memory (window <= 104 tokens in both checkpoints, any length up to kMaxWin supported), each warp
// owns 16-query row tiles and runs an online-softmax loop over 16-key blocks with
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  FLOPs here are ~1% of the encoder; the point of the
// kernel is zero HBM round trips: packed QKV is read once, O written once.
// ---------------------------------------------------------------------------------------------
constexpr int HD = 64;
constexpr int KV_STRIDE = 72;  // bf16 elements per smem row (144 B): conflict-free fragment + ldmatrix access
constexpr int ATT_WARPS = 8;  // a 104-token window is 7 query tiles of 16 rows: one per warp

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__global__ void __launch_bounds__(ATT_WARPS * 32) window_attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                          __nv_bfloat16* __restrict__ out,
                                                                          const int2* __restrict__ win, int d, int heads, long long head_rows,
                                                                          float scale_log2e) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int2 wd = win[blockIdx.x];
  const int start = wd.x, wl = wd.y;
  const int head = blockIdx.y;
  const int wl16 = (wl + 15) & ~15;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sV = sK + wl16 * KV_STRIDE;
  __nv_bfloat16* sQ = sV + wl16 * KV_STRIDE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // head-major qkv: plane (section, head) = [head_rows tokens][64]
  const __nv_bfloat16* qb = qkv + (static_cast<long long>(head) * head_rows + start) * HD;
  const __nv_bfloat16* kb = qkv + (static_cast<long long>(heads + head) * head_rows + start) * HD;
  const __nv_bfloat16* vb = qkv + (static_cast<long long>(2 * heads + head) * head_rows + start) * HD;

  // stage Q, K, V with 16-byte loads, 128 contiguous bytes per row (rows >= wl zero-filled so masked P = 0 never meets garbage)
  for (int i = tid; i < wl16 * 8; i += ATT_WARPS * 32) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q;
    if (r < wl) {
      q = *reinterpret_cast<const uint4*>(qb + r * HD + c);
      k = *reinterpret_cast<const uint4*>(kb + r * HD + c);
      v = *reinterpret_cast<const uint4*>(vb + r * HD + c);
    }
    *reinterpret_cast<uint4*>(sQ + r * KV_STRIDE + c) = q;
    *reinterpret_cast<uint4*>(sK + r * KV_STRIDE + c) = k;
    *reinterpret_cast<uint4*>(sV + r * KV_STRIDE + c) = v;
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  for (int qt = warp; qt * 16 < wl; qt += ATT_WARPS) {
    const int r0 = qt * 16 + g, r1 = r0 + 8;
    // Q fragments for the 4 k16 steps over head_dim (row pitch 144 B: the 32 lanes hit 32 distinct banks)
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + 2 * t;
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(sQ + r0 * KV_STRIDE + c);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(sQ + r1 * KV_STRIDE + c);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(sQ + r0 * KV_STRIDE + c + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(sQ + r1 * KV_STRIDE + c + 8);
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kb = 0; kb < wl16; kb += 16) {
      float s[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const __nv_bfloat16* krow = sK + (kb + nt * 8 + g) * KV_STRIDE + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + ks * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8);
          mma_bf16_16816(s[nt], qa[ks], b0, b1);
        }
      }
      // scale (log2 domain) + mask keys beyond the window
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int key = kb + nt * 8 + 2 * t;
        s[nt][0] = key < wl ? s[nt][0] * scale_log2e : -INFINITY;
        s[nt][1] = key + 1 < wl ? s[nt][1] * scale_log2e : -INFINITY;
        s[nt][2] = key < wl ? s[nt][2] * scale_log2e : -INFINITY;
        s[nt][3] = key + 1 < wl ? s[nt][3] * scale_log2e : -INFINITY;
        bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);  // finite: every block holds >= 1 live key
      const float corr0 = exp2f(m0 - mn0), corr1 = exp2f(m1 - mn1);
      m0 = mn0; m1 = mn1;
      uint32_t pa[4];
      float ps0 = 0.f, ps1 = 0.f;
      {
        const float p00 = exp2f(s[0][0] - mn0), p01 = exp2f(s[0][1] - mn0);
        const float p02 = exp2f(s[0][2] - mn1), p03 = exp2f(s[0][3] - mn1);
        const float p10 = exp2f(s[1][0] - mn0), p11 = exp2f(s[1][1] - mn0);
        const float p12 = exp2f(s[1][2] - mn1), p13 = exp2f(s[1][3] - mn1);
        ps0 = p00 + p01 + p10 + p11;
        ps1 = p02 + p03 + p12 + p13;
        pa[0] = pack_bf16x2(p00, p01);
        pa[1] = pack_bf16x2(p02, p03);
        pa[2] = pack_bf16x2(p10, p11);
        pa[3] = pack_bf16x2(p12, p13);
      }
      l0 = l0 * corr0 + ps0;
      l1 = l1 * corr1 + ps1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
      // O += P V: V^T fragments through ldmatrix.trans; one x4 covers two 8-wide dim tiles
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        // matrices: (keys kb..+7, dims 16dp..+7), (keys kb+8..+15, same dims), then the same for dims 16dp+8..+15
        const int mi = lane >> 3, ri = lane & 7;
        const __nv_bfloat16* vrow = sV + (kb + (mi & 1) * 8 + ri) * KV_STRIDE + dp * 16 + (mi >> 1) * 8;
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, vrow);
        mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
        mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    __nv_bfloat16* ob = out + static_cast<long long>(start) * d + head * HD + 2 * t;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (r0 < wl) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r0) * d + i * 8) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
      if (r1 < wl) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r1) * d + i * 8) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// conv1 on the tensor cores (product path).  The 9-tap, 1-input-channel convolution is a [pixels x 16] x [16 x channels]
// product (taps 9..15 are zero), small enough for warp-level mma.sync.m16n8k16: one instruction replaces 1152 FFMAs, which
// leaves the kernel with what it cannot avoid -- the erf-GELU, the two bf16 roundings and the 3 GB of stores.
//   CTA  = one chunk x 26 output-column slots, mel tile (bf16) in shared memory
//   warp = one block of 32 output channels (15 warps = 480 channels): its weight fragments (4 n-tiles) and biases stay in
//          registers for the whole CTA; it walks the CTA's (column, 16-row block) pixel groups
//   The n-columns of the four n-tiles are PERMUTED over the 32 channels (tile j, column 2t+e <-> channel 8t + 2j + e) so that
//   the accumulator fragments of a thread are 8 CONSECUTIVE channels of its two rows: one 16-byte store per row and thread,
//   64 contiguous bytes per row and quad -- full sectors without a shared-memory transpose.
// ---------------------------------------------------------------------------------------------
constexpr int C1M_COLS = 26;                   // column slots per CTA at batch sizes that fill the GPU (52 / 2)
constexpr int C1M_COLS_SMALL = 4;              // ... and for a handful of chunks (a WS window): 13 CTAs per chunk instead of 2
constexpr int C1M_PITCH = 58;                  // bf16 elements per tile row: <= 53 frames used; 2 rows = 58 words = 26 mod 32 banks
constexpr int C1M_TILE_H = 130;                // mel bins -1 .. 128
constexpr int C1M_WARPS = 15;                  // 480 channels / 32
constexpr int C1M_THREADS = C1M_WARPS * 32;

template <typename MelT>
__device__ __forceinline__ unsigned short mel_load_bf16_bits(const MelT* p);
template <>
__device__ __forceinline__ unsigned short mel_load_bf16_bits<float>(const float* p) {  // the .to(bf16) cast
  return __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p)));
}
template <>
__device__ __forceinline__ unsigned short mel_load_bf16_bits<__nv_bfloat16>(const __nv_bfloat16* p) {
  return *reinterpret_cast<const unsigned short*>(p);
}

template <typename MelT, int COLS, bool LUT>
__global__ void __launch_bounds__(C1M_THREADS, 2) conv1_mma_kernel(const MelT* __restrict__ mel, long long ld, const ChunkDesc* __restrict__ chunks,
                                                                  const float* __restrict__ w, const float* __restrict__ bias,
                                                                  const float* __restrict__ gelu_lut, __nv_bfloat16* __restrict__ act1) {
  constexpr int channels = 480;
  __shared__ __align__(16) unsigned short tile[C1M_TILE_H * C1M_PITCH];
  __shared__ float lut_s[LUT ? gelu_tab::WORDS : 1];
  if constexpr (LUT)
    for (int i = threadIdx.x; i < gelu_tab::WORDS; i += C1M_THREADS) lut_s[i] = __ldg(gelu_lut + i);
  const int chunk = blockIdx.x;
  const int slot0 = blockIdx.y * COLS;
  const ChunkDesc cd = chunks[chunk];
  const int tid = threadIdx.x, lane = tid & 31, cb = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  // mel frames f = 2*(slot-1) - 1 + j for the slots of this CTA: f0 = 2*slot0 - 3
  const int f0 = 2 * slot0 - 3;
  for (int i = tid; i < C1M_TILE_H * C1M_PITCH; i += C1M_THREADS) {
    const int r = i / C1M_PITCH, cc = i - r * C1M_PITCH;
    const int bin = r - 1, f = f0 + cc;
    unsigned short v = 0;
    if (bin >= 0 && bin < 128 && f >= 0 && f < cd.valid && cc < 2 * COLS + 1) v = mel_load_bf16_bits<MelT>(mel + bin * ld + cd.mel_col0 + f);
    tile[i] = v;
  }

  // weight fragments (B, "col" layout: b0 = {k = 2t, 2t+1}, b1 = {k = 2t+8, 2t+9} of column n = g) and the biases of the
  // channels this thread's accumulators belong to
  uint32_t b0[4], b1[4];
  float bia[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = cb * 32 + 8 * (g >> 1) + 2 * j + (g & 1);
    b0[j] = pack_bf16x2(__ldg(w + ch * 9 + 2 * t), __ldg(w + ch * 9 + 2 * t + 1));
    b1[j] = t == 0 ? pack_bf16x2(__ldg(w + ch * 9 + 8), 0.f) : 0u;
    bia[j][0] = __ldg(bias + cb * 32 + 8 * t + 2 * j);
    bia[j][1] = __ldg(bias + cb * 32 + 8 * t + 2 * j + 1);
  }
  // tile offsets of this lane's two taps k = 2t, 2t+1 (k = 3 kh + kw) relative to (row 2h, column 2s)
  const int offA = ((2 * t) / 3) * C1M_PITCH + (2 * t) % 3;
  const int offB = ((2 * t + 1) / 3) * C1M_PITCH + (2 * t + 1) % 3;
  constexpr int off8 = 2 * C1M_PITCH + 2;
  __syncthreads();

  for (int s = 0; s < COLS; ++s) {
    const int slot = slot0 + s;
    const int ow = slot - 1;  // output column of the chunk; slot 0 / 51 are padding
    const bool live = ow >= 0 && ow < cd.w1;
    __nv_bfloat16* dst = act1 + ((static_cast<long long>(chunk) * ACT1_PITCH + slot) * ACT1_H) * channels + cb * 32 + 8 * t;
    if (!live) {
#pragma unroll
      for (int i = 0; i < ACT1_H / 8; ++i) *reinterpret_cast<uint4*>(dst + static_cast<long long>(i * 8 + g) * channels) = make_uint4(0, 0, 0, 0);
      continue;
    }
#pragma unroll 1
    for (int h0 = 0; h0 < ACT1_H; h0 += 16) {
      // A fragment: pixel rows h0 + g and h0 + g + 8 of output column s; pixel (h, s) reads tile rows 2h .. 2h+2, columns 2s .. 2s+2
      const unsigned short* p0 = tile + (2 * (h0 + g)) * C1M_PITCH + 2 * s;
      const unsigned short* p1 = p0 + 16 * C1M_PITCH;
      uint32_t a[4];
      a[0] = static_cast<uint32_t>(p0[offA]) | (static_cast<uint32_t>(p0[offB]) << 16);
      a[1] = static_cast<uint32_t>(p1[offA]) | (static_cast<uint32_t>(p1[offB]) << 16);
      a[2] = t == 0 ? static_cast<uint32_t>(p0[off8]) : 0u;
      a[3] = t == 0 ? static_cast<uint32_t>(p1[off8]) : 0u;
      float c[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j][0] = bia[j][0]; c[j][1] = bia[j][1]; c[j][2] = bia[j][0]; c[j][3] = bia[j][1];
        mma_bf16_16816(c[j], a, b0[j], b1[j]);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {  // row g, then row g + 8
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (LUT) {
            o[j] = gelu_lut_bf16x2(lut_s, c[j][2 * half], c[j][2 * half + 1]);
          } else {
            const float2 xr = bf16_round2(c[j][2 * half], c[j][2 * half + 1]);
            o[j] = pack_bf16x2(gelu_erf(xr.x), gelu_erf(xr.y));
          }
        }
        *reinterpret_cast<uint4*>(dst + static_cast<long long>(h0 + g + 8 * half) * channels) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

}  // namespace

cudaError_t launch_conv1(const void* mel, int mel_is_bf16, long long mel_ld, const ChunkDesc* chunks, int n_chunks, const float* w,
                         const float* bias, const float* gelu_lut, int channels, __nv_bfloat16* act1, bool simt, cudaStream_t stream) {
  if (n_chunks == 0) return cudaSuccess;
  if (channels > 480 || (channels & 1)) return cudaErrorInvalidValue;
  if (!simt) {
    if (channels != 480) return cudaErrorInvalidValue;
    static_assert(ACT1_PITCH % C1M_COLS == 0 && ACT1_PITCH % C1M_COLS_SMALL == 0, "column groups tile the 52 slots");
    const bool big = n_chunks * (ACT1_PITCH / C1M_COLS) >= 2 * kNumSMs;
    // few chunks: more, smaller CTAs (the weight fragments are reloaded per CTA, which is noise next to an idle GPU)
    const dim3 grid(n_chunks, ACT1_PITCH / (big ? C1M_COLS : C1M_COLS_SMALL));
    auto go = [&](auto kern, auto melp) { kern<<<grid, C1M_THREADS, 0, stream>>>(melp, mel_ld, chunks, w, bias, gelu_lut, act1); };
    const auto* mb = static_cast<const __nv_bfloat16*>(mel);
    const auto* mf = static_cast<const float*>(mel);
    const bool lut = gelu_lut != nullptr;
    if (big) {
      if (mel_is_bf16) lut ? go(conv1_mma_kernel<__nv_bfloat16, C1M_COLS, true>, mb) : go(conv1_mma_kernel<__nv_bfloat16, C1M_COLS, false>, mb);
      else lut ? go(conv1_mma_kernel<float, C1M_COLS, true>, mf) : go(conv1_mma_kernel<float, C1M_COLS, false>, mf);
    } else {
      if (mel_is_bf16) lut ? go(conv1_mma_kernel<__nv_bfloat16, C1M_COLS_SMALL, true>, mb) : go(conv1_mma_kernel<__nv_bfloat16, C1M_COLS_SMALL, false>, mb);
      else lut ? go(conv1_mma_kernel<float, C1M_COLS_SMALL, true>, mf) : go(conv1_mma_kernel<float, C1M_COLS_SMALL, false>, mf);
    }
    return cudaGetLastError();
  }
  // checker (QASR_DEBUG_SIMT=1): the fp32 FFMA version
  dim3 grid(n_chunks, ACT1_PITCH / C1_COLS);
  constexpr int smem = 0;
  if (mel_is_bf16)
    conv1_kernel<__nv_bfloat16><<<grid, C1_THREADS, smem, stream>>>(static_cast<const __nv_bfloat16*>(mel), mel_ld, chunks, w, bias, channels, act1);
  else
    conv1_kernel<float><<<grid, C1_THREADS, smem, stream>>>(static_cast<const float*>(mel), mel_ld, chunks, w, bias, channels, act1);
  return cudaGetLastError();
}

namespace {
// (mean, rstd) of every row: what LayerNorm would normalise with (same two-pass fp32 arithmetic as layernorm_kernel), for the
// Linears that have LayerNorm folded in (epilogues.cuh, LnFold).  One warp per row, nothing kept but the row: 48 warps per SM.
template <int NV>
__global__ void __launch_bounds__(256) ln_stats_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ stats, int rows, int d, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    const uint4 raw = col < d ? *reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * d + col) : make_uint4(0, 0, 0, 0);
    float2 t;
    t = unpack_bf16x2(raw.x); v[i][0] = t.x; v[i][1] = t.y;
    t = unpack_bf16x2(raw.y); v[i][2] = t.x; v[i][3] = t.y;
    t = unpack_bf16x2(raw.z); v[i][4] = t.x; v[i][5] = t.y;
    t = unpack_bf16x2(raw.w); v[i][6] = t.x; v[i][7] = t.y;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += v[i][j];
  }
  const float inv_d = 1.0f / d;
  const float mean = warp_sum(sum) * inv_d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    if (col < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; sq = fmaf(t, t, sq); }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}
}  // namespace

cudaError_t launch_ln_stats(const __nv_bfloat16* x, float2* stats, int rows, int d, float eps, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (d % 8 != 0 || d > 2048) return cudaErrorInvalidValue;
  const int grid = (rows + 7) / 8;
  const int nv = (d + 255) / 256;
  if (nv <= 1) return launch_pdl(ln_stats_kernel<1>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, d, eps);
  if (nv <= 2) return launch_pdl(ln_stats_kernel<2>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, d, eps);
  if (nv <= 4) return launch_pdl(ln_stats_kernel<4>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, d, eps);
  return launch_pdl(ln_stats_kernel<8>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, d, eps);
}

namespace {
template <bool FP8_OUT>
cudaError_t launch_ln(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* out, uint8_t* q_out, float* row_scale,
                      int rows, int d, float eps, cudaStream_t stream) {
  // ~5 rows per warp at the C2 batch (12480 rows): enough rows to amortise the gamma/beta registers, enough warps to fill the chip
  const int grid = std::max(1, std::min((rows + LN_WARPS - 1) / LN_WARPS, 2 * kNumSMs));
  const int nv = (d + 255) / 256;
  const dim3 g(grid), b(LN_WARPS * 32);
  if (nv <= 1) return launch_pdl(layernorm_kernel<1, FP8_OUT>, g, b, 0, stream, 1, x, gamma, beta, out, q_out, row_scale, rows, d, eps);
  if (nv <= 2) return launch_pdl(layernorm_kernel<2, FP8_OUT>, g, b, 0, stream, 1, x, gamma, beta, out, q_out, row_scale, rows, d, eps);
  if (nv <= 4) return launch_pdl(layernorm_kernel<4, FP8_OUT>, g, b, 0, stream, 1, x, gamma, beta, out, q_out, row_scale, rows, d, eps);
  return launch_pdl(layernorm_kernel<8, FP8_OUT>, g, b, 0, stream, 1, x, gamma, beta, out, q_out, row_scale, rows, d, eps);
}
}  // namespace

cudaError_t launch_layernorm(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* out, int rows, int d,
                             float eps, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (d % 8 != 0 || d > 2048) return cudaErrorInvalidValue;
  return launch_ln<false>(x, gamma, beta, out, nullptr, nullptr, rows, d, eps, stream);
}

cudaError_t launch_layernorm_fp8(const __nv_bfloat16* x, const float* gamma, const float* beta, uint8_t* q_out, float* row_scale, int rows,
                                 int d, float eps, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (d % 8 != 0 || d > 2048) return cudaErrorInvalidValue;
  return launch_ln<true>(x, gamma, beta, nullptr, q_out, row_scale, rows, d, eps, stream);
}

cudaError_t launch_window_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* win, int n_win, int max_win_len, int d,
                                    int heads, int head_rows, cudaStream_t stream) {
  if (n_win == 0) return cudaSuccess;
  if (d != heads * HD) return cudaErrorInvalidValue;
  const int wl16 = (max_win_len + 15) & ~15;
  const size_t smem = static_cast<size_t>(3) * wl16 * KV_STRIDE * sizeof(__nv_bfloat16);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  const float scale_log2e = 0.125f * 1.44269504088896340736f;  // head_dim^-0.5 * log2(e)
  dim3 grid(n_win, heads);
  window_attention_kernel<<<grid, ATT_WARPS * 32, smem, stream>>>(qkv, out, win, d, heads, head_rows, scale_log2e);
  return cudaGetLastError();
}


// ---- GELU table (common.cuh, gelu_tab): built once per handle on the host in double, then checked over all 65 536 bf16 inputs ----
namespace {
// round-to-nearest-even of a double to bf16 (normal and denormal range), returned as the float it is
float bf16_rn_from_double(double v) {
  if (v == 0.0 || std::isnan(v) || std::isinf(v)) return static_cast<float>(v);
  int e = std::ilogb(std::fabs(v));
  if (e < -126) e = -126;
  const double quantum = std::ldexp(1.0, e - 7);
  return static_cast<float>(std::nearbyint(v / quantum) * quantum);   // default rounding mode: ties to even
}
uint32_t f32_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
float bits_f32(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
float bf16_rn_from_float(float f) {   // what cvt.rn.bf16x2.f32 does (finite inputs)
  const uint32_t u = f32_bits(f);
  return bits_f32((u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u);
}
}  // namespace

bool build_gelu_lut(float* lut) {
  using namespace gelu_tab;
  auto target = [](float x) { return bf16_rn_from_double(0.5 * static_cast<double>(x) * (1.0 + std::erf(static_cast<double>(x) * 0.70710678118654752440))); };
  for (int sgn = 0; sgn < 2; ++sgn) {
    float* t = lut + sgn * N;
    t[0] = 0.5f;
    t[N - 1] = sgn ? 0.0f : 1.0f;
    for (uint32_t a = A_LO; a <= A_HI; ++a) {
      const float x = bits_f32(((sgn ? 0x8000u : 0u) | a) << 16);
      t[a - (A_LO - 1)] = static_cast<float>(static_cast<double>(target(x)) / static_cast<double>(x));
    }
  }
  // the device evaluates bf16_rn(x * r): it must be the correctly rounded GELU for EVERY finite input
  for (uint32_t b = 0; b < 65536u; ++b) {
    const float x = bits_f32(b << 16);
    if (std::isnan(x) || std::isinf(x)) continue;
    const uint32_t a = b & 0x7fffu;
    const uint32_t i = std::min(std::max(a, A_LO - 1u), A_HI + 1u) - (A_LO - 1u);
    const float got = bf16_rn_from_float(x * lut[(b >> 15) * N + i]);
    const float want = target(x);
    if (f32_bits(got) != f32_bits(want) && !(got == 0.f && want == 0.f)) return false;
  }
  return true;
}

}  // namespace qasr

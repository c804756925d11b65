// Fused GEMM epilogues.  Each consumes 16 consecutive fp32 accumulator columns of one row and
// writes bf16.  Rounding points mirror the reference's bf16 deployment:
// every nn.Module output is rounded to bf16 before the next op consumes it.
#pragma once

#include "common.cuh"

namespace qasr {

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf16x2(v[0], v[1]);   a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]);   a.w = pack_bf16x2(v[6], v[7]);
  b.x = pack_bf16x2(v[8], v[9]);   b.y = pack_bf16x2(v[10], v[11]);
  b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}
__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* src, float (&v)[16]) {
  const uint4 a = reinterpret_cast<const uint4*>(src)[0];
  const uint4 b = reinterpret_cast<const uint4*>(src)[1];
  float2 t;
  t = unpack_bf16x2(a.x); v[0] = t.x;  v[1] = t.y;
  t = unpack_bf16x2(a.y); v[2] = t.x;  v[3] = t.y;
  t = unpack_bf16x2(a.z); v[4] = t.x;  v[5] = t.y;
  t = unpack_bf16x2(a.w); v[6] = t.x;  v[7] = t.y;
  t = unpack_bf16x2(b.x); v[8] = t.x;  v[9] = t.y;
  t = unpack_bf16x2(b.y); v[10] = t.x; v[11] = t.y;
  t = unpack_bf16x2(b.z); v[12] = t.x; v[13] = t.y;
  t = unpack_bf16x2(b.w); v[14] = t.x; v[15] = t.y;
}
__device__ __forceinline__ void load16_f32(const float* src, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src) + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

enum EpiAct : int { ACT_NONE = 0, ACT_GELU = 1 };

struct NoPrefetch {};

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// The GEMM kernel computes y = act(bf16(acc + bias[n])) per element (zero for rows that are not live), stages it as
// bf16 and then hands 8 consecutive columns at a time to finish(), which stores them (after an optional addend).

// out[m, n] = bf16( act( bf16(acc + bias[n]) ) (+ residual[m, n]) )      -- nn.Linear (+GELU) (+ residual add)
template <int ACT, bool RESIDUAL>
struct EpiLinear {
  __nv_bfloat16* out;
  const float* bias;               // [N] or nullptr
  const __nv_bfloat16* residual;   // [M, ldo] (may alias out) when RESIDUAL
  long long ldo;
  int m_valid, n_valid;
  struct Prefetch { uint4 r; };
  static constexpr bool kScaled = false;
  static constexpr bool kLnFold = false;
  static constexpr bool kLnPart = false;
  static constexpr bool kRowStats = false;
  static constexpr bool kRowAtomic = false;
  __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  __device__ __forceinline__ int n_cols() const { return n_valid; }
  __device__ __forceinline__ bool row_live(int) const { return true; }
  __device__ __forceinline__ int stats_row(int m) const { return m < m_valid ? m : -1; }
  __device__ __forceinline__ float act(float v) const { return ACT == ACT_GELU ? gelu_erf(v) : v; }
  __device__ __forceinline__ long long offset(int m, int n) const { return (m < m_valid && n < n_valid) ? m * ldo + n : -1; }
  __device__ __forceinline__ Prefetch prefetch(int, int, long long off) const {
    Prefetch p;
    if (RESIDUAL) p.r = *reinterpret_cast<const uint4*>(residual + off);
    return p;
  }
  // stores and returns the 8 final bf16 values
  __device__ __forceinline__ uint4 finish(long long off, const uint4& y, const Prefetch& p) const {
    if (RESIDUAL) {
      float a[8], r[8];
      unpack8(y, a);
      unpack8(p.r, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += r[j];
      const uint4 v = pack8(a);
      *reinterpret_cast<uint4*>(out + off) = v;
      return v;
    } else {
      *reinterpret_cast<uint4*>(out + off) = y;
      return y;
    }
  }
};

// nn.Linear whose output rows are scattered: GEMM row m lands in row row_map[m] of `out` (the fused
// inputs_embeds.masked_scatter of qasr_encode_scatter).  A separate type so that the hot residual / GELU epilogues carry no
// run-time check (a nullable row map inside EpiLinear cost the K = 1024 out_proj GEMM 20 %).
struct EpiLinearRows : EpiLinear<ACT_NONE, false> {
  const long long* row_map;  // [M] device
  __device__ __forceinline__ long long offset(int m, int n) const {
    return (m < m_valid && n < n_valid) ? __ldg(row_map + m) * ldo + n : -1;
  }
};

// The fused q|k|v projection writes HEAD-MAJOR: out[(section * heads + head) * head_rows + token][64], section = q, k, v.
// A 64-column group of the [tokens, 3d] product is one head's 128-byte row, so the stores are the same full-line segments
// as in the row-major layout, but a (window, head) tile of the attention kernel becomes ONE contiguous block of
// tokens x 128 bytes instead of 128-byte pieces 6 KB apart.
struct EpiQkv : EpiLinear<ACT_NONE, false> {
  long long head_rows;  // token capacity of one (section, head) plane
  __device__ __forceinline__ long long offset(int m, int n) const {
    return (m < m_valid && n < n_valid) ? (static_cast<long long>(n >> 6) * head_rows + m) * 64 + (n & 63) : -1;
  }
};

// FP8 (e4m3 x e4m3) variants: the fp32 accumulator is first dequantised with the dynamic activation scale of its row
// and the weight scale of its column -- y = acc * sa[m] * sw[n] + bias, what torch._scaled_mm computes for torchao's
// Float8DynamicActivationFloat8WeightConfig (reference src/server.py:362-371) -- then follows the bf16 epilogue unchanged.
template <class Base>
struct Scaled : Base {
  const float* row_scale_p;   // [M] activation scales (per-tensor mode: all equal)
  const float* col_scale_p;   // [N] weight scales (per-tensor mode: equal within a module)
  static constexpr bool kScaled = true;
  __device__ __forceinline__ const float* col_scale_ptr() const { return col_scale_p; }
  __device__ __forceinline__ float row_scale(int m) const { return m < this->m_valid ? __ldg(row_scale_p + m) : 0.f; }
};

// LayerNorm folded into the Linear that consumes it:  LN(x) W^T + b = rstd * (x (gamma . W)^T - mean * colsum) + (beta W^T + b),
// colsum[n] = sum_k (gamma . W)[n, k].  The GEMM runs on the raw residual stream x with the pre-scaled weight; the epilogue
// applies the row statistics: y = (acc - mean[m] * colsum[n]) * rstd[m] + bias'[n], then the base epilogue (rounding, activation,
// layout) unchanged.  Saves the LayerNorm kernel's pass over the activation and one bf16 rounding point.
template <class Base>
struct LnFold : Base {
  const float2* stats_p;    // [M] (mean, rstd) of the rows of x
  const float* colsum_p;    // [N]
  static constexpr bool kLnFold = true;
  static constexpr bool kLnPart = false;
  __device__ __forceinline__ const float* col_scale_ptr() const { return colsum_p; }
  __device__ __forceinline__ float2 row_stats(int m) const { return m < this->m_valid ? __ldg(stats_p + m) : make_float2(0.f, 0.f); }
};
// ... with the row statistics finalised INSIDE the consuming GEMM: the producers' residual epilogues (RowStats<>) left per-panel
// (sum, sum of squares) partials; the GEMM kernel's two otherwise idle warps reduce the partials of the tile's 128 rows into a
// shared-memory table while the tile's MMAs run.  No statistics kernel, no extra launch.
template <class Base>
struct LnFoldPart : LnFold<Base> {
  const float2* part_p;     // [M][n_panels]
  int n_panels;             // K / 32 (<= 64)
  float inv_d, eps;
  static constexpr bool kLnPart = true;
};

// Implicit-GEMM convolution epilogue: row m = (global output column g, output row h) with
// g = chunk * slots + ow.  Writes bf16(gelu(bf16(acc + bias))) into the next layer's
// [column][row][channel] layout at column chunk * out_pitch + out_off + ow; columns at or beyond
// the chunk's valid width are written as zeros (they are the next conv's zero padding), slots
// beyond `max_w` are skipped (they alias the neighbouring chunk's padding column).
struct EpiConv {
  __nv_bfloat16* out;
  const float* bias;       // [C]
  const int* width;        // [n_chunks] valid output width of this layer
  int hc;                  // output rows (mel axis) per column: 32 (conv2) or 16 (conv3)
  int slots;               // output column slots per chunk: 26 (conv2) or 13 (conv3)
  int max_w;               // 25 / 13
  int out_pitch, out_off;  // destination column = chunk * out_pitch + out_off + ow
  int n_chunks, c;         // c = channels (480)
  typedef NoPrefetch Prefetch;
  static constexpr bool kScaled = false;
  static constexpr bool kLnFold = false;
  static constexpr bool kLnPart = false;
  static constexpr bool kRowStats = false;
  static constexpr bool kRowAtomic = false;
  __device__ __forceinline__ const float* bias_ptr() const { return bias; }
  __device__ __forceinline__ int n_cols() const { return c; }
  __device__ __forceinline__ bool row_live(int m) const {
    const int g = m / hc;
    const int chunk = g / slots, ow = g - chunk * slots;
    return chunk < n_chunks && ow < __ldg(width + chunk);
  }
  __device__ __forceinline__ float act(float v) const { return gelu_erf(v); }
  __device__ __forceinline__ long long offset(int m, int n) const {
    if (n >= c) return -1;
    const int g = m / hc, h = m - g * hc;
    const int chunk = g / slots, ow = g - chunk * slots;
    if (chunk >= n_chunks || ow >= max_w) return -1;
    const long long col = static_cast<long long>(chunk) * out_pitch + out_off + ow;
    return (col * hc + h) * c + n;
  }
  __device__ __forceinline__ Prefetch prefetch(int, int, long long) const { return Prefetch(); }
  __device__ __forceinline__ uint4 finish(long long off, const uint4& y, const Prefetch&) const {
    *reinterpret_cast<uint4*>(out + off) = y;
    return y;
  }
};

// conv_out epilogue: row m = chunk * 13 + t.  x[token] = bf16( bf16(acc) + pe[t] ), dropped when the
// row is not a valid token (row_token[m] < 0).  Positions restart at 0 in every chunk.
struct EpiConvOut {
  __nv_bfloat16* out;      // [tokens, d]
  const float* pe;         // [tok_per_chunk, d] (values already bf16-representable)
  const int* row_token;    // [m_valid]
  int tok_per_chunk;       // 13
  int d, m_valid;
  struct Prefetch { float4 a, b; };
  static constexpr bool kScaled = false;
  static constexpr bool kLnFold = false;
  static constexpr bool kLnPart = false;
  static constexpr bool kRowStats = false;
  static constexpr bool kRowAtomic = false;
  __device__ __forceinline__ int stats_row(int m) const { return m < m_valid ? __ldg(row_token + m) : -1; }
  __device__ __forceinline__ const float* bias_ptr() const { return nullptr; }
  __device__ __forceinline__ int n_cols() const { return d; }
  __device__ __forceinline__ bool row_live(int) const { return true; }
  __device__ __forceinline__ float act(float v) const { return v; }
  __device__ __forceinline__ long long offset(int m, int n) const {
    if (m >= m_valid || n >= d) return -1;
    const int tok = __ldg(row_token + m);
    return tok < 0 ? -1 : static_cast<long long>(tok) * d + n;
  }
  __device__ __forceinline__ Prefetch prefetch(int m, int n, long long) const {
    const float* p = pe + (m % tok_per_chunk) * d + n;
    Prefetch r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  __device__ __forceinline__ uint4 finish(long long off, const uint4& y, const Prefetch& p) const {
    float a[8];
    unpack8(y, a);
    a[0] += p.a.x; a[1] += p.a.y; a[2] += p.a.z; a[3] += p.a.w;
    a[4] += p.b.x; a[5] += p.b.y; a[6] += p.b.z; a[7] += p.b.w;
    const uint4 v = pack8(a);
    *reinterpret_cast<uint4*>(out + off) = v;
    return v;
  }
};

// The epilogues that write the residual stream (conv_out, out_proj, fc2) can also leave what the NEXT LayerNorm needs: per row and
// 32-column panel the sum and the sum of squares of the bf16 values just stored, in fixed slots part[row][panel] (deterministic:
// no atomics).  The consuming GEMM (LnFoldPart<>) turns the d / 32 partials of a row into (mean, rstd) in its idle warps.
template <class Base>
struct RowStats : Base {
  float2* part;      // [rows][n_panels]
  int n_panels;      // d / 32
  static constexpr bool kRowStats = true;
};

// ... or ACCUMULATE them: every (row, 32-column panel) adds its (sum, sum of squares) to ONE pair of 64-bit integers per row with
// atomicAdd, in fixed point (sum * 2^20, sum of squares * 2^16, each rounded to nearest).  Integer addition is associative, so the
// totals do not depend on the order in which the CTAs arrive: deterministic and batch-invariant like everything else.  The
// consuming GEMM (LnFoldAcc<>) reads the pair of its row in its epilogue -- nothing to finalise, no statistics kernel, no table.
template <class Base>
struct RowStatsAtomic : Base {
  unsigned long long* acc;   // [rows][2], zeroed before the forward
  float2* part = nullptr;    // unused (the kernel's non-atomic branch is compiled out)
  int n_panels = 0;
  static constexpr bool kRowStats = true;
  static constexpr bool kRowAtomic = true;
};
template <class Base>
struct LnFoldAcc : Base {
  const unsigned long long* acc_p;   // [M][2]
  const float* colsum_p;             // [N]
  float inv_d, eps;
  static constexpr bool kLnFold = true;
  static constexpr bool kLnPart = false;
  __device__ __forceinline__ const float* col_scale_ptr() const { return colsum_p; }
  __device__ __forceinline__ float2 row_stats(int m) const {
    if (m >= this->m_valid) return make_float2(0.f, 0.f);
    const ulonglong2 a = __ldcg(reinterpret_cast<const ulonglong2*>(acc_p) + m);
    const float mean = static_cast<float>(static_cast<long long>(a.x)) * (inv_d / kStatSumScale);
    const float ex2 = static_cast<float>(static_cast<long long>(a.y)) * (inv_d / kStatSqScale);
    const float var = fmaxf(fmaf(-mean, mean, ex2), 0.f);
    return make_float2(mean, rsqrtf(var + eps));
  }
};

}  // namespace qasr

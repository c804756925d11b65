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

// out[m, n] = bf16( act( bf16(acc + bias[n]) ) (+ residual[m, n]) )      -- nn.Linear (+GELU) (+ residual add)
template <int ACT, bool RESIDUAL>
struct EpiLinear {
  __nv_bfloat16* out;
  const float* bias;               // [N] or nullptr
  const __nv_bfloat16* residual;   // [M, ldo] (may alias out) when RESIDUAL
  long long ldo;
  int m_valid, n_valid;
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[16]) const {
    if (m >= m_valid || n0 >= n_valid) return;
    float v[16], b[16];
    if (bias != nullptr) {
      load16_f32(bias + n0, b);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) b[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x = bf16_round(acc[j] + b[j]);
      if (ACT == ACT_GELU) x = gelu_erf(x);
      v[j] = x;
    }
    if (RESIDUAL) {
      float r[16];
      load16_bf16(residual + m * ldo + n0, r);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = bf16_round(v[j]) + r[j];
    }
    store16_bf16(out + m * ldo + n0, v);
  }
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
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[16]) const {
    if (n0 >= c) return;
    const int g = m / hc, h = m - g * hc;
    const int chunk = g / slots, ow = g - chunk * slots;
    if (chunk >= n_chunks || ow >= max_w) return;
    float v[16];
    if (ow < __ldg(width + chunk)) {
      float b[16];
      load16_f32(bias + n0, b);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(bf16_round(acc[j] + b[j]));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    const long long col = static_cast<long long>(chunk) * out_pitch + out_off + ow;
    store16_bf16(out + (col * hc + h) * c + n0, v);
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
  __device__ __forceinline__ void operator()(int m, int n0, const float (&acc)[16]) const {
    if (m >= m_valid || n0 >= d) return;
    const int tok = __ldg(row_token + m);
    if (tok < 0) return;
    const int t = m % tok_per_chunk;
    float p[16], v[16];
    load16_f32(pe + t * d + n0, p);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = bf16_round(acc[j]) + p[j];
    store16_bf16(out + static_cast<long long>(tok) * d + n0, v);
  }
};

}  // namespace qasr

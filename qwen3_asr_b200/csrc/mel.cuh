// Fused log-mel frontend (K1 of SURVEY.md section 2.3): PCM f32 -> Hann-400/hop-160 centred STFT ->
// |X|^2 -> 128-bin Slaney mel -> log10(max(.,1e-10)), plus a per-clip running max for the
// max-8 clamp; a second tiny pass applies clamp and (x+4)/4 in place.
//
// Restates transformers/models/whisper/feature_extraction_whisper.py:135-164 per clip (SURVEY.md
// appendix A.2).  The 400-point real DFT is computed as a 200-point complex Stockham FFT
// (radices 5,5,4,2) in shared memory in fp32, followed by the real-input split.
//
// The per-item math is __host__ __device__ so tests/host/mel_host_test.cu can execute the very
// same functions on the CPU.
#pragma once

#include "common.cuh"

namespace qasr {
namespace mel {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_MELS = 128;
constexpr int N_BINS = 201;
constexpr int NC = 200;          // complex FFT length
constexpr int FB = 16;           // frames per CTA slab
constexpr int SLAB = FB * HOP + (N_FFT - HOP);  // 2800 samples
constexpr int P_PITCH = 208;     // power row pitch (floats)
constexpr int THREADS = 256;
constexpr int MAX_NNZ = 512;

struct Tables {                  // device-resident constants built once on the host in double
  float2 w400[N_FFT];            // exp(-2 pi i k / 400)
  float window[N_FFT];           // periodic Hann
  int fptr[N_MELS + 1];          // CSR over filters: weights [fptr[m], fptr[m+1]) ...
  int flo[N_MELS];               // ... apply to power bins flo[m] + j
  float fw[MAX_NNZ];
};

struct cf { float x, y; };
__host__ __device__ __forceinline__ cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y}; }
__host__ __device__ __forceinline__ cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y}; }
__host__ __device__ __forceinline__ cf cmul(cf a, cf b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__host__ __device__ __forceinline__ cf mul_negi(cf a) { return {a.y, -a.x}; }  // -i * a
__host__ __device__ __forceinline__ cf mul_posi(cf a) { return {-a.y, a.x}; }  // +i * a

// One Stockham radix-R butterfly (decimation in frequency):
//   a_j = X[q + s (p + m j)],  Y[q + s (R p + k)] = (sum_j a_j W_R^{jk}) * exp(-2 pi i p k / n),  n = R m
// `it` in [0, m*s): p = it / s, q = it % s.  tw_step = 400 / n indexes the shared W400 table.
template <int R>
__host__ __device__ __forceinline__ void butterfly(const float2* __restrict__ X, float2* __restrict__ Y,
                                                   const float2* __restrict__ w400, int it, int s, int m, int tw_step) {
  const int p = it / s, q = it - p * s;
  cf a[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const float2 v = X[q + s * (p + m * j)];
    a[j] = {v.x, v.y};
  }
  cf b[R];
  if (R == 2) {
    b[0] = cadd(a[0], a[1]);
    b[1] = csub(a[0], a[1]);
  } else if (R == 4) {
    const cf t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
    const cf t2 = cadd(a[1], a[3]), t3 = csub(a[1], a[3]);
    b[0] = cadd(t0, t2);
    b[2] = csub(t0, t2);
    b[1] = cadd(t1, mul_negi(t3));
    b[3] = cadd(t1, mul_posi(t3));
  } else {  // R == 5
    const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
    const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
    const cf t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]);
    const cf t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
    b[0] = {a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y};
    const cf m1 = {a[0].x + c1 * t1.x + c2 * t2.x, a[0].y + c1 * t1.y + c2 * t2.y};
    const cf m2 = {a[0].x + c2 * t1.x + c1 * t2.x, a[0].y + c2 * t1.y + c1 * t2.y};
    const cf n1 = {s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
    const cf n2 = {s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
    b[1] = cadd(m1, mul_negi(n1));
    b[4] = cadd(m1, mul_posi(n1));
    b[2] = cadd(m2, mul_negi(n2));
    b[3] = cadd(m2, mul_posi(n2));
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    cf o = b[k];
    if (k > 0 && p > 0) {
      const float2 w = w400[(p * k * tw_step) % N_FFT];
      o = cmul(o, cf{w.x, w.y});
    }
    Y[q + s * (R * p + k)] = make_float2(o.x, o.y);
  }
}

// Real-input split: power of bin k of the 400-point real DFT from the 200-point FFT Z of
// z[n] = x[2n] + i x[2n+1].
__host__ __device__ __forceinline__ float power_bin(const float2* __restrict__ Z, const float2* __restrict__ w400, int k) {
  const float2 zk = Z[k == NC ? 0 : k];
  const float2 zc = Z[k == 0 ? 0 : NC - k];  // conj taken below
  const cf e = {0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y)};
  const cf o = {0.5f * (zk.x - zc.x), 0.5f * (zk.y + zc.y)};
  const float2 w = w400[k];
  const cf wo = cmul(cf{w.x, w.y}, o);
  const cf x = cadd(e, mul_negi(wo));
  return x.x * x.x + x.y * x.y;
}

__host__ __device__ __forceinline__ int reflect_index(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

}  // namespace mel
}  // namespace qasr

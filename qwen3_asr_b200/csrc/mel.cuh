// Fused log-mel frontend (K1 of SURVEY.md section 2.3): PCM f32 -> Hann-400 / hop-160 centred STFT -> |X|^2 ->
// 128-bin Slaney mel -> log10(max(., 1e-10)) -> max(x, clip max - 8) -> (x + 4) / 4, one kernel.
//
// Restates transformers/models/whisper/feature_extraction_whisper.py:135-164 per clip.
//
// The 400-point real DFT of a windowed frame u[n] is computed in registers as a two-stage real-input
// Cooley-Tukey, n = 16 m + j, k = k1 + 25 k2:
//   stage 1   Y_j[k1] = sum_m u[16 m + j] W25^(k1 m)              16 real 25-point DFTs (5 x 5), k1 = 0..12 only:
//                                                                 Y_j[25 - k1] = conj(Y_j[k1]) for real input
//   stage 2   X[k1 + 25 k2] = sum_j (Y_j[k1] W400^(k1 j)) W16^(k2 j)   13 complex 16-point FFTs (4 x 4)
// The 13 x 16 = 208 stage-2 outputs are exactly the bins k and 400 - k needed for |X[0..200]|^2 (201 unique), so
// there is no Hermitian post-processing pass and only ONE shared-memory exchange between the stages.
#pragma once

#include "common.cuh"

namespace qasr {
namespace mel {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_MELS = 128;
constexpr int N_BINS = 201;
constexpr int FB = 32;                  // frames per work item (one lane per frame in the mel phase)
constexpr int THREADS = 256;
constexpr int SLAB_ROWS = FB + 2;       // hops of 160 samples covering FB frames: 31 * 160 + 400 = 5360 <= 34 * 160
constexpr int SLAB_PITCH = 176;         // floats per hop row: 176 = 16 mod 32, so the two frames of a warp hit disjoint banks
constexpr int ROW_COPY = HOP + 4;       // floats per bulk-copied hop row: 160 + up to 3 of 16-byte alignment slack
constexpr int E_PITCH = 18;             // float2 per exchange row (16 used): 144 B, conflict-free 16-byte reads, no swizzle
constexpr int K1 = 13;                  // stage-1 outputs kept per 25-point DFT
constexpr int P_PITCH = 209;            // power row pitch (odd: conflict-free for lanes = frames and lanes = bins)
constexpr int CLAMP_TILE = 256;         // frames per clamp work item
constexpr int MAX_NNZ = 512;

#include "mel_structure.inc"

struct Tables {                         // device-resident constants built once on the host in double
  float window[N_FFT];                  // periodic Hann
  float2 tw[K1][16];                    // exp(-2 pi i k1 j / 400)
  float fw[MAX_NNZ];                    // mel weights, CSR order (structure: mel_structure.inc)
};

// One unit of work of the persistent kernel, claimed through a global ticket counter.
struct Item {
  int kind;           // 0: FB frames of one clip -> raw log10 mel + clip max;  1: clamp + rescale a tile of a finished clip
  int clip;
  int frame0;         // first frame (kind 0) / first column of the tile (kind 1) within the clip
  int n_frames;       // valid frames (<= FB) / tile width (<= CLAMP_TILE)
  int n_samples;      // clip length in samples (kind 0)
  int need;           // kind 1: number of kind-0 items of the clip that must have finished
  int bulk;           // kind 0: the slab lies inside the clip and inside the PCM buffer -> 34 bulk async row copies
  int pad_;
  long long pcm_off;  // clip start in the packed PCM buffer
  long long col0;     // clip start column in the packed mel
};

// v3 (the default kernel): one unit of work = 32 frames of a clip to transform (n_frames == 0: none) PLUS the clamp of a 32-frame
// tile of a clip that finished `lag` items earlier (c_n_frames == 0: none).  Items are claimed in list order through a ticket.
struct Item3 {
  int clip;
  int frame0;           // first frame within the clip
  int n_frames;         // valid frames (<= FB)
  int n_samples;        // clip length in samples
  long long pcm_off;    // clip start in the packed PCM buffer
  long long col0;       // clip start column in the packed mel
  int c_clip;           // clamp tile: clip, first column within the clip, width (<= FB), frame items of the clip that must be done
  int c_frame0;
  int c_n_frames;
  int c_need;
  long long c_col0;     // that clip's start column
  long long pad_;
};
constexpr int P3_PITCH = 204;           // power row pitch of v3: 16-byte row loads, conflict-free for lanes = frames (204 = 12 mod 32)
constexpr int P3_RING = 3;              // power tiles in flight per CTA
constexpr int v3_grid(int num_sms) { return 2 * num_sms; }
constexpr int v3_clamp_lag(int num_sms) { return 4 * v3_grid(num_sms); }   // a tile is clamped four grid-fulls of items after its clip's last one (19 MB of log-mel)
// counters of a launch: ticket, per-clip done, per-clip max, dump word; padded so that the 16-byte poll of a done counter stays inside
constexpr size_t counter_words(int n_clips) { return 2 * static_cast<size_t>(n_clips) + 8; }

// The per-unit math is __host__ __device__ so tests/host/mel_host_test.cu executes the very same functions on the CPU.
// ---- small complex helpers ---------------------------------------------------------------------
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__host__ __device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

constexpr float C5_1 = 0.30901699437494742f, C5_2 = -0.80901699437494742f;   // cos(2 pi / 5), cos(4 pi / 5)
constexpr float S5_1 = 0.95105651629515357f, S5_2 = 0.58778525229247313f;    // sin(2 pi / 5), sin(4 pi / 5)

// 5-point DFT of real input: r0 real, r1 = X[1], r2 = X[2]  (X[3] = conj r2, X[4] = conj r1)
__host__ __device__ __forceinline__ void rdft5(float a0, float a1, float a2, float a3, float a4, float& r0, float2& r1, float2& r2) {
  const float t1 = a1 + a4, t2 = a2 + a3, t3 = a1 - a4, t4 = a2 - a3;
  r0 = a0 + t1 + t2;
  r1.x = fmaf(C5_2, t2, fmaf(C5_1, t1, a0));
  r1.y = fmaf(-S5_2, t4, -S5_1 * t3);
  r2.x = fmaf(C5_1, t2, fmaf(C5_2, t1, a0));
  r2.y = fmaf(S5_1, t4, -S5_2 * t3);
}
// 5-point DFT of complex input
__host__ __device__ __forceinline__ void cdft5(const float2 (&a)[5], float2 (&b)[5]) {
  const float2 t1 = cadd(a[1], a[4]), t2 = cadd(a[2], a[3]), t3 = csub(a[1], a[4]), t4 = csub(a[2], a[3]);
  b[0] = make_float2(a[0].x + t1.x + t2.x, a[0].y + t1.y + t2.y);
  const float2 m1 = make_float2(fmaf(C5_2, t2.x, fmaf(C5_1, t1.x, a[0].x)), fmaf(C5_2, t2.y, fmaf(C5_1, t1.y, a[0].y)));
  const float2 m2 = make_float2(fmaf(C5_1, t2.x, fmaf(C5_2, t1.x, a[0].x)), fmaf(C5_1, t2.y, fmaf(C5_2, t1.y, a[0].y)));
  const float2 n1 = make_float2(fmaf(S5_2, t4.x, S5_1 * t3.x), fmaf(S5_2, t4.y, S5_1 * t3.y));
  const float2 n2 = make_float2(fmaf(-S5_1, t4.x, S5_2 * t3.x), fmaf(-S5_1, t4.y, S5_2 * t3.y));
  b[1] = make_float2(m1.x + n1.y, m1.y - n1.x);  // m1 - i n1
  b[4] = make_float2(m1.x - n1.y, m1.y + n1.x);  // m1 + i n1
  b[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
  b[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
}
// radix-4 butterfly, forward transform
__host__ __device__ __forceinline__ void r4(float2 a0, float2 a1, float2 a2, float2 a3, float2& o0, float2& o1, float2& o2, float2& o3) {
  const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  o0 = cadd(t0, t2);
  o2 = csub(t0, t2);
  o1 = make_float2(t1.x + t3.y, t1.y - t3.x);  // t1 - i t3
  o3 = make_float2(t1.x - t3.y, t1.y + t3.x);  // t1 + i t3
}

// W25^e = exp(-2 pi i e / 25)
#define QASR_W25(e_cos, e_sin) make_float2(e_cos, -(e_sin))
constexpr float W25C1 = 0.96858316112863108f, W25S1 = 0.24868988716485479f;
constexpr float W25C2 = 0.87630668004386358f, W25S2 = 0.48175367410171532f;
constexpr float W25C3 = 0.72896862742141155f, W25S3 = 0.68454710592868873f;
constexpr float W25C4 = 0.53582679497899666f, W25S4 = 0.84432792550201508f;
constexpr float W25C6 = 0.06279051952931337f, W25S6 = 0.99802672842827156f;
constexpr float W25C8 = -0.42577929156507272f, W25S8 = 0.90482705246601958f;
// W16^e
constexpr float W16C1 = 0.92387953251128674f, W16S1 = 0.38268343236508977f;
constexpr float RSQRT2 = 0.70710678118654752f;

// Stage 1: real 25-point DFT of v[m] (m = 5 m1 + m2), outputs V[0..12] (V[0] real).
__host__ __device__ __forceinline__ void rdft25(const float (&v)[25], float2 (&V)[K1]) {
  float a0[5];
  float2 a1[5], a2[5];
#pragma unroll
  for (int m2 = 0; m2 < 5; ++m2) rdft5(v[m2], v[5 + m2], v[10 + m2], v[15 + m2], v[20 + m2], a0[m2], a1[m2], a2[m2]);
  // twiddle W25^(q m2)
  a1[1] = cmul(a1[1], QASR_W25(W25C1, W25S1));
  a1[2] = cmul(a1[2], QASR_W25(W25C2, W25S2));
  a1[3] = cmul(a1[3], QASR_W25(W25C3, W25S3));
  a1[4] = cmul(a1[4], QASR_W25(W25C4, W25S4));
  a2[1] = cmul(a2[1], QASR_W25(W25C2, W25S2));
  a2[2] = cmul(a2[2], QASR_W25(W25C4, W25S4));
  a2[3] = cmul(a2[3], QASR_W25(W25C6, W25S6));
  a2[4] = cmul(a2[4], QASR_W25(W25C8, W25S8));
  float r0;
  rdft5(a0[0], a0[1], a0[2], a0[3], a0[4], r0, V[5], V[10]);
  V[0] = make_float2(r0, 0.f);
  float2 b[5];
  cdft5(a1, b);  // k = 1, 6, 11, 16, 21
  V[1] = b[0]; V[6] = b[1]; V[11] = b[2]; V[9] = cconj(b[3]); V[4] = cconj(b[4]);
  cdft5(a2, b);  // k = 2, 7, 12, 17, 22
  V[2] = b[0]; V[7] = b[1]; V[12] = b[2]; V[8] = cconj(b[3]); V[3] = cconj(b[4]);
}

// Stage 2: 16-point complex FFT, j = 4 j1 + j2 in, k2 = r + 4 s out (natural order in x[]).
__host__ __device__ __forceinline__ void fft16(float2 (&z)[16]) {
  float2 c[4][4];  // [j2][r]
#pragma unroll
  for (int j2 = 0; j2 < 4; ++j2) r4(z[j2], z[4 + j2], z[8 + j2], z[12 + j2], c[j2][0], c[j2][1], c[j2][2], c[j2][3]);
  // twiddle W16^(r j2)
  c[1][1] = cmul(c[1][1], make_float2(W16C1, -W16S1));                                                    // e = 1
  c[1][2] = make_float2((c[1][2].x + c[1][2].y) * RSQRT2, (c[1][2].y - c[1][2].x) * RSQRT2);              // e = 2
  c[1][3] = cmul(c[1][3], make_float2(W16S1, -W16C1));                                                    // e = 3
  c[2][1] = make_float2((c[2][1].x + c[2][1].y) * RSQRT2, (c[2][1].y - c[2][1].x) * RSQRT2);              // e = 2
  c[2][2] = make_float2(c[2][2].y, -c[2][2].x);                                                           // e = 4: -i
  c[2][3] = make_float2((c[2][3].y - c[2][3].x) * RSQRT2, -(c[2][3].x + c[2][3].y) * RSQRT2);             // e = 6
  c[3][1] = cmul(c[3][1], make_float2(W16S1, -W16C1));                                                    // e = 3
  c[3][2] = make_float2((c[3][2].y - c[3][2].x) * RSQRT2, -(c[3][2].x + c[3][2].y) * RSQRT2);             // e = 6
  c[3][3] = cmul(c[3][3], make_float2(-W16C1, W16S1));                                                    // e = 9
#pragma unroll
  for (int r = 0; r < 4; ++r) r4(c[0][r], c[1][r], c[2][r], c[3][r], z[r], z[r + 4], z[r + 8], z[r + 12]);
}

// Stage-2 output (k1, k2) is X[k1 + 25 k2]: FFT bin k for k <= 200, else its mirror 400 - k (same power for a real
// signal).  The k1 = 0 family meets its own mirrors from k2 = 9 on (stage2_unique false).
__host__ __device__ __forceinline__ int stage2_bin(int k1, int k2) { return k2 <= 7 ? 25 * k2 + k1 : (400 - 25 * k2) - k1; }
__host__ __device__ __forceinline__ bool stage2_unique(int k1, int k2) { return k2 <= 8 || k1 != 0; }

__host__ __device__ __forceinline__ int reflect_index(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

}  // namespace mel
}  // namespace qasr

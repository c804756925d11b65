// Host-side launchers of the tcgen05 GEMM family (tc_gemm.cuh) and the TMA tensor-map builders.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace qasr {

// Resolves cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda, so the
// library loads -- and exports its symbols -- on a box without a driver).  0 on success.
int tmap_api_init();

// bf16 row-major [rows, cols] (leading dimension ld elements), box = box_rows x 64 columns, 128B swizzle.
int make_tmap_rowmajor(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);
// e4m3 (one byte per element) row-major [rows, cols], box = box_rows x 128 columns, 128B swizzle.
int make_tmap_rowmajor_u8(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);
// bf16 activation [g_in columns][h_in rows][c channels] read by a 3x3 stride-2 convolution as
// implicit GEMM: box = 64 channels x hc output rows x gt output columns, traversal stride 2 on
// rows and columns.
int make_tmap_conv(CUtensorMap* tm, const void* base, long long g_in, int h_in, int c, int hc, int gt);

// Largest supported N tile (256 / 128 / 64) that divides n; 0 if none.
int pick_bn(int n);
// Box rows of the small-M weight map for an [n, k] weight: 32 if it divides n; 0 if the path does not apply.
int pick_bn_small(int n, int k);

// The GEMM family runs as CTA pairs (tcgen05.mma.cta_group::2, see tc_gemm.cuh) unless the library is built with
// -DQASR_GEMM_1CTA=1 (A/B builds).  In pair mode each CTA stages half of the weight tile, so the weight tensor maps
// carry a box of bn / 2 rows.
#ifndef QASR_GEMM_1CTA
#define QASR_GEMM_1CTA 0
#endif
constexpr bool kGemmCta2 = QASR_GEMM_1CTA == 0;
constexpr int gemm_b_box_rows(int bn) { return kGemmCta2 ? bn / 2 : bn; }

enum LinearEpi : int { LIN_PLAIN = 0, LIN_GELU = 1, LIN_RESIDUAL = 2, LIN_QKV = 3 /* head-major output, see EpiQkv */ };

struct LinearArgs {
  const CUtensorMap* tm_a;     // [m_cap, k] activations
  const CUtensorMap* tm_b;     // [n, k] weight, box rows = bn
  int bn;
  const __nv_bfloat16* a;      // raw pointers for the SIMT checker
  long long lda;
  const __nv_bfloat16* b;
  long long ldb;
  int m, n, k;
  int epi;                     // LinearEpi
  __nv_bfloat16* out;
  long long ldo;
  const float* bias;           // [n] or nullptr
  const __nv_bfloat16* residual;  // LIN_RESIDUAL: [m, ldo]
  long long head_rows;         // LIN_QKV: token capacity of one (section, head) plane of the head-major output
  const long long* row_map;    // LIN_PLAIN only, optional: device array, output row of each GEMM row (fused scatter)
  // LayerNorm folded into this Linear (LIN_GELU / LIN_QKV, bf16 only): A is the raw residual stream, b the gamma-scaled weight,
  // bias the beta-folded bias; the epilogue applies (acc - mean * colsum) * rstd
  const float2* ln_stats;      // [m] (mean, rstd) or nullptr
  const float* ln_colsum;      // [n]
  const float2* ln_part;       // alternative to ln_stats: [m][k / 32] per-panel (sum, sum of squares) left by the producer's RowStats<>
                               // epilogue; the kernel's idle warps finalise them per tile
  // LIN_RESIDUAL, bf16 only, optional: also leave per-row / per-32-column-panel (sum, sum of squares) of the stored values
  float2* stats_part;          // [m][n / 32] or nullptr
  // ... or accumulate them (LIN_RESIDUAL) / read them (LIN_GELU / LIN_QKV with ln_colsum) as one fixed-point pair per row
  unsigned long long* stats_acc;        // [m][2] written with integer atomics (zeroed by the caller) or nullptr
  const unsigned long long* ln_acc;     // [m][2] alternative to ln_stats / ln_part
  // small-M path (m <= 128, bf16): the same weight with a box of bn_small = 32 rows: the small-M kernel (tc_gemm_small.cuh) for
  // k <= 1024, 64-column tiles of the pair kernel beyond; nullptr = not offered
  const CUtensorMap* tm_b_small;
  int bn_small;
  // FP8 variant (QUANTIZE=fp8): tm_a / tm_b are e4m3 maps (make_tmap_rowmajor_u8), k counts e4m3 elements,
  // y = acc * row_scale[m] * col_scale[n] + bias
  int fp8;
  const float* row_scale;      // [m]
  const float* col_scale;      // [n]
};
cudaError_t gemm_linear(const LinearArgs& a, bool simt, int num_sms, cudaStream_t stream);

struct ConvArgs {
  const CUtensorMap* tm_a;     // make_tmap_conv over the input activation
  const CUtensorMap* tm_b;     // [480, 9*512] packed weight, box rows = 240
  const __nv_bfloat16* a;      // input activation [g_in][h_in][c]
  long long g_in;
  int h_in;
  const __nv_bfloat16* b;      // [c][9*512]
  int n_chunks;
  int hc;                      // output rows per column (32 / 16)
  int gt;                      // output columns per 128-row tile (4 / 8)
  int slots;                   // output column slots per chunk (26 / 13)
  int max_w;                   // 25 / 13
  int out_pitch, out_off;
  const int* width;            // [n_chunks] valid output columns
  const float* bias;           // [c]
  __nv_bfloat16* out;
  int c;                       // 480
};
cudaError_t gemm_conv(const ConvArgs& a, bool simt, int num_sms, cudaStream_t stream);

struct ConvOutArgs {
  const CUtensorMap* tm_a;     // [chunks*13, 7680]
  const CUtensorMap* tm_b;     // [d, 7680] (K axis permuted to f*480+c), box rows = bn
  int bn;
  const __nv_bfloat16* a;
  const __nv_bfloat16* b;
  int m, d, k;
  const float* pe;             // [13, d]
  const int* row_token;        // [m]
  int tok_per_chunk;
  __nv_bfloat16* out;          // [tokens, d]
  int fp8;                     // as in LinearArgs
  const float* row_scale;
  const float* col_scale;
  float2* stats_part;          // optional (bf16 only): [tokens][d / 32] partial LayerNorm statistics of the rows written
  unsigned long long* stats_acc;   // optional (bf16 only): [tokens][2] fixed-point (sum, sum of squares), integer atomics
};
cudaError_t gemm_conv_out(const ConvOutArgs& a, bool simt, int num_sms, cudaStream_t stream);

}  // namespace qasr

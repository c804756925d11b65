// Host-callable launch wrappers for the qasr_b200 kernels.  Every wrapper enqueues on `stream`
// and returns a cudaError_t from the launch; none of them synchronises.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <vector>

#include <cuda.h>

#include "mel.cuh"

namespace qasr {

// ---- log-mel (mel.cu) ----------------------------------------------------------------------
// build_mel_tables returns false if the computed filter-bank structure differs from the compiled-in one
// (mel_structure.inc); upload_mel_constants copies the mel weights into constant memory of the current device.
bool build_mel_tables(mel::Tables* host_tables);
cudaError_t upload_mel_constants(const mel::Tables* host_tables);
// One persistent kernel over `items` (mel::Item, device).  counters: 1 + 2 * n_clips uint32 (ticket, per-clip done, per-clip max),
// zeroed by the launcher on `stream`.
// variant: 1 = the CTA-synchronous kernel of round 1 (three barriers per item, bulk-copied PCM slabs), 2 = warp-synchronous FFT stages
cudaError_t launch_logmel(const float* pcm, const mel::Item* items, int n_items, const mel::Tables* tables, float* mel_out,
                          long long mel_ld, unsigned int* counters, int n_clips, int num_sms, cudaStream_t stream);
// v3 (default): ticketed, warp-autonomous (no CTA barrier); its own item format (mel::Item3), counters of mel::counter_words(n_clips)
cudaError_t launch_logmel_v3(const float* pcm, const mel::Item3* items, int n_items, const mel::Tables* tables, const mel::Tables* host_tables,
                             float* mel_out, long long mel_ld, unsigned int* counters, int n_clips, int num_sms, cudaStream_t stream);

// ---- WS pre-frontend (prefrontend.cu) ---------------------------------------------------------
struct WsStream {       // one WS window
  long long in_off;     // first int16 sample in the packed input
  long long in_len;     // int16 samples present
  long long flt_len;    // in_len + flush silence: the span the band-pass runs over
  long long out_off;    // first float sample in the packed output
  long long out_len;    // max(flt_len, minimum clip length): [flt_len, out_len) is zero
};
struct RsStream {
  long long in_off, in_len, out_off, out_len;
};
void design_resample_taps(int up, int down, std::vector<double>* taps, int* half_len);
int sos_warmup(const double* sos, int n_sections);  // < 0: filter not (safely) stable
cudaError_t launch_ws_window(const int16_t* pcm16, const WsStream* streams_dev, int n_streams, long long max_out_len, const double* sos,
                             int n_sections, int warm, float* out, cudaStream_t stream);
cudaError_t launch_resample_pcm16(const int16_t* in, const RsStream* streams_dev, int n_streams, long long max_out_len,
                                  const double* taps_dev, int n_taps, int half_len, int up, int down, int16_t* out, int num_sms,
                                  cudaStream_t stream);

// float32 (interleaved channels) -> mono float32 at the new rate, torchaudio sinc_interp_hann (prefrontend.cu)
void design_sinc_hann_kernel(int orig, int newf, std::vector<float>* kernel /*[newf][2 * width + orig]*/, int* width);
cudaError_t launch_resample_f32(const float* in, const RsStream* streams_dev, int n_streams, long long max_out_len, int channels,
                                const float* taps_dev, const int* lo_dev, int n_keep, int width, int orig, int newf, float* out,
                                int num_sms, cudaStream_t stream);

// long-audio splitter: quietest sample of the quietest `win`-sample window inside [left, right) -> *boundary_out_dev
cudaError_t launch_split_scan(const float* x, long long left, long long right, int win, long long* boundary_out_dev, cudaStream_t stream);

// ---- conv1 (elementwise.cu) ------------------------------------------------------------------
struct ChunkDesc {
  long long mel_col0;  // first mel column of the chunk in the packed mel
  int valid;           // valid mel frames in the chunk
  int w1;              // valid conv1 output columns (conv_len(padded frames))
};
constexpr int ACT1_PITCH = 52;  // columns per chunk in conv1's output: [zero][50][zero]
constexpr int ACT1_H = 64;
cudaError_t launch_conv1(const void* mel, int mel_is_bf16, long long mel_ld, const ChunkDesc* chunks, int n_chunks,
                         const float* w /*[C][9]*/, const float* bias /*[C]*/, const float* gelu_lut /*[gelu_tab::WORDS], device*/, int channels,
                         __nv_bfloat16* act1, bool simt /*fp32 FFMA checker instead of the mma.sync kernel*/, cudaStream_t stream);

// ---- FP8 dynamic quantisation (quant.cu) -----------------------------------------------------
// x: bf16 [rows, k] (ld ldx) -> q: e4m3 [rows, k] (ld ldq bytes) + row_scale[rows].  per_row = false: one scale for
// the whole tensor (amax_slot: a zeroed uint32 on the device), written to every row_scale entry.
cudaError_t launch_quant_fp8(const __nv_bfloat16* x, long long ldx, int rows, int k, uint8_t* q, long long ldq, float* row_scale,
                             unsigned int* amax_slot, bool per_row, int num_sms, cudaStream_t stream);
void quantize_weight_e4m3(const float* w, int n, int k, int modules, bool per_row, uint8_t* q_out, float* scale_out);

// ---- LayerNorm -------------------------------------------------------------------------------
cudaError_t launch_layernorm(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* out,
                             int rows, int d, float eps, cudaStream_t stream);

// (mean, rstd) per row only: for the Linears that have the LayerNorm folded in (epilogues.cuh LnFold)
cudaError_t launch_ln_stats(const __nv_bfloat16* x, float2* stats, int rows, int d, float eps, cudaStream_t stream);

// LayerNorm whose bf16-rounded output row is quantised to e4m3 with a per-row dynamic scale in the same pass (fp8 per-row mode)
cudaError_t launch_layernorm_fp8(const __nv_bfloat16* x, const float* gamma, const float* beta, uint8_t* q_out, float* row_scale,
                                 int rows, int d, float eps, cudaStream_t stream);

// ---- windowed attention ----------------------------------------------------------------------
// qkv: head-major [3 (q, k, v)][heads][head_rows tokens][64] (written by the QKV GEMM's EpiQkv epilogue); out: [tokens, d];
// win: [n_win] (start, len)
cudaError_t launch_window_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* win, int n_win,
                                    int max_win_len, int d, int heads, int head_rows, cudaStream_t stream);

// tcgen05 version (attention_tc.cu): windows up to 128 tokens; tm_qkv = 128B-swizzled map over the same buffer viewed as
// [3 * heads * head_rows, 64], box 64 x attention_tc_tile_rows(longest window) -- pass the same value as tile_rows
int attention_tc_tile_rows(int max_win_len);
cudaError_t launch_window_attention_tc(const CUtensorMap* tm_qkv, int tile_rows, __nv_bfloat16* out, const int2* win, int n_win,
                                       int max_win_len, int d, int heads, int head_rows, int num_sms, cudaStream_t stream);

}  // namespace qasr

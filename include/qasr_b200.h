/* qasr_b200 -- C ABI of the B200-native log-mel frontend + Qwen3-ASR audio-encoder backend.
 *
 * This is the drop-in boundary for the encoder-backend slot of jaaacki/qwen3-asr
 * (reference src/server.py).  The reference binds its two existing backends from Python:
 *   - TRT slot : loader src/server.py:237-251, dispatch src/server.py:873-893
 *   - ONNX slot: loader src/server.py:461-475, dispatch src/server.py:895-914
 * and both replace one call: encoder.forward(input_features) -> hidden states.  The real module
 * behind that call is m.model.thinker.audio_tower, whose forward is
 * Qwen3OmniMoeAudioEncoder.forward(input_features[128, sum T], feature_lens[B])
 * (transformers modeling_qwen3_omni_moe.py:698-766), fed by WhisperFeatureExtractor
 * (feature_extraction_whisper.py:135-164, 189-342).  Each entry point below names the piece of
 * that interface it replaces.  INTEGRATION.md shows the ctypes binding and the server.py patch.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; qasr_last_error() then holds a
 *     message (thread-local).  There is no CPU fallback inside the library.
 *   - device pointers are plain CUDA device addresses on the handle's device (torch tensors are
 *     passed as tensor.data_ptr()); "stream" is a cudaStream_t passed as void* (0 = default
 *     stream; pass torch.cuda.current_stream().cuda_stream).  Work is enqueued on that stream and
 *     the call returns without synchronising, except the *_host entry points, which synchronise
 *     the stream before returning because they hand back host data.
 *   - a handle may be used by one host thread at a time (the reference has exactly one inference
 *     thread, src/server.py:45-48).
 */
#ifndef QASR_B200_H_
#define QASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QASR_ABI_VERSION 1

#if defined(__GNUC__)
#define QASR_API __attribute__((visibility("default")))
#else
#define QASR_API
#endif

typedef struct qasr_handle_s* qasr_handle_t;

/* qasr_config_t.flags.  QASR_FLAG_FP8 mirrors the reference's QUANTIZE=fp8 (src/server.py:362-371: torchao
 * Float8DynamicActivationFloat8WeightConfig on every nn.Linear, convolutions untouched): e4m3 weights, activations
 * quantised dynamically before every Linear, fp32 accumulate, bf16 out.  Default granularity is per-tensor (one scale
 * per weight module and per activation tensor of the call); QASR_FLAG_FP8_PER_ROW switches to per-output-channel weight
 * scales and per-token activation scales (torchao's PerRow), which is invariant to how clips are batched. */
#define QASR_FLAG_FP8 1
#define QASR_FLAG_FP8_PER_ROW 2

/* dtypes for qasr_set_weight / qasr_encode */
#define QASR_F32 0
#define QASR_BF16 1
#define QASR_F16 2

/* Fields of thinker_config.audio_config (vllm transformers_utils/configs/qwen3_asr.py:27-130);
 * values per checkpoint in SURVEY.md appendix A.1.  All run-time, none compiled in. */
typedef struct qasr_config_s {
  int32_t d_model;              /* 1024 (1.7B) / 896 (0.6B) */
  int32_t encoder_layers;       /* 24 / 18 */
  int32_t encoder_attention_heads; /* 16 / 14 (head_dim must be 64) */
  int32_t encoder_ffn_dim;      /* 4096 / 3584 */
  int32_t output_dim;           /* 2048 / 1024 */
  int32_t n_window;             /* 50  -> 100-frame conv chunks (must be 50) */
  int32_t n_window_infer;       /* 800 -> 104-token attention windows */
  int32_t downsample_hidden_size; /* 480 (must be 480) */
  int32_t num_mel_bins;         /* 128 (must be 128) */
  int32_t max_source_positions; /* 1500 */
  /* capacity of one internal micro-batch; larger requests are split (exactly) into several.
   * 0 = defaults (1024 chunks = 1024 audio-seconds, 13312 tokens) */
  int32_t max_chunks;
  int32_t max_tokens;
  int32_t flags;                /* QASR_FLAG_* bits, 0 = bf16 */
} qasr_config_t;

QASR_API int qasr_abi_version(void);
QASR_API const char* qasr_last_error(void);

/* Create a backend instance on CUDA device `device` (replaces _try_load_trt_encoder /
 * _try_load_onnx_encoder, src/server.py:237-251, 461-475). */
QASR_API int qasr_create(const qasr_config_t* cfg, int device, qasr_handle_t* out);

/* Hand over one parameter of the audio tower by its state_dict name without the
 * "thinker.audio_tower." prefix (names and shapes: SURVEY.md appendix A.5), e.g.
 * "layers.3.self_attn.q_proj.weight".  `data` may be a host or a device pointer; the library
 * copies.  Optional extra name: "positional_embedding" [>=13, d_model] overrides the internally
 * generated sinusoid table (modeling_qwen3_omni_moe.py:88-106). */
QASR_API int qasr_set_weight(qasr_handle_t h, const char* name, const void* data, int dtype, const int64_t* shape, int ndim);

/* Pack weights for the kernels (fused QKV, conv weights to [out][tap][in], conv_out K-axis
 * permuted from c*16+f to f*480+c), upload, allocate the workspace.  Must be called once after
 * all qasr_set_weight calls and before any encode. */
QASR_API int qasr_finalize(qasr_handle_t h);

/* Device bytes held by the handle (weights + workspace). */
QASR_API size_t qasr_workspace_bytes(qasr_handle_t h);

/* Token count the encoder emits for a clip of `feature_len` mel frames
 * (_get_feat_extract_output_lengths, modeling_qwen3_omni_moe.py:145-153). */
QASR_API int64_t qasr_token_len(int64_t feature_len);

/* The table behind every GELU of the device path (host-only entry point, no GPU needed; csrc/common.cuh gelu_tab).  The encoder's
 * GELUs (ACT2FN["gelu"] = erf GELU: conv stem modeling_qwen3_omni_moe.py:693-700, fc1 :452, proj1 :760) take a value that was just
 * rounded to bf16 and round their result to bf16, i.e. they are maps between 16-bit patterns; the device evaluates
 * bf16(x * table[sign][clamp(|x| bits)]) and the table is built so that this IS the correctly rounded float64 erf GELU for all
 * 65 536 inputs.  Writes the 2 x 1666 float32 ratios (positive x, then negative x; |x| bits 0x3AFF..0x4180) to table_out and
 * returns the number of floats written, or a negative value if capacity is too small / the library's exhaustive self-check fails. */
QASR_API int qasr_gelu_table(float* table_out, int capacity);

/* The work list the log-mel kernel walks for a batch (host-only entry point, no GPU needed; csrc/mel.cuh Item3): item i = the i-th
 * 32-frame tile to transform, clip by clip, plus the 32-frame tile whose max-8 clamp it carries -- a tile of a clip whose last
 * frame item lies at least 8 * num_sms items earlier (FIFO), so that the clamp finds the clip maximum final and the tile still in
 * L2; tiles left over ride on trailing items without frames.  Writes 8 int64 per item (clip, frame0, n_frames, col0, clamp clip,
 * clamp frame0, clamp n_frames, frame items of the clamp's clip) when items_out is not NULL; returns the item count, -1 on a bad
 * argument.  Replaces nothing in the reference (WhisperFeatureExtractor clamps after the whole clip, feature_extraction_whisper.py:
 * 160-163); it exists so that the scheduling invariants of the kernel are testable on the CPU. */
QASR_API int64_t qasr_mel_plan(const int64_t* clip_offsets, int n_clips, int num_sms, int64_t* items_out, int64_t capacity_items);

/* Log-mel of n_clips mono 16 kHz float32 clips packed back to back on the device
 * (replaces WhisperFeatureExtractor._torch_extract_fbank_features per clip, standalone semantics).
 *   pcm_dev            float32, clip i = samples [clip_offsets[i], clip_offsets[i+1])  (host int64 [n_clips+1])
 *   mel_out_dev        float32 [128, mel_ld], clip i occupies columns [sum_{j<i} T_j, +T_i), T_i = N_i / 160
 *   feature_lens_out   host int64 [n_clips] (may be NULL)
 * Every clip needs more than 200 samples (as torch.stft's reflect padding does). */
QASR_API int qasr_logmel(qasr_handle_t h, const float* pcm_dev, const int64_t* clip_offsets, int n_clips, float* mel_out_dev,
                int64_t mel_ld, int64_t* feature_lens_out, void* stream);

/* Audio-tower forward (replaces audio_tower.forward / the patched encoder.forward of
 * src/server.py:873-914).
 *   mel_dev            [128, mel_ld] packed features, dtype QASR_F32 or QASR_BF16; clip i occupies
 *                      columns [sum_{j<i} feature_lens[j], +feature_lens[i])
 *   feature_lens       host int64 [n_clips]
 *   out_dev            bf16 [sum tokens, output_dim], clip-major
 *   token_lens_out     host int64 [n_clips] (may be NULL) */
QASR_API int qasr_encode(qasr_handle_t h, const void* mel_dev, int mel_dtype, int64_t mel_ld, const int64_t* feature_lens, int n_clips,
                void* out_dev, int64_t* token_lens_out, void* stream);

/* qasr_encode whose last GEMM writes every encoder token straight into its row of the decoder's input embeddings: the
 * fused form of `inputs_embeds.masked_scatter(audio_mask, audio_features)` (transformers modeling_qwen3_omni_moe.py:2135-2143,
 * SURVEY.md section 8 row a15 / 8f-3) -- the [tokens, output_dim] intermediate and the scatter pass disappear.
 *   embeds_dev         bf16 [rows, embeds_ld], embeds_ld >= output_dim (the decoder's hidden size), 16-byte aligned
 *   token_rows_dev     DEVICE int64 [sum tokens]: row of embeds_dev for each encoder token, in clip-major token order
 *                      (= the flattened positions where audio_mask is true) */
QASR_API int qasr_encode_scatter(qasr_handle_t h, const void* mel_dev, int mel_dtype, int64_t mel_ld, const int64_t* feature_lens, int n_clips,
                                 void* embeds_dev, int64_t embeds_ld, const int64_t* token_rows_dev, int64_t* token_lens_out, void* stream);

/* Fused PCM -> log-mel -> encoder on the device (the mel stays in the handle's workspace). */
QASR_API int qasr_encode_pcm(qasr_handle_t h, const float* pcm_dev, const int64_t* clip_offsets, int n_clips, void* out_dev,
                    int64_t* token_lens_out, void* stream);

/* Same, end to end with HOST buffers: copies pcm_host (pinned memory recommended) to the device,
 * encodes, copies the bf16 hidden states back into out_host and synchronises `stream`.
 *   out_host           bf16 [sum tokens, output_dim]; out_capacity_tokens bounds the copy. */
QASR_API int qasr_encode_pcm_host(qasr_handle_t h, const float* pcm_host, const int64_t* clip_offsets, int n_clips, void* out_host,
                         int64_t out_capacity_tokens, int64_t* token_lens_out, void* stream);

/* Pipelined form of qasr_encode_pcm_host for back-to-back batches (the serving loop): returns as soon as the
 * work is enqueued -- host->device copy on an internal copy stream, log-mel + encoder on `stream`, device->host copy
 * on a second internal stream -- so that the copies of one batch overlap the compute of its neighbours (device
 * buffers are double-buffered).  qasr_wait(ticket) blocks until out_host holds that batch.  pcm_host and out_host
 * must stay valid (and should be pinned) until then.  At most two submits may be un-waited at any time. */
QASR_API int qasr_submit_pcm_host(qasr_handle_t h, const float* pcm_host, const int64_t* clip_offsets, int n_clips, void* out_host,
                         int64_t out_capacity_tokens, int64_t* token_lens_out, void* stream, uint64_t* ticket_out);
QASR_API int qasr_wait(qasr_handle_t h, uint64_t ticket);
/* Non-blocking form of qasr_wait: *done_out = 1 once out_host holds the batch of `ticket` (a qasr_wait would return at once), else 0.
 * What an event loop -- or the pool's worker threads -- call between other work instead of parking a thread in qasr_wait. */
QASR_API int qasr_poll(qasr_handle_t h, uint64_t ticket, int* done_out);
/* Device-side durations of the three legs of a finished ticket (CUDA events on the copy-in, compute and copy-out streams): host->device
 * copy, log-mel + encoder, device->host copy.  Valid after qasr_wait(ticket) until the same slot's next submit (ticket + 2). */
QASR_API int qasr_pipe_times(qasr_handle_t h, uint64_t ticket, float* h2d_ms, float* compute_ms, float* d2h_ms);
/* qasr_submit_pcm_host for clips that are NOT contiguous in pcm_host (what an LPT-sharded pool member receives): clip i is
 * pcm_host[clip_begin[i], clip_begin[i] + clip_len[i]) and its tokens are written at row out_row[i] of out_host (bf16
 * [out_capacity_tokens, output_dim]); one copy per clip each way instead of one per batch.  Same ticket rules. */
QASR_API int qasr_submit_clips_host(qasr_handle_t h, const float* pcm_host, const int64_t* clip_begin, const int64_t* clip_len,
                                    const int64_t* out_row, int n_clips, void* out_host, int64_t out_capacity_tokens, int64_t* token_lens_out,
                                    void* stream, uint64_t* ticket_out);

/* Log-mel end to end with host buffers (float32 [128, sum T] out). */
QASR_API int qasr_logmel_host(qasr_handle_t h, const float* pcm_host, const int64_t* clip_offsets, int n_clips, float* mel_out_host,
                     int64_t* feature_lens_out, void* stream);

/* ---- WebSocket pre-frontend: the per-window CPU passes of the reference's
 * WS path, batched over streams on the device.  Streams are packed back to back; offsets are HOST int64 [n_streams + 1]. */

/* Output length of qasr_resample_pcm16 for an input of n_in samples: ceil(n_in * up / down). */
QASR_API int64_t qasr_resample_len(int64_t n_in, int up, int down);

/* int16 -> resample by up/down -> int16 (replaces _resample_pcm_bytes, src/server.py:32-42: astype(float32) ->
 * librosa.resample -> astype(int16), i.e. truncation toward zero; this library saturates where numpy would wrap).
 * The resampler is the polyphase Kaiser-windowed sinc of scipy.signal.resample_poly (librosa's soxr_hq is not
 * reproducible offline: parity unpinned against the reference, pinned against scipy), evaluated in float64.
 *   taps               host float64 [n_taps], n_taps odd, already scaled by `up` (scipy: firwin(...) * up), or NULL for
 *                      the built-in design firwin(20 * max(up, down) + 1, 1 / max(up, down), ("kaiser", 5.0)) * up
 *   out_dev            int16, stream i at [out_offsets_out[i], out_offsets_out[i + 1]); out_capacity in samples
 *   out_offsets_out    host int64 [n_streams + 1], written by the call */
QASR_API int qasr_resample_pcm16(qasr_handle_t h, const int16_t* pcm16_dev, const int64_t* in_offsets, int n_streams, int up, int down,
                                 const double* taps, int n_taps, int16_t* out_dev, int64_t out_capacity, int64_t* out_offsets_out,
                                 void* stream);

/* Upload normalisation of the HTTP path (src/server.py:867 hands (audio, sr) of any rate / channel count to the SDK, which turns it
 * into mono float32 at 16 kHz): float32 frames with `channels` interleaved channels at orig_sr -> mono float32 at new_sr.  Defined as
 * the reference's own spelling of that step in src/debug_audio.py:24-33: audio.mean(axis=1), then
 * torchaudio.functional.resample(waveform, orig_sr, new_sr) = sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99.  The SDK's
 * internal resampler is not available offline: pinned against torchaudio, unpinned against the SDK.
 *   in_offsets         host int64 [n_streams + 1] in FRAMES (stream i = frames [in_offsets[i], in_offsets[i+1]), each `channels` floats)
 *   taps               host float32 [new_sr / g][2 * width + orig_sr / g] (g = gcd; torchaudio's _get_sinc_resample_kernel layout), or
 *                      NULL for the built-in design (the same float32 recipe, see qasr_resample_f32_taps)
 *   out_dev            float32, stream i at [out_offsets_out[i], out_offsets_out[i + 1]), length ceil(n * new_sr / orig_sr) -- directly
 *                      usable as the pcm_dev / clip_offsets of qasr_logmel and qasr_encode_pcm; out_capacity in samples */
QASR_API int64_t qasr_resample_f32_len(int64_t n_in_frames, int orig_sr, int new_sr);
QASR_API int qasr_resample_f32(qasr_handle_t h, const float* pcm_dev, const int64_t* in_offsets, int n_streams, int channels, int orig_sr,
                               int new_sr, const float* taps, float* out_dev, int64_t out_capacity, int64_t* out_offsets_out, void* stream);
/* The built-in resampling kernel on its own (pure host code, no GPU needed): writes [*n_phases][*n_taps] float32 into taps_out
 * (may be NULL to query the sizes); *width = the zero padding torchaudio applies on the left. */
QASR_API int qasr_resample_f32_taps(int orig_sr, int new_sr, float* taps_out, int64_t capacity, int* n_phases, int* n_taps, int* width);

/* Long-audio splitter: the SDK cuts audio longer than max_chunk_sec (1200 s) into chunks at low-energy points
 * before encoding them as independent clips -- split_audio_into_chunks, which the reference relies on (LEARNING_LOG.md:215-219:
 * "sliding window convolution with +/-5s search range"; src/server.py:867 enters it through model.transcribe).  For every cut wanted
 * at start + max_chunk_sec * sr: inside [cut - expand, cut + expand) find the min_window_ms window with the smallest sum of |x|
 * and cut at its quietest sample (first one on ties).  Arithmetic is exact (|x| quantised to floor(|x| * 2^40), int64 sums), so the
 * boundaries are defined bit for bit.  The SDK's source is not available offline: PARITY UNPINNED against it, pinned against
 * oracle/prefrontend.py::split_points, which restates the description above.  Synchronises `stream` (every cut depends on the last).
 *   pcm_dev            mono float32 on the device, n_samples long
 *   boundaries_out     host int64 [capacity]: 0 = b_0 < b_1 < ... < b_n = n_samples; chunk i is [b_i, b_i+1) -- directly usable as
 *                      the clip_offsets of qasr_encode_pcm / qasr_pool_submit;  *n_chunks_out = n */
QASR_API int qasr_split_audio(qasr_handle_t h, const float* pcm_dev, int64_t n_samples, int sample_rate, double max_chunk_sec,
                              double search_expand_sec, double min_window_ms, int64_t* boundaries_out, int capacity, int* n_chunks_out,
                              void* stream);

/* One WS window per stream: int16 16 kHz samples (+ pad_samples[i] int16 zeros appended: the 600 ms flush silence)
 * -> float32 / 32768 -> SOS band-pass in float64, zero initial state, cast to float32 -> zero-padded to min_samples
 * (replaces _transcribe_with_context's numpy prologue, src/server.py:1321-1338, and _telephony_bandpass, :26-29 =
 * scipy.signal.sosfilt(butter(4, [300, 3400], "bandpass", fs=16000, output="sos"), audio).astype(float32)).
 *   pad_samples        host int32 [n_streams] or NULL
 *   sos                host float64 [n_sections, 6] in scipy's layout (b0 b1 b2 a0 a1 a2), n_sections <= 8; NULL / 0 = no filter
 *   min_samples        shorter windows are zero-padded to this length AFTER filtering (the SDK's 0.5 s minimum); 0 = off
 *   out_dev            float32, stream i at [out_offsets_out[i], out_offsets_out[i + 1]) -- directly usable as the
 *                      pcm_dev / clip_offsets of qasr_logmel and qasr_encode_pcm; out_capacity in samples */
QASR_API int qasr_ws_window(qasr_handle_t h, const int16_t* pcm16_dev, const int64_t* in_offsets, int n_streams, const int32_t* pad_samples,
                            const double* sos, int n_sections, int min_samples, float* out_dev, int64_t out_capacity,
                            int64_t* out_offsets_out, void* stream);

QASR_API void qasr_destroy(qasr_handle_t h);

/* ---- one process, several GPUs -----------------------------------------------------------
 * The path shards by clip with no exchange step, so the pool is N independent handles (weights replicated), one worker
 * thread + CUDA stream per device, and NO collective (NCCL is not used).  A batch is cut into at most N contiguous clip
 * ranges of near-equal work; every shard reads its slice of the caller's (pinned) PCM buffer and writes its slice of the
 * caller's output buffer directly, in clip order.  Stands where a multi-GPU deployment of the reference would run one
 * server process per GPU behind the gateway (src/gateway.py / src/worker.py). */
typedef struct qasr_pool_s* qasr_pool_t;
QASR_API int qasr_pool_create(const qasr_config_t* cfg, const int* devices, int n_devices, qasr_pool_t* out);
QASR_API int qasr_pool_size(qasr_pool_t p);
/* qasr_set_weight / qasr_finalize / qasr_workspace_bytes for every handle of the pool */
QASR_API int qasr_pool_set_weight(qasr_pool_t p, const char* name, const void* data, int dtype, const int64_t* shape, int ndim);
QASR_API int qasr_pool_finalize(qasr_pool_t p);
QASR_API size_t qasr_pool_workspace_bytes(qasr_pool_t p);
/* qasr_submit_pcm_host across the pool.  Returns at once; token_lens_out [n_clips] and (optionally) clip_device_out
 * [n_clips] (the CUDA device each clip was sent to) are filled before it returns and not touched afterwards (clip_offsets
 * is copied).  pcm_host and out_host must stay valid until qasr_pool_collect(ticket) returns; any number of batches may be in flight (each worker
 * keeps two on its GPU).  For throughput keep THREE batches in flight (collect the oldest, submit the next): with two, every GPU's next
 * shard is submitted only after the previous batch has been collected, i.e. one device->host copy late (measured on 8 B200s: 613 k
 * audio-s/s with three, 542 k with two).  out_host: bf16 [sum tokens, output_dim] in clip order. */
QASR_API int qasr_pool_submit(qasr_pool_t p, const float* pcm_host, const int64_t* clip_offsets, int n_clips, void* out_host,
                              int64_t out_capacity_tokens, int64_t* token_lens_out, int32_t* clip_device_out, uint64_t* ticket_out);
QASR_API int qasr_pool_collect(qasr_pool_t p, uint64_t ticket);
/* Per-member averages since creation (or the last call with reset != 0), QASR_POOL_STAT_FIELDS doubles per member:
 * shards, host ms spent enqueueing a shard, host ms blocked waiting for one, device ms of its host->device copy, of its compute, of
 * its device->host copy, host ms from enqueue start to completion.  Tells a launch-bound pool from a copy-bound one. */
#define QASR_POOL_STAT_FIELDS 7
QASR_API int qasr_pool_stats(qasr_pool_t p, double* out, int capacity_members, int reset);
/* How qasr_pool_submit cuts a batch.  CONTIGUOUS: at most N contiguous clip ranges of near-equal mel-frame
 * count -- every shard is one copy each way, in place.  LPT: longest-processing-time-first by mel frames (sort descending, each clip to
 * the least-loaded member) -- near-optimal balance when clips are few and unequal, at one copy per clip.  AUTO (default): LPT when the
 * batch has fewer than four clips per member, else CONTIGUOUS.  The output is in clip order either way. */
#define QASR_SHARD_AUTO 0
#define QASR_SHARD_CONTIGUOUS 1
#define QASR_SHARD_LPT 2
QASR_API int qasr_pool_set_sharding(qasr_pool_t p, int mode);
QASR_API int qasr_pool_plan_mode(const int64_t* clip_offsets, int n_clips, int n_devices, int mode, int32_t* clip_shard_out);
/* The CONTIGUOUS sharding rule on its own (pure host code, no GPU needed): clip_shard_out[i] = index (0 .. n_devices-1)
 * of the pool member clip i would be sent to -- contiguous ranges of near-equal mel-frame count. */
QASR_API int qasr_pool_plan(const int64_t* clip_offsets, int n_clips, int n_devices, int32_t* clip_shard_out);
QASR_API void qasr_pool_destroy(qasr_pool_t p);

/* ---- launch accounting and per-launch timing (measurement; bench.py's roofline figures) ------- */
/* Number of CUDA kernels this handle has launched so far. */
QASR_API uint64_t qasr_launch_count(qasr_handle_t h);
/* Graph cache: qasr_encode / qasr_encode_pcm calls whose shape (entry point, dtype, every clip length) repeats are captured into a
 * CUDA graph on the second sighting and replayed from then on (same kernels and arguments, entry-owned buffers: bit-identical
 * results, one launch instead of ~130-180).  Calls of more than 128 one-second chunks run eagerly unless QASR_GRAPH=all;
 * QASR_GRAPH=0 switches the cache off.  Number of live graphs, replays so far, device bytes they hold (counters the kernels
 * launched by a replay are included in qasr_launch_count). */
QASR_API int qasr_graph_stats(qasr_handle_t h, int* n_graphs, uint64_t* replays, size_t* bytes);
/* on != 0: bracket every kernel launch with CUDA events on the launching stream (clears earlier
 * records).  qasr_profile_read synchronises the device and returns, aggregated by kernel name in
 * first-launch order: '\n'-separated names, summed milliseconds, summed algorithmic work (FLOPs;
 * bytes for "logmel") and launch counts. */
QASR_API int qasr_profile_enable(qasr_handle_t h, int on);
QASR_API int qasr_profile_read(qasr_handle_t h, char* names, size_t names_cap, double* ms, double* work, int32_t* counts,
                               int max_entries, int* n_entries);

/* ---- test / bring-up hooks (not part of the serving path) ---------------------------------- */
/* Copy a named intermediate of the LAST encode call to the host (synchronises the device).
 * names: "act1","act2","act3","embed","mel".  Returns the number of bytes copied in *nbytes. */
QASR_API int qasr_debug_read(qasr_handle_t h, const char* name, void* dst_host, size_t capacity, size_t* nbytes);
/* D[M,N] = act(A[M,K] * B[N,K]^T + bias) (+ residual) through the encoder's tcgen05 GEMM
 * (impl 0) or the SIMT checker (impl 1).  All pointers are device pointers, A/B/D/residual bf16,
 * bias float32 or NULL; act: 0 none, 1 GELU. */
QASR_API int qasr_debug_gemm(const void* a, const void* b, const float* bias, const void* residual, void* d, int m, int n, int k, int act,
                    int impl, void* stream);
/* bf16 [rows, k] -> e4m3 [rows, k] + row_scale[rows] (per_row = 0: one tensor-wide scale replicated); synchronises. */
QASR_API int qasr_debug_quant_fp8(const void* x, int rows, int k, void* q_out, float* row_scale_out, int per_row, void* stream);
/* D[M,N] = act(A8[M,K] * B8[N,K]^T * row_scale[m] * col_scale[n] + bias) (+ residual), e4m3 operands, bf16 out. */
QASR_API int qasr_debug_gemm_fp8(const void* a8, const void* b8, const float* row_scale, const float* col_scale, const float* bias,
                        const void* residual, void* d, int m, int n, int k, int act, void* stream);
QASR_API int qasr_debug_layernorm(const void* x, const float* gamma, const float* beta, void* out, int rows, int d, void* stream);
QASR_API int qasr_debug_attention(const void* qkv, void* out, const int32_t* win_start_len_host, int n_win, int d, int heads, void* stream);

/* The tcgen05 attention kernel (windows <= 128 tokens); qkv has `tokens` rows. */
QASR_API int qasr_debug_attention_tc(const void* qkv, void* out, const int32_t* win_start_len_host, int n_win, int tokens, int d, int heads,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QASR_B200_H_ */

// WebSocket pre-frontend on the device: what the reference does to a WS audio
// window on the CPU before the log-mel.
//   resample_pcm16_kernel  <- src/server.py:32-42  _resample_pcm_bytes: int16 -> polyphase FIR resample -> astype(int16)
//   resample_f32_kernel    <- the HTTP upload path's normalisation (src/server.py:867 -> SDK: mono mean, resample to 16 kHz,
//                             float32), in the definition the reference itself spells out in src/debug_audio.py:24-33:
//                             audio.mean(axis=1) -> torchaudio.functional.resample(sinc_interp_hann, width 6, rolloff 0.99)
//   ws_window_kernel       <- src/server.py:1321-1338 + :26-29: [window] (+ flush silence) -> /32768 -> 300-3400 Hz
//                             Butterworth SOS cascade (scipy sosfilt, float64, direct form II transposed) -> float32
// Both are HBM-trivial (2-6 bytes per sample); the work is the float64 recursion, which is sequential in time.  It is
// parallelised by cutting every stream into blocks of SOS_L samples, one thread per block, each thread first running the
// cascade from zero state over `warm` samples before its block: the band-pass poles lie at radius <= 0.961, so after
// warm = 1184 samples (pole radius -> 1e-18, plus a margin per section) the state error is < 1e-18 of the signal -- far below the float32 the result is cast to.  Blocks
// closer than `warm` to the stream start begin at sample 0 with the true zero state.
// The arithmetic mirrors scipy's _sosfilt operation by operation with separately rounded multiplies and adds (no FMA
// contraction), and the four sections are skewed across iterations (section s works on sample t - s) so that the four
// recurrences of one iteration are independent instruction chains.
#include <climits>
#include <cmath>
#include <vector>

#include "kernels.h"
#include "common.cuh"

namespace qasr {
namespace {

constexpr int SOS_L = 512;         // output samples per thread
constexpr int SOS_THREADS = 64;
constexpr int MAX_SECTIONS = 8;

struct SosCoef {
  double b0[MAX_SECTIONS], b1[MAX_SECTIONS], b2[MAX_SECTIONS], a1[MAX_SECTIONS], a2[MAX_SECTIONS];
};

// scipy _sosfilt_float inner statement for one section: returns y, updates the two states
__device__ __forceinline__ double sos_step(double x, double b0, double b1, double b2, double a1, double a2, double& s0, double& s1) {
  const double y = __dadd_rn(__dmul_rn(b0, x), s0);
  s0 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x), __dmul_rn(a1, y)), s1);
  s1 = __dsub_rn(__dmul_rn(b2, x), __dmul_rn(a2, y));
  return y;
}

template <int NS>
__global__ void __launch_bounds__(SOS_THREADS) ws_window_kernel(const int16_t* __restrict__ pcm16, const WsStream* __restrict__ streams,
                                                               const SosCoef coef, int warm, float* __restrict__ out) {
  const WsStream st = streams[blockIdx.y];
  const long long blk = static_cast<long long>(blockIdx.x) * SOS_THREADS + threadIdx.x;
  const long long n0 = blk * SOS_L;
  if (n0 >= st.out_len) return;
  const long long n1 = min(n0 + SOS_L, st.out_len);
  const int16_t* __restrict__ in = pcm16 + st.in_off;
  float* __restrict__ o = out + st.out_off;
  auto sample = [&](long long n) -> double {
    // int16 -> float32 / 32768 (exact in float32: a power-of-two scale), then scipy promotes to float64; the flush
    // silence [in_len, flt_len) is int16 zeros
    return n < st.in_len ? static_cast<double>(static_cast<float>(in[n]) * (1.0f / 32768.0f)) : 0.0;
  };
  if constexpr (NS == 0) {
    for (long long n = n0; n < n1; ++n) o[n] = n < st.flt_len ? static_cast<float>(sample(n)) : 0.f;
    return;
  } else {
    const long long f1 = min(n1, st.flt_len);           // filtered region of this block ends here; the rest is the minimum-length zero pad
    const long long start = max(0LL, n0 - warm);
    double s0[NS], s1[NS], carry[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { s0[s] = 0.0; s1[s] = 0.0; carry[s] = 0.0; }
    // iteration t: section s processes sample t - s (its input is what section s - 1 produced one iteration earlier)
    const long long t_end = f1 + NS - 1;
    for (long long t = start; t < t_end; ++t) {
      double xin[NS];
      xin[0] = t < f1 ? sample(t) : 0.0;
#pragma unroll
      for (int s = 1; s < NS; ++s) xin[s] = carry[s - 1];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const long long n = t - s;                       // the sample section s works on
        if (n >= start && n < f1) carry[s] = sos_step(xin[s], coef.b0[s], coef.b1[s], coef.b2[s], coef.a1[s], coef.a2[s], s0[s], s1[s]);
      }
      const long long n_out = t - (NS - 1);
      if (n_out >= n0 && n_out < f1) o[n_out] = static_cast<float>(carry[NS - 1]);
    }
    for (long long n = max(n0, st.flt_len); n < n1; ++n) o[n] = 0.f;
  }
}

// y[n] = sum_j h[j] x_up[n * down + half_len - j] over the taps j = (n * down + half_len) mod up, += up  (ascending j, the
// oracle's order), float64, then numpy's astype(int16): truncation toward zero (saturating).
__global__ void __launch_bounds__(256) resample_pcm16_kernel(const int16_t* __restrict__ in, const RsStream* __restrict__ streams,
                                                             const double* __restrict__ taps, int n_taps, int half_len, int up, int down,
                                                             int16_t* __restrict__ out) {
  const RsStream st = streams[blockIdx.y];
  const int16_t* __restrict__ x = in + st.in_off;
  int16_t* __restrict__ y = out + st.out_off;
  for (long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < st.out_len;
       n += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m0 = n * down + half_len;
    double acc = 0.0;
    for (long long j = m0 % up; j < n_taps; j += up) {
      const long long i = (m0 - j) / up;
      if (i >= 0 && i < st.in_len) acc = __dadd_rn(acc, __dmul_rn(taps[j], static_cast<double>(x[i])));
    }
    const double t = trunc(acc);
    y[n] = static_cast<int16_t>(fmin(fmax(t, -32768.0), 32767.0));
  }
}

// torchaudio _apply_sinc_resample_kernel (functional.py): y[j * new + p] = sum_k kernel[p][k] * xpad[j * orig + k], xpad = x padded
// by `width` zeros on the left and width + orig on the right, i.e. xpad[i] = x[i - width].  Only the taps inside the Hann window's
// support are kept per phase (lo[p] .. lo[p] + n_keep): outside it torchaudio's float32 kernel holds cos^2 of a rounded pi / 2
// times the sinc, |tap| < 1e-20.  Products of float32 values are exact in float64; the sum is accumulated in float64 and rounded
// once (torch's conv1d accumulates in float32 in an unspecified order: its result scatters around this one by ~1e-7).
// channels > 1: the mono sample is float32(float64 sum over the interleaved channels / channels) -- numpy's audio.mean(axis=1)
// on the float64 array soundfile returns, followed by the .float() cast (debug_audio.py:24-31).
__global__ void __launch_bounds__(256) resample_f32_kernel(const float* __restrict__ in, const RsStream* __restrict__ streams, int channels,
                                                           const float* __restrict__ taps, const int* __restrict__ lo, int n_keep,
                                                           int width, int orig, int newf, float* __restrict__ out) {
  const RsStream st = streams[blockIdx.y];
  const float* __restrict__ x = in + st.in_off * channels;
  float* __restrict__ y = out + st.out_off;
  const double inv_c = 1.0 / channels;
  for (long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < st.out_len;
       n += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long j = n / newf;
    const int p = static_cast<int>(n - j * newf);
    const float* __restrict__ tp = taps + static_cast<long long>(p) * n_keep;
    const long long i0 = j * orig + lo[p] - width;   // input frame of the first kept tap
    double acc = 0.0;
    for (int k = 0; k < n_keep; ++k) {
      const long long i = i0 + k;
      if (i < 0 || i >= st.in_len) continue;
      float xv;
      if (channels == 1) {
        xv = __ldg(x + i);
      } else {
        double s = 0.0;
        for (int c = 0; c < channels; ++c) s = __dadd_rn(s, static_cast<double>(__ldg(x + i * channels + c)));
        xv = static_cast<float>(s * inv_c);
      }
      acc = fma(static_cast<double>(__ldg(tp + k)), static_cast<double>(xv), acc);
    }
    y[n] = static_cast<float>(acc);
  }
}

// orig == new: torchaudio returns the waveform unchanged; only the channel mean remains
__global__ void __launch_bounds__(256) downmix_f32_kernel(const float* __restrict__ in, const RsStream* __restrict__ streams, int channels,
                                                          float* __restrict__ out) {
  const RsStream st = streams[blockIdx.y];
  const float* __restrict__ x = in + st.in_off * channels;
  float* __restrict__ y = out + st.out_off;
  const double inv_c = 1.0 / channels;
  for (long long n = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; n < st.out_len;
       n += static_cast<long long>(gridDim.x) * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < channels; ++c) s = __dadd_rn(s, static_cast<double>(__ldg(x + n * channels + c)));
    y[n] = channels == 1 ? __ldg(x + n) : static_cast<float>(s * inv_c);
  }
}

// ---- long-audio splitter ---------------------------------------------------------------------------
// The SDK cuts audio longer than 1200 s at low-energy points before it is encoded (split_audio_into_chunks, described in the
// reference's LEARNING_LOG.md:215-219: "sliding window convolution with +/-5 s search range"): for a cut wanted at sample c, look
// at [c - expand, c + expand), take the window of `win` samples (100 ms) with the smallest sum of |x|, and cut at the quietest
// sample inside it.  Restated here with EXACT integer arithmetic so that the result is defined bit for bit: |x| is quantised to
// floor(|x| * 2^40) (exact for float32 inputs of magnitude >= 2^-16 at 24 significant bits; smaller samples lose their low bits
// identically everywhere), window sums are int64, ties go to the first position (numpy argmin).  One CTA per cut: every thread
// slides its own run of consecutive windows, then a block-wide argmin over (sum, position), then over the samples of the winner.
constexpr int SPLIT_THREADS = 1024;

__device__ __forceinline__ long long split_quant(float x) { return static_cast<long long>(floor(fabs(static_cast<double>(x)) * 1099511627776.0)); }

__global__ void __launch_bounds__(SPLIT_THREADS) split_scan_kernel(const float* __restrict__ x, long long left, long long right, int win,
                                                                    long long* __restrict__ boundary_out) {
  __shared__ long long s_val[SPLIT_THREADS];
  __shared__ long long s_pos[SPLIT_THREADS];
  const int tid = threadIdx.x;
  const long long n_win = (right - left) - win + 1;          // window start positions 0 .. n_win - 1 (relative to left)
  const long long per = (n_win + SPLIT_THREADS - 1) / SPLIT_THREADS;
  const long long p0 = static_cast<long long>(tid) * per, p1 = min(p0 + per, n_win);
  long long best = LLONG_MAX, best_pos = LLONG_MAX;
  if (p0 < p1) {
    long long sum = 0;
    for (int k = 0; k < win; ++k) sum += split_quant(__ldg(x + left + p0 + k));
    best = sum;
    best_pos = p0;
    for (long long p = p0 + 1; p < p1; ++p) {
      sum += split_quant(__ldg(x + left + p - 1 + win)) - split_quant(__ldg(x + left + p - 1));
      if (sum < best) { best = sum; best_pos = p; }
    }
  }
  s_val[tid] = best;
  s_pos[tid] = best_pos;
  __syncthreads();
  for (int o = SPLIT_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o) {
      const long long v = s_val[tid + o], q = s_pos[tid + o];
      if (v < s_val[tid] || (v == s_val[tid] && q < s_pos[tid])) { s_val[tid] = v; s_pos[tid] = q; }
    }
    __syncthreads();
  }
  const long long wstart = s_pos[0];
  __syncthreads();
  // quietest sample inside the winning window (first on ties)
  best = LLONG_MAX;
  best_pos = LLONG_MAX;
  for (int k = tid; k < win; k += SPLIT_THREADS) {
    const long long v = split_quant(__ldg(x + left + wstart + k));
    if (v < best) { best = v; best_pos = k; }
  }
  s_val[tid] = best;
  s_pos[tid] = best_pos;
  __syncthreads();
  for (int o = SPLIT_THREADS / 2; o > 0; o >>= 1) {
    if (tid < o) {
      const long long v = s_val[tid + o], q = s_pos[tid + o];
      if (v < s_val[tid] || (v == s_val[tid] && q < s_pos[tid])) { s_val[tid] = v; s_pos[tid] = q; }
    }
    __syncthreads();
  }
  if (tid == 0) *boundary_out = left + wstart + s_pos[0];
}

double bessel_i0(double x) {
  const double q = (x / 2.0) * (x / 2.0);
  double term = 1.0, s = 1.0;
  for (int k = 1; k < 64; ++k) {
    term = term * q / (static_cast<double>(k) * k);
    s += term;
  }
  return s;
}

}  // namespace

// scipy.signal.resample_poly's default design: firwin(2 * half_len + 1, 1 / max(up, down), window = ("kaiser", 5.0)) * up
void design_resample_taps(int up, int down, std::vector<double>* taps, int* half_len) {
  const double PI = 3.14159265358979323846;
  const int max_rate = up > down ? up : down;
  const int hl = 10 * max_rate;
  const int n = 2 * hl + 1;
  const double cutoff = 1.0 / max_rate, alpha = 0.5 * (n - 1), beta = 5.0;
  taps->resize(n);
  double sum = 0.0;
  for (int k = 0; k < n; ++k) {
    const double m = k - alpha;
    const double px = PI * cutoff * m;
    const double sinc = m == 0.0 ? 1.0 : std::sin(px) / px;
    const double r = (k - alpha) / alpha;
    const double w = bessel_i0(beta * std::sqrt(std::fmax(0.0, 1.0 - r * r))) / bessel_i0(beta);
    (*taps)[k] = cutoff * sinc * w;
    sum += (*taps)[k];
  }
  for (int k = 0; k < n; ++k) (*taps)[k] = (*taps)[k] / sum * up;
  *half_len = hl;
}

// torchaudio.functional._get_sinc_resample_kernel(orig, new, gcd, lowpass_filter_width = 6, rolloff = 0.99, "sinc_interp_hann",
// dtype = float32) restated operation by operation IN FLOAT32, as torchaudio evaluates it for a float32 waveform (functional.py:
// 1340-1400): kernel[p][k] for p < new, k < 2 * width + orig.  Differences to torch are confined to its vectorised sinf / cosf
// (<= 1 ulp each).  orig / new are already divided by their gcd.
void design_sinc_hann_kernel(int orig, int newf, std::vector<float>* kernel, int* width_out) {
  const int lpw = 6;
  const double rolloff = 0.99;
  double base_freq = static_cast<double>(orig < newf ? orig : newf);
  base_freq *= rolloff;
  const int width = static_cast<int>(std::ceil(lpw * orig / base_freq));
  const int n_taps = 2 * width + orig;
  const float base_f = static_cast<float>(base_freq);
  const float scale_f = static_cast<float>(base_freq / orig);
  const float pi_f = static_cast<float>(3.14159265358979323846);
  kernel->assign(static_cast<size_t>(newf) * n_taps, 0.f);
  for (int p = 0; p < newf; ++p) {
    const float tp = static_cast<float>(-p) / static_cast<float>(newf);
    for (int k = 0; k < n_taps; ++k) {
      const float idx = static_cast<float>(k - width) / static_cast<float>(orig);
      float t = tp + idx;
      t *= base_f;
      t = std::fmin(std::fmax(t, -static_cast<float>(lpw)), static_cast<float>(lpw));
      const float c = cosf(t * pi_f / static_cast<float>(lpw) / 2.0f);
      const float window = c * c;
      t *= pi_f;
      float v = t == 0.f ? 1.0f : sinf(t) / t;
      v *= window * scale_f;
      (*kernel)[static_cast<size_t>(p) * n_taps + k] = v;
    }
  }
  *width_out = width;
}

// Samples after which the cascade has forgotten its initial state to 1e-18: from the largest pole radius of the sections.
int sos_warmup(const double* sos, int n_sections) {
  double rmax = 0.0;
  for (int s = 0; s < n_sections; ++s) {
    const double a1 = sos[6 * s + 4] / sos[6 * s + 3], a2 = sos[6 * s + 5] / sos[6 * s + 3];
    const double disc = a1 * a1 - 4.0 * a2;
    double r;
    if (disc < 0.0) r = std::sqrt(a2);
    else r = std::fmax(std::fabs((-a1 + std::sqrt(disc)) / 2.0), std::fabs((-a1 - std::sqrt(disc)) / 2.0));
    rmax = std::fmax(rmax, r);
  }
  if (!(rmax < 0.99999)) return -1;  // (marginally) unstable filter: the blocked evaluation does not apply
  if (rmax < 1e-3) return 32;
  const double w = std::log(1e-18) / std::log(rmax);
  return static_cast<int>(std::ceil(w / 32.0)) * 32 + 32 * n_sections;  // + a margin for the transient gain of the cascade
}

cudaError_t launch_ws_window(const int16_t* pcm16, const WsStream* streams_dev, int n_streams, long long max_out_len, const double* sos,
                             int n_sections, int warm, float* out, cudaStream_t stream) {
  if (n_streams == 0 || max_out_len == 0) return cudaSuccess;
  if (n_sections < 0 || n_sections > MAX_SECTIONS || n_streams > 65535) return cudaErrorInvalidValue;
  SosCoef c{};
  for (int s = 0; s < n_sections; ++s) {
    const double a0 = sos[6 * s + 3];
    c.b0[s] = sos[6 * s + 0] / a0; c.b1[s] = sos[6 * s + 1] / a0; c.b2[s] = sos[6 * s + 2] / a0;
    c.a1[s] = sos[6 * s + 4] / a0; c.a2[s] = sos[6 * s + 5] / a0;
  }
  const long long blocks = (max_out_len + SOS_L - 1) / SOS_L;
  dim3 grid(static_cast<unsigned int>((blocks + SOS_THREADS - 1) / SOS_THREADS), n_streams);
  switch (n_sections) {
    case 0: ws_window_kernel<0><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 1: ws_window_kernel<1><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 2: ws_window_kernel<2><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 3: ws_window_kernel<3><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 4: ws_window_kernel<4><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 5: ws_window_kernel<5><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 6: ws_window_kernel<6><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    case 7: ws_window_kernel<7><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
    default: ws_window_kernel<8><<<grid, SOS_THREADS, 0, stream>>>(pcm16, streams_dev, c, warm, out); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_resample_pcm16(const int16_t* in, const RsStream* streams_dev, int n_streams, long long max_out_len, const double* taps_dev,
                                  int n_taps, int half_len, int up, int down, int16_t* out, int num_sms, cudaStream_t stream) {
  if (n_streams == 0 || max_out_len == 0) return cudaSuccess;
  if (n_streams > 65535) return cudaErrorInvalidValue;
  const long long want = (max_out_len + 255) / 256;
  dim3 grid(static_cast<unsigned int>(want < 4LL * num_sms ? want : 4LL * num_sms), n_streams);
  resample_pcm16_kernel<<<grid, 256, 0, stream>>>(in, streams_dev, taps_dev, n_taps, half_len, up, down, out);
  return cudaGetLastError();
}

cudaError_t launch_split_scan(const float* x, long long left, long long right, int win, long long* boundary_out_dev, cudaStream_t stream) {
  if (right - left < win || win < 1) return cudaErrorInvalidValue;
  split_scan_kernel<<<1, SPLIT_THREADS, 0, stream>>>(x, left, right, win, boundary_out_dev);
  return cudaGetLastError();
}

// taps_dev: compact [newf][n_keep] float32, lo_dev: [newf] first kept tap of every phase
cudaError_t launch_resample_f32(const float* in, const RsStream* streams_dev, int n_streams, long long max_out_len, int channels,
                                const float* taps_dev, const int* lo_dev, int n_keep, int width, int orig, int newf, float* out,
                                int num_sms, cudaStream_t stream) {
  if (n_streams == 0 || max_out_len == 0) return cudaSuccess;
  if (n_streams > 65535 || channels < 1) return cudaErrorInvalidValue;
  const long long want = (max_out_len + 255) / 256;
  dim3 grid(static_cast<unsigned int>(want < 8LL * num_sms ? want : 8LL * num_sms), n_streams);
  if (orig == newf) downmix_f32_kernel<<<grid, 256, 0, stream>>>(in, streams_dev, channels, out);
  else resample_f32_kernel<<<grid, 256, 0, stream>>>(in, streams_dev, channels, taps_dev, lo_dev, n_keep, width, orig, newf, out);
  return cudaGetLastError();
}

}  // namespace qasr

// WebSocket pre-frontend on the device: what the reference does to a WS audio
// window on the CPU before the log-mel.
//   resample_pcm16_kernel  <- src/server.py:32-42  _resample_pcm_bytes: int16 -> polyphase FIR resample -> astype(int16)
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

}  // namespace qasr

// Log-mel kernels.  See mel.cuh for the math; this file holds the CTA-level choreography.
#include <cmath>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace qasr {

// Host: constants in double, rounded once.  Filter bank follows transformers/audio_utils.py:263-332,
// 356-375, 453-544 (Slaney scale, Slaney norm, 0..8000 Hz, 201 bins, 128 filters), cast f64 -> f32 as
// at feature_extraction_whisper.py:152.
void build_mel_tables(mel::Tables* t) {
  using namespace mel;
  std::memset(t, 0, sizeof(Tables));
  const double PI = 3.14159265358979323846;
  for (int k = 0; k < N_FFT; ++k) {
    t->w400[k] = make_float2(static_cast<float>(std::cos(2.0 * PI * k / N_FFT)), static_cast<float>(-std::sin(2.0 * PI * k / N_FFT)));
    t->window[k] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * PI * k / N_FFT));
  }
  auto hz_to_mel = [](double f) { return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4)) : 3.0 * f / 200.0; };
  auto mel_to_hz = [](double m) { return m >= 15.0 ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; };
  const double mel_min = hz_to_mel(0.0), mel_max = hz_to_mel(8000.0);
  std::vector<double> ff(N_MELS + 2);
  const double mel_step = (mel_max - mel_min) / (N_MELS + 1);  // numpy.linspace: arange * step + start, last = stop
  for (int i = 0; i < N_MELS + 2; ++i) ff[i] = mel_to_hz(i == N_MELS + 1 ? mel_max : i * mel_step + mel_min);
  int nnz = 0;
  for (int m = 0; m < N_MELS; ++m) {
    t->fptr[m] = nnz;
    t->flo[m] = 0;
    bool started = false;
    const double enorm = 2.0 / (ff[m + 2] - ff[m]);
    for (int k = 0; k < N_BINS; ++k) {
      const double f = 8000.0 * k / (N_BINS - 1);
      const double down = (f - ff[m]) / (ff[m + 1] - ff[m]);
      const double up = (ff[m + 2] - f) / (ff[m + 2] - ff[m + 1]);
      const double v = std::fmax(0.0, std::fmin(down, up)) * enorm;
      if (v > 0.0) {
        if (!started) { t->flo[m] = k; started = true; }
        // bins of one triangle are contiguous
        if (nnz < MAX_NNZ) t->fw[nnz] = static_cast<float>(v);
        ++nnz;
      }
    }
  }
  t->fptr[N_MELS] = nnz;
}

namespace {

using namespace mel;

struct MelSmem {
  float slab[SLAB];
  float2 X[FB * NC];
  float2 Y[FB * NC];   // reused as the power buffer [FB][P_PITCH] (FB*P_PITCH floats <= 2*FB*NC)
  float2 w400[N_FFT];
  float window[N_FFT];
  int fptr[N_MELS + 1];
  int flo[N_MELS];
  float fw[MAX_NNZ];
  float red[THREADS / 32];
};
static_assert(FB * P_PITCH <= 2 * FB * NC, "power buffer must fit in Y");

template <int R>
__device__ __forceinline__ void fft_stage(const float2* X, float2* Y, const float2* w400, int s, int m, int tw_step, int nf) {
  const int per_frame = m * s;  // butterflies per frame
  for (int item = threadIdx.x; item < nf * per_frame; item += THREADS) {
    const int f = item / per_frame, it = item - f * per_frame;
    butterfly<R>(X + f * NC, Y + f * NC, w400, it, s, m, tw_step);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS) logmel_kernel(const float* __restrict__ pcm, const MelSlab* __restrict__ slabs,
                                                         const Tables* __restrict__ tables, float* __restrict__ out,
                                                         long long ld, unsigned int* __restrict__ clip_max) {
  extern __shared__ uint8_t smem_raw[];
  MelSmem& sm = *reinterpret_cast<MelSmem*>(smem_raw);
  const MelSlab sl = slabs[blockIdx.x];
  const int tid = threadIdx.x;

  // constants -> smem
  for (int i = tid; i < N_FFT; i += THREADS) { sm.w400[i] = tables->w400[i]; sm.window[i] = tables->window[i]; }
  for (int i = tid; i < N_MELS + 1; i += THREADS) sm.fptr[i] = tables->fptr[i];
  for (int i = tid; i < N_MELS; i += THREADS) sm.flo[i] = tables->flo[i];
  for (int i = tid; i < MAX_NNZ; i += THREADS) sm.fw[i] = tables->fw[i];

  // PCM slab with centred reflect padding resolved by index mirroring
  const float* clip = pcm + sl.pcm_off;
  const int s0 = sl.frame0 * HOP - N_FFT / 2;
  const int need = (sl.n_frames - 1) * HOP + N_FFT;
  for (int u = tid; u < SLAB; u += THREADS) {
    float v = 0.f;
    if (u < need) v = __ldg(clip + reflect_index(s0 + u, sl.n_samples));
    sm.slab[u] = v;
  }
  __syncthreads();

  // window + pack: z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1]
  const int nf = sl.n_frames;
  for (int item = tid; item < nf * NC; item += THREADS) {
    const int f = item / NC, n = item - f * NC;
    const float* x = sm.slab + f * HOP + 2 * n;
    sm.X[item] = make_float2(x[0] * sm.window[2 * n], x[1] * sm.window[2 * n + 1]);
  }
  __syncthreads();

  // 200-point complex FFT, Stockham autosort, radices 5,5,4,2 (n = 200, 40, 8, 2)
  fft_stage<5>(sm.X, sm.Y, sm.w400, 1, 40, 2, nf);
  fft_stage<5>(sm.Y, sm.X, sm.w400, 5, 8, 10, nf);
  fft_stage<4>(sm.X, sm.Y, sm.w400, 25, 2, 50, nf);
  fft_stage<2>(sm.Y, sm.X, sm.w400, 100, 1, 200, nf);

  // power spectrum of the real DFT -> Y (as floats)
  float* P = reinterpret_cast<float*>(sm.Y);
  for (int item = tid; item < nf * N_BINS; item += THREADS) {
    const int f = item / N_BINS, k = item - f * N_BINS;
    P[f * P_PITCH + k] = power_bin(sm.X + f * NC, sm.w400, k);
  }
  __syncthreads();

  // sparse mel + log10; item = (filter m, frame f) with f fastest so a filter row writes FB contiguous floats
  float vmax = -INFINITY;
  float* orow = out + sl.col0 + sl.frame0;
  for (int item = tid; item < N_MELS * FB; item += THREADS) {
    const int m = item / FB, f = item - m * FB;
    if (f < nf) {
      const int b = sm.fptr[m], e = sm.fptr[m + 1];
      const float* p = P + f * P_PITCH + sm.flo[m];
      float acc = 0.f;
      for (int j = b; j < e; ++j) acc = fmaf(sm.fw[j], p[j - b], acc);
      const float v = log10f(fmaxf(acc, 1e-10f));
      vmax = fmaxf(vmax, v);
      orow[m * ld + f] = v;
    }
  }
  vmax = warp_max(vmax);
  if ((tid & 31) == 0) sm.red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    float v = sm.red[0];
#pragma unroll
    for (int i = 1; i < THREADS / 32; ++i) v = fmaxf(v, sm.red[i]);
    atomicMax(clip_max + sl.clip, float_to_ordered(v));
  }
}

// clamp to (clip max - 8) and rescale, in place.  grid = (column tiles, clips)
__global__ void logmel_finish_kernel(float* __restrict__ out, long long ld, const long long* __restrict__ clip_cols,
                                     const unsigned int* __restrict__ clip_max) {
  const int clip = blockIdx.y;
  const long long c0 = clip_cols[clip], c1 = clip_cols[clip + 1];
  const float floor_v = ordered_to_float(clip_max[clip]) - 8.0f;
  const long long t = c1 - c0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < t * N_MELS;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / t, c = i - m * t;
    float* p = out + m * ld + c0 + c;
    *p = (fmaxf(*p, floor_v) + 4.0f) * 0.25f;
  }
}

}  // namespace

cudaError_t launch_logmel(const float* pcm, const MelSlab* slabs, int n_slabs, const mel::Tables* tables, float* mel_out,
                          long long mel_ld, unsigned int* clip_max, cudaStream_t stream) {
  if (n_slabs == 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(MelSmem)));
  if (e != cudaSuccess) return e;
  logmel_kernel<<<n_slabs, mel::THREADS, sizeof(MelSmem), stream>>>(pcm, slabs, tables, mel_out, mel_ld, clip_max);
  return cudaGetLastError();
}

cudaError_t launch_logmel_finish(float* mel_out, long long mel_ld, const long long* clip_cols, int n_clips,
                                 const unsigned int* clip_max, cudaStream_t stream) {
  if (n_clips == 0) return cudaSuccess;
  dim3 grid(64, n_clips);
  logmel_finish_kernel<<<grid, 256, 0, stream>>>(mel_out, mel_ld, clip_cols, clip_max);
  return cudaGetLastError();
}

}  // namespace qasr

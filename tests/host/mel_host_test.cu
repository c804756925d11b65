// Host execution of the log-mel kernel's per-item math (mel.cuh is __host__ __device__): the same
// butterflies, real-input split, reflect indexing and filter tables the CUDA kernel uses, run on the
// CPU over one clip read from stdin-specified files.  Used by tests/test_host_math.py (no GPU).
//   mel_host_test <pcm_f32.bin> <out_f32.bin>   -> log10(max(mel,1e-10)) [128, T] (before clamp/scale)
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../qwen3_asr_b200/csrc/kernels.h"

using namespace qasr;
using namespace qasr::mel;

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<float> pcm;
  float buf[4096];
  size_t n;
  while ((n = fread(buf, 4, 4096, f)) > 0) pcm.insert(pcm.end(), buf, buf + n);
  fclose(f);
  const int N = static_cast<int>(pcm.size());
  const int T = N / HOP;
  Tables* t = new Tables();
  build_mel_tables(t);
  std::vector<float> out(static_cast<size_t>(N_MELS) * T);
  std::vector<float2> X(NC), Y(NC);
  std::vector<float> P(N_BINS);
  for (int fr = 0; fr < T; ++fr) {
    const int s0 = fr * HOP - N_FFT / 2;
    for (int i = 0; i < NC; ++i) {
      const float a = pcm[reflect_index(s0 + 2 * i, N)], b = pcm[reflect_index(s0 + 2 * i + 1, N)];
      X[i] = make_float2(a * t->window[2 * i], b * t->window[2 * i + 1]);
    }
    for (int it = 0; it < 40; ++it) butterfly<5>(X.data(), Y.data(), t->w400, it, 1, 40, 2);
    for (int it = 0; it < 40; ++it) butterfly<5>(Y.data(), X.data(), t->w400, it, 5, 8, 10);
    for (int it = 0; it < 50; ++it) butterfly<4>(X.data(), Y.data(), t->w400, it, 25, 2, 50);
    for (int it = 0; it < 100; ++it) butterfly<2>(Y.data(), X.data(), t->w400, it, 100, 1, 200);
    for (int k = 0; k < N_BINS; ++k) P[k] = power_bin(X.data(), t->w400, k);
    for (int m = 0; m < N_MELS; ++m) {
      float acc = 0.f;
      for (int j = t->fptr[m]; j < t->fptr[m + 1]; ++j) acc = fmaf(t->fw[j], P[t->flo[m] + j - t->fptr[m]], acc);
      out[static_cast<size_t>(m) * T + fr] = log10f(fmaxf(acc, 1e-10f));
    }
  }
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, out.size(), f);
  // filter table dump: nnz triples appended after the mel
  fclose(f);
  printf("%d %d %d\n", N, T, t->fptr[N_MELS]);
  return 0;
}

// Host execution of the log-mel kernel's per-unit math (mel.cuh is __host__ __device__): the same real 25-point DFTs,
// twiddles, 16-point FFTs, bin mapping, reflect indexing and filter tables the CUDA kernel uses, run on the CPU over
// one clip.  Used by tests/test_host_cpu.py (no GPU).
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
  if (!build_mel_tables(t)) return 3;  // structure mismatch with mel_structure.inc
  std::vector<float> out(static_cast<size_t>(N_MELS) * T);
  std::vector<float> P(N_BINS);
  for (int fr = 0; fr < T; ++fr) {
    const int s0 = fr * HOP - N_FFT / 2;
    float2 Y[16][K1];
    for (int j = 0; j < 16; ++j) {
      float v[25];
      for (int m = 0; m < 25; ++m) v[m] = pcm[reflect_index(s0 + 16 * m + j, N)] * t->window[16 * m + j];
      rdft25(v, Y[j]);
    }
    for (int k1 = 0; k1 < K1; ++k1) {
      float2 z[16];
      z[0] = Y[0][k1];
      for (int j = 1; j < 16; ++j) z[j] = cmul(Y[j][k1], t->tw[k1][j]);
      fft16(z);
      for (int k2 = 0; k2 < 16; ++k2)
        if (stage2_unique(k1, k2)) P[stage2_bin(k1, k2)] = fmaf(z[k2].x, z[k2].x, z[k2].y * z[k2].y);
    }
    for (int m = 0; m < N_MELS; ++m) {
      float acc = 0.f;
      for (int j = 0; j < kMelCnt[m]; ++j) acc = fmaf(t->fw[kMelPtr[m] + j], P[kMelLo[m] + j], acc);
      out[static_cast<size_t>(m) * T + fr] = log10f(fmaxf(acc, 1e-10f));
    }
  }
  f = fopen(argv[2], "wb");
  fwrite(out.data(), 4, out.size(), f);
  fclose(f);
  printf("%d %d %d\n", N, T, kMelNnz);
  return 0;
}

// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a, plus MUFU and LDS rates.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rate fp32_rate.cu && ./fp32_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{.reg .b64 ra, rb, rc, rd;\n\t mov.b64 ra, {%2,%3};\n\t mov.b64 rb, {%4,%5};\n\t mov.b64 rc, {%6,%7};\n\t fma.rn.f32x2 rd, ra, rb, rc;\n\t mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int CH = 8;  // independent chains per thread

__global__ void k_ffma(float* o, int n, float s) {
  float x[2 * CH];
  for (int i = 0; i < 2 * CH; ++i) x[i] = threadIdx.x * 0.001f + i;
  const float a = s, b = 1.0f - s;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) x[i] = fmaf(x[i], a, b);
  }
  float r = 0;
  for (int i = 0; i < 2 * CH; ++i) r += x[i];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_ffma2(float* o, int n, float s) {
  float2 x[CH];
  for (int i = 0; i < CH; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, i);
  const float2 a = make_float2(s, s), b = make_float2(1.0f - s, 1.0f - s);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fma2(x[i], a, b);
  }
  float r = 0;
  for (int i = 0; i < CH; ++i) r += x[i].x + x[i].y;
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_mufu(float* o, int n, float s) {
  float x[2 * CH];
  for (int i = 0; i < 2 * CH; ++i) x[i] = threadIdx.x * 0.001f + i * s;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float r = 0;
  for (int i = 0; i < 2 * CH; ++i) r += x[i];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_lds(float* o, int n) {
  __shared__ float4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = sm[(idx + i * 32) & 1023];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx += 7;
  }
  o[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* o; cudaMalloc(&o, sms * 8 * 1024 * sizeof(float));
  const int n = 20000, blocks = sms * 2, threads = 1024;
  const double lanes = double(blocks) * threads;
  float ms;
  ms = time_ms([&] { k_ffma<<<blocks, threads>>>(o, n, 0.5f); });
  printf("FFMA : %.3f ms  %.2f TFLOP/s (%.1f fma/clk/SM @1.9GHz)\n", ms, lanes * n * 2 * CH * 2 / ms / 1e9, lanes * n * 2 * CH / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(o, n, 0.5f); });
  printf("FFMA2: %.3f ms  %.2f TFLOP/s (%.1f fma/clk/SM @1.9GHz)\n", ms, lanes * n * CH * 4 / ms / 1e9, lanes * n * CH * 2 / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_mufu<<<blocks, threads>>>(o, n / 4, 0.5f); });
  printf("MUFU : %.3f ms  %.2f Tops/s (%.1f /clk/SM @1.9GHz)\n", ms, lanes * (n / 4) * 2 * CH / ms / 1e9, lanes * (n / 4) * 2 * CH / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_lds<<<blocks, threads>>>(o, n / 4); });
  printf("LDS.128: %.3f ms  %.1f B/clk/SM @1.9GHz\n", ms, lanes * (n / 4) * 8 * 16 / (ms * 1e-3) / sms / 1.9e9);
  return 0;
}

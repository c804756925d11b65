// Dynamic FP8 (e4m3) activation quantisation for the QUANTIZE=fp8 variant of the encoder's nn.Linear layers
// (reference src/server.py:362-371 -> torchao Float8DynamicActivationFloat8WeightConfig; torchao is not installable
// offline, so the arithmetic restated here is torchao's documented recipe and parity is pinned to oracle/encoder.py's
// emulation of it, not to torchao itself):
//     scale = max(amax(|x|), 1e-12) / 448        amax over the whole tensor (per-tensor) or over each row (per-row)
//     q     = e4m3_rn_satfinite(x / scale)
// The GEMM epilogue multiplies the fp32 accumulator by scale[m] * weight_scale[n].
#include <cuda_fp8.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "kernels.h"

namespace qasr {
namespace {

constexpr int Q_WARPS = 8;
constexpr float kE4M3Max = 448.0f;
constexpr float kAmaxEps = 1e-12f;

__device__ __forceinline__ float amax8(const uint4& u) {
  // |x| of 8 packed bf16: clear the sign bits, compare as floats
  float2 a = unpack_bf16x2(u.x & 0x7fff7fffu), b = unpack_bf16x2(u.y & 0x7fff7fffu);
  float2 c = unpack_bf16x2(u.z & 0x7fff7fffu), d = unpack_bf16x2(u.w & 0x7fff7fffu);
  return fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)), fmaxf(fmaxf(c.x, c.y), fmaxf(d.x, d.y)));
}
__device__ __forceinline__ uint32_t quant4(float2 lo, float2 hi, const FastDivisor& sc) {
  const unsigned short p0 = __nv_cvt_float2_to_fp8x2(make_float2(sc.div(lo.x), sc.div(lo.y)), __NV_SATFINITE, __NV_E4M3);
  const unsigned short p1 = __nv_cvt_float2_to_fp8x2(make_float2(sc.div(hi.x), sc.div(hi.y)), __NV_SATFINITE, __NV_E4M3);
  return static_cast<uint32_t>(p0) | (static_cast<uint32_t>(p1) << 16);
}

// per-tensor pass 1: amax over [rows, k] -> atomicMax on the float's bit pattern (non-negative floats order like uints)
__global__ void __launch_bounds__(256) absmax_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int rows, int k, unsigned int* __restrict__ amax_bits) {
  const int vec_per_row = k / 8;
  const long long total = static_cast<long long>(rows) * vec_per_row;
  float m = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row);
    m = fmaxf(m, amax8(*reinterpret_cast<const uint4*>(x + r * ld + c * 8)));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}

// one warp per row: (per-row) amax of the row, then convert; (per-tensor) scale from the finished amax of pass 1
template <bool PER_ROW>
__global__ void __launch_bounds__(Q_WARPS * 32) quant_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int rows, int k,
                                                                  uint8_t* __restrict__ q, long long ldq, float* __restrict__ row_scale,
                                                                  const unsigned int* __restrict__ amax_bits) {
  const int lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * Q_WARPS;
  const int vecs = k / 8;
  float tensor_scale = 0.f;
  if (!PER_ROW) tensor_scale = fmaxf(__uint_as_float(*amax_bits), kAmaxEps) / kE4M3Max;
  for (int row = blockIdx.x * Q_WARPS + (threadIdx.x >> 5); row < rows; row += n_warps) {
    const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<long long>(row) * ldx);
    float scale = tensor_scale;
    if (PER_ROW) {
      float m = 0.f;
      for (int v = lane; v < vecs; v += 32) m = fmaxf(m, amax8(src[v]));
      scale = fmaxf(warp_max(m), kAmaxEps) / kE4M3Max;
    }
    uint2* dst = reinterpret_cast<uint2*>(q + static_cast<long long>(row) * ldq);
    const FastDivisor sc(scale);
    for (int v = lane; v < vecs; v += 32) {
      const uint4 u = src[v];  // second touch of the row: L1 / L2 hit
      uint2 o;
      o.x = quant4(unpack_bf16x2(u.x), unpack_bf16x2(u.y), sc);
      o.y = quant4(unpack_bf16x2(u.z), unpack_bf16x2(u.w), sc);
      dst[v] = o;
    }
    if (lane == 0) row_scale[row] = scale;
  }
}

}  // namespace

cudaError_t launch_quant_fp8(const __nv_bfloat16* x, long long ldx, int rows, int k, uint8_t* q, long long ldq, float* row_scale,
                             unsigned int* amax_slot, bool per_row, int num_sms, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (k % 8 != 0 || ldx % 8 != 0 || ldq % 8 != 0) return cudaErrorInvalidValue;
  const int grid = std::max(1, std::min((rows + Q_WARPS - 1) / Q_WARPS, 8 * num_sms));
  if (per_row) {
    quant_rows_kernel<true><<<grid, Q_WARPS * 32, 0, stream>>>(x, ldx, rows, k, q, ldq, row_scale, nullptr);
  } else {
    // amax_slot was zeroed by the caller on this stream
    absmax_kernel<<<std::min(4 * num_sms, grid), 256, 0, stream>>>(x, ldx, rows, k, amax_slot);
    quant_rows_kernel<false><<<grid, Q_WARPS * 32, 0, stream>>>(x, ldx, rows, k, q, ldq, row_scale, amax_slot);
  }
  return cudaGetLastError();
}

// Host: weight quantisation at finalize.  w is [n, k] float (bf16-representable values); rows are split into `modules`
// equal blocks (3 for the fused q|k|v weight) that each own a per-tensor scale; per_row gives every output row its own.
void quantize_weight_e4m3(const float* w, int n, int k, int modules, bool per_row, uint8_t* q_out, float* scale_out) {
  const int rows_per_module = n / modules;
  for (int mdl = 0; mdl < modules; ++mdl) {
    float amax_t = 0.f;
    if (!per_row)
      for (long long i = static_cast<long long>(mdl) * rows_per_module * k; i < static_cast<long long>(mdl + 1) * rows_per_module * k; ++i)
        amax_t = std::max(amax_t, std::fabs(w[i]));
    for (int r = mdl * rows_per_module; r < (mdl + 1) * rows_per_module; ++r) {
      const float* row = w + static_cast<long long>(r) * k;
      float amax = amax_t;
      if (per_row) {
        amax = 0.f;
        for (int c = 0; c < k; ++c) amax = std::max(amax, std::fabs(row[c]));
      }
      const float scale = std::max(amax, kAmaxEps) / kE4M3Max;
      scale_out[r] = scale;
      for (int c = 0; c < k; ++c)
        q_out[static_cast<long long>(r) * k + c] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(row[c] / scale, __NV_SATFINITE, __NV_E4M3));
    }
  }
}

}  // namespace qasr

// Shared device/host helpers for the qasr_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "qasr_b200 targets sm_100a only"
#endif

namespace qasr {

// ---- host-side error plumbing -------------------------------------------------------------
void set_last_error(const std::string& msg);

#define QASR_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::qasr::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                             __FILE__ + ":" + std::to_string(__LINE__));                   \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

#define QASR_REQUIRE(cond, msg)                   \
  do {                                            \
    if (!(cond)) {                                \
      ::qasr::set_last_error(std::string(msg));   \
      return 1;                                   \
    }                                             \
  } while (0)

// ---- numerics ------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// erf GELU (torch F.gelu default, ACT2FN["gelu"]): 0.5 x (1 + erf(x / sqrt 2)) = x - x q(|x|) for x >= 0, x q(|x|)
// for x < 0 (together: relu(x) - |x| q), with q(a) = 0.5 erfc(a / sqrt 2) from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): branch-free,
// 2 MUFU + ~12 FP32 ops, no cancellation in the negative tail.  Over all bf16 inputs its bf16-rounded result differs
// from the float64 GELU in 169 of 35898 values (by 1 ulp), torch's own float32 erff path in 129 (tests/test_host_cpu.py).
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fabsf(x);
  const float t = rcp_approx(fmaf(0.231641888f, a, 1.0f));          // 1 / (1 + 0.3275911 a / sqrt 2)
  float p = fmaf(0.5307027145f, t, -0.7265760135f);                  // 0.5 * (a5 .. a1), Horner in t
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  const float e = ex2_approx(a * a * -0.72134752044f);               // exp(-a^2 / 2)
  const float q = p * t * e;
  return fmaf(-a, q, fmaxf(x, 0.f));                                 // relu(x) - |x| q: no predicate, no select
}

// erf GELU of a bf16 VALUE by table (round 2).  Every GELU of the encoder takes an input that has just been rounded to bf16 and
// rounds its output to bf16 again, so the function is a map between 16-bit patterns.  gelu(x) = x * r(x): the table holds r for the
// |x| in [2^-9, 16) of either sign (1664 patterns each) plus one sentinel below (r = 1/2: bf16(x / 2) is the exact result for every
// smaller |x|) and one above (r = 1 for x >= 16; r = 0 for x <= -16, so the product is -0 like torch's and NaN for -inf like torch's).
// r(b) is the float nearest to target(b) / x(b), target = the float64 erf GELU rounded to nearest-even bf16, so bf16(x * r) IS the
// correctly rounded GELU for every one of the 65 536 inputs (build_gelu_lut checks all of them) -- the 14-instruction formula above
// is off by one bf16 ulp on 169 of them, torch's own float32 path on 129.  One shared-memory load + one multiply instead of 2 MUFU
// + 12 FP32 operations.
namespace gelu_tab {
constexpr uint32_t A_LO = 0x3B00u;            // bf16 bits of 2^-9
constexpr uint32_t A_HI = 0x417Fu;            // largest bf16 below 16
constexpr int N = A_HI - A_LO + 3;            // per sign, with the two sentinels
constexpr int WORDS = 2 * N;                  // 3332 floats = 13.3 KB
}  // namespace gelu_tab
bool build_gelu_lut(float* lut /* [gelu_tab::WORDS] */);   // elementwise.cu (host); false if the exhaustive check fails

// ratio r for the bf16 value with bit pattern `bits16` (low 16 bits)
__device__ __forceinline__ float gelu_ratio(const float* __restrict__ lut_s, uint32_t bits16) {
  const uint32_t a = bits16 & 0x7fffu;
  const uint32_t i = min(max(a, gelu_tab::A_LO - 1u), gelu_tab::A_HI + 1u) - (gelu_tab::A_LO - 1u);
  return lut_s[(bits16 >> 15) * gelu_tab::N + i];
}
// two accumulators -> bf16 round -> GELU -> bf16, packed
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi);
__device__ __forceinline__ uint32_t gelu_lut_bf16x2(const float* __restrict__ lut_s, float c0, float c1) {
  const uint32_t p = pack_bf16x2(c0, c1);
  const float x0 = __uint_as_float(p << 16), x1 = __uint_as_float(p & 0xffff0000u);
  return pack_bf16x2(x0 * gelu_ratio(lut_s, p & 0xffffu), x1 * gelu_ratio(lut_s, p >> 16));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Round two floats to bf16 and back with ONE packed conversion (F2FP, ALU pipe) + two logic ops; the scalar
// cvt.rn.bf16.f32 (F2F) runs on the 16-lane XU pipe shared with MUFU and was a bottleneck of the GELU epilogues.
__device__ __forceinline__ float2 bf16_round2(float a, float b) {
  const uint32_t p = pack_bf16x2(a, b);
  return make_float2(__uint_as_float(p << 16), __uint_as_float(p & 0xffff0000u));
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// x / s for many x and one s: r = 1 / s once, then two FMAs give the correctly rounded quotient (q0 = x r; q = q0 + (x - q0 s) r),
// i.e. what IEEE division returns, without the slow path of the generic division sequence (no denormals / overflow here:
// s >= 1e-12 / 448 and |x / s| <= 448 by construction).
struct FastDivisor {
  float s, r;
  __device__ __forceinline__ explicit FastDivisor(float scale) : s(scale), r(1.0f / scale) {}
  __device__ __forceinline__ float div(float x) const {
    const float q0 = x * r;
    return fmaf(fmaf(-q0, s, x), r, q0);
  }
};

// order-preserving float <-> uint32 map for atomicMax on floats of either sign
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(b);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kNumSMs = 148;

// fixed-point scales of the integer-atomic LayerNorm statistics (epilogues.cuh RowStatsAtomic / LnFoldAcc)
constexpr float kStatSumScale = 1048576.0f;     // 2^20: |sum of a row| < 2^43, far beyond any activation
constexpr float kStatSqScale = 65536.0f;        // 2^16

// ---- programmatic dependent launch --------------------------------------------------------------
// The step is a chain of ~176 dependent kernels; between two of them the GPU otherwise idles for the launch latency and the
// next kernel's prologue (barrier init, TMEM allocation, cluster sync, descriptor prefetch).  Kernels on the chain call
// pdl_launch_dependents() first (the next grid may be scheduled onto SMs as they drain) and pdl_wait() before they touch global
// memory (it returns once the previous grid has completed and its writes are visible), and are launched with launch_pdl().
// Both device calls are no-ops for a kernel launched without the attribute.  QASR_PDL=0 disables the attribute (A/B).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // qasr.cu

template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attrs[2];
  int n = 0;
  if (cluster_x > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster_x;
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attrs;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace qasr

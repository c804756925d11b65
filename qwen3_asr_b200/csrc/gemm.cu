// Launchers for the tcgen05 GEMM family and TMA tensor-map construction.
#include "gemm.h"

#include <algorithm>
#include <string>

#include "epilogues.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm_small.cuh"

namespace qasr {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;

}  // namespace

int tmap_api_init() {
  if (g_encode_tiled != nullptr) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_last_error(std::string("cuTensorMapEncodeTiled not available from the driver: ") + cudaGetErrorString(e));
    return 2;
  }
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

int make_tmap_rowmajor(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  if (tmap_api_init() != 0) return 2;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(tc::BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(2d) failed with CUresult " + std::to_string(static_cast<int>(r)) + " rows=" +
                   std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld));
    return 2;
  }
  return 0;
}

int make_tmap_rowmajor_u8(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  if (tmap_api_init() != 0) return 2;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(tc::BLOCK_K_BYTES), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(2d, u8) failed with CUresult " + std::to_string(static_cast<int>(r)) + " rows=" +
                   std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld));
    return 2;
  }
  return 0;
}

int make_tmap_conv(CUtensorMap* tm, const void* base, long long g_in, int h_in, int c, int hc, int gt) {
  if (tmap_api_init() != 0) return 2;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(h_in), static_cast<cuuint64_t>(g_in)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(c) * 2 * static_cast<cuuint64_t>(h_in)};
  // with a traversal stride s the box extent is (elements to load) * s
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(tc::BLOCK_K), static_cast<cuuint32_t>(2 * hc), static_cast<cuuint32_t>(2 * gt)};
  const cuuint32_t estr[3] = {1, 2, 2};
  const CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled(conv) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 2;
  }
  return 0;
}

int pick_bn_small(int n, int k) { return (k % tc::BLOCK_K == 0 && n % 32 == 0) ? 32 : 0; }

int pick_bn(int n) {
  if (n % 256 == 0) return 256;
  if (n % 128 == 0) return 128;
  if (n % 64 == 0) return 64;
  return 0;
}

namespace {

template <int BN, int AMODE, int KIND, class Epi>
cudaError_t launch_tc(const CUtensorMap& tm_a, const CUtensorMap& tm_b, const tc::GemmShape& shape, const Epi& epi, int num_sms,
                      cudaStream_t stream) {
  constexpr bool CTA2 = kGemmCta2;
  // LayerNorm-folded Linears double-buffer two per-column tables: one ring stage less pays for them
  constexpr int ST = tc::default_stages<BN, CTA2>() - (Epi::kLnFold ? 1 : 0);
  using L = tc::SmemLayout<BN, ST, Epi::kScaled, CTA2, Epi::kLnFold>;
  auto kern = tc::gemm_tc_kernel<BN, ST, AMODE, KIND, CTA2, Epi>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
  if (e != cudaSuccess) return e;
  if (shape.m_tiles <= 0 || shape.n_tiles <= 0) return cudaSuccess;
  if constexpr (!CTA2) {
    const int grid = std::min(shape.m_tiles * shape.n_tiles, num_sms);
    kern<<<grid, tc::kThreads, L::TOTAL, stream>>>(tm_a, tm_b, shape, epi);
    return cudaGetLastError();
  } else {
    // persistent CTA pairs: as many 2-CTA clusters as the device can keep resident at once (one CTA per SM; a GPC with an
    // odd number of free SMs leaves one idle), never more than there are 256-row tiles
    static thread_local int max_clusters_cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int max_clusters = dev < 64 ? max_clusters_cached[dev] : 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t cfg{};
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.blockDim = dim3(tc::kThreads);
      cfg.dynamicSmemBytes = L::TOTAL;
      cfg.stream = stream;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      cfg.gridDim = dim3(2 * (num_sms / 2));
      e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
      if (e != cudaSuccess) return e;
      if (max_clusters <= 0) return cudaErrorLaunchOutOfResources;
      max_clusters = std::min(max_clusters, num_sms / 2);
      if (dev < 64) max_clusters_cached[dev] = max_clusters;
    }
    const int tiles = ((shape.m_tiles + 1) / 2) * shape.n_tiles;
    return launch_pdl(kern, dim3(2 * std::min(tiles, max_clusters)), dim3(tc::kThreads), L::TOTAL, stream, 2, tm_a, tm_b, shape, epi);
  }
}

// one CTA per BN-column slice of the weight, the whole slice resident (tc_gemm_small.cuh)
template <int BN, class Epi>
cudaError_t launch_small(const CUtensorMap& tm_a, const CUtensorMap& tm_b, int n_tiles, int num_kb, const Epi& epi, cudaStream_t stream) {
  using L = tc::SmallLayout<BN>;
  auto kern = tc::gemm_smallm_kernel<BN, Epi>;
  const int smem = L::total(num_kb);
  if (num_kb > tc::kSmallMaxKb || smem > 232448) return cudaErrorInvalidValue;
  // the attribute is per function AND device: set it on every launch like launch_tc does (a host-side table lookup)
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  return launch_pdl(kern, dim3(n_tiles), dim3(tc::kSmallThreads), smem, stream, 1, tm_a, tm_b, num_kb, epi);
}

template <class ALoad, class Epi>
cudaError_t launch_simt(const ALoad& aload, const __nv_bfloat16* b, long long ldb, int m, int n, int k, const Epi& epi, cudaStream_t stream) {
  const long long items = static_cast<long long>(m) * (n / 8);
  if (items <= 0) return cudaSuccess;
  const int threads = 128;
  const long long blocks = (items + threads - 1) / threads;
  tc::gemm_simt_kernel<<<static_cast<unsigned int>(blocks), threads, 0, stream>>>(aload, b, ldb, m, n, k, epi);
  return cudaGetLastError();
}

template <int KIND, class Epi>
cudaError_t linear_dispatch(const LinearArgs& a, const Epi& epi, bool simt, int num_sms, cudaStream_t stream) {
  if (simt) {
    if (KIND != tc::K_BF16) return cudaErrorNotSupported;  // the SIMT checker reads bf16 operands
    tc::ALoadLinear al{a.a, a.lda, a.m};
    return launch_simt(al, a.b, a.ldb, a.m, a.n, a.k, epi, stream);
  }
  int bn = a.bn;
  const CUtensorMap* tm_b = a.tm_b;
  if constexpr (KIND == tc::K_BF16 && !Epi::kScaled && !Epi::kLnPart) {
    // a single window or chunk (the unbatched reference's call shape) is latency-bound on 256-column tiles.  K <= 1024: the
    // small-M kernel (tc_gemm_small.cuh), one CTA per 32-column weight slice.  Longer K (fc2): the slice would not fit shared
    // memory, and 16-column slices would change the 32-column granularity at which the residual epilogues round their LayerNorm
    // partial sums (bit-exact batch invariance) -- the pair kernel on 64-column tiles instead: twice the pairs at work.
    if (a.m <= tc::BLOCK_M && a.tm_b_small != nullptr && a.bn_small == 32) {
      if (a.k <= 1024) return launch_small<32>(*a.tm_a, *a.tm_b_small, a.n / 32, a.k / tc::BLOCK_K, epi, stream);
      if (kGemmCta2 && a.n % 64 == 0) { bn = 64; tm_b = a.tm_b_small; }   // a pair stages 2 x 32 weight rows: the same box
    }
  }
  tc::GemmShape sh{};
  sh.m_tiles = (a.m + tc::BLOCK_M - 1) / tc::BLOCK_M;
  sh.n_tiles = a.n / bn;
  sh.num_kb = a.k / tc::block_k_elems<KIND>();
  sh.kb_per_tap = 1;
  sh.gt = 1;
  switch (bn) {
    case 256: return launch_tc<256, tc::A_LINEAR, KIND>(*a.tm_a, *tm_b, sh, epi, num_sms, stream);
    case 128: return launch_tc<128, tc::A_LINEAR, KIND>(*a.tm_a, *tm_b, sh, epi, num_sms, stream);
    case 64: return launch_tc<64, tc::A_LINEAR, KIND>(*a.tm_a, *tm_b, sh, epi, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <class Epi>
cudaError_t linear_kind(const LinearArgs& a, const Epi& e, bool simt, int num_sms, cudaStream_t stream) {
  if (a.fp8) {
    Scaled<Epi> se{e, a.row_scale, a.col_scale};
    return linear_dispatch<tc::K_E4M3>(a, se, simt, num_sms, stream);
  }
  return linear_dispatch<tc::K_BF16>(a, e, simt, num_sms, stream);
}

// LayerNorm-folded Linears exist in bf16 only and never run through the SIMT checker (handles in those modes keep the separate
// LayerNorm kernel)
template <class Epi>
cudaError_t linear_ln(const LinearArgs& a, const Epi& e, bool simt, int num_sms, cudaStream_t stream) {
  if (simt || a.fp8 || a.ln_colsum == nullptr) return cudaErrorNotSupported;
  if (a.ln_acc != nullptr) {
    LnFoldAcc<Epi> la{e, a.ln_acc, a.ln_colsum, 1.0f / static_cast<float>(a.k), 1e-5f};
    return linear_dispatch<tc::K_BF16>(a, la, false, num_sms, stream);
  }
  if (a.ln_part != nullptr) {
    if (a.k % 64 != 0) return cudaErrorInvalidValue;  // an even number of 32-column panels: 16-byte aligned rows of partials
    LnFoldPart<Epi> lp{{e, nullptr, a.ln_colsum}, a.ln_part, a.k / 32, 1.0f / static_cast<float>(a.k), 1e-5f};
    return linear_dispatch<tc::K_BF16>(a, lp, false, num_sms, stream);
  }
  LnFold<Epi> le{e, a.ln_stats, a.ln_colsum};
  return linear_dispatch<tc::K_BF16>(a, le, false, num_sms, stream);
}

}  // namespace

cudaError_t gemm_linear(const LinearArgs& a, bool simt, int num_sms, cudaStream_t stream) {
  if (a.m <= 0) return cudaSuccess;
  const int kb = a.fp8 ? tc::block_k_elems<tc::K_E4M3>() : tc::BLOCK_K;
  if (a.k % kb != 0 || a.n % 16 != 0 || (!simt && (a.bn == 0 || a.n % a.bn != 0))) return cudaErrorInvalidValue;
  if (a.fp8 && (a.row_scale == nullptr || a.col_scale == nullptr)) return cudaErrorInvalidValue;
  if (a.row_map != nullptr && a.epi != LIN_PLAIN) return cudaErrorInvalidValue;
  const bool ln_any = a.ln_stats != nullptr || a.ln_part != nullptr || a.ln_acc != nullptr;
  if (ln_any && a.epi != LIN_GELU && a.epi != LIN_QKV) return cudaErrorInvalidValue;
  switch (a.epi) {
    case LIN_PLAIN:
      if (a.row_map != nullptr) return linear_kind(a, EpiLinearRows{{a.out, a.bias, nullptr, a.ldo, a.m, a.n}, a.row_map}, simt, num_sms, stream);
      return linear_kind(a, EpiLinear<ACT_NONE, false>{a.out, a.bias, nullptr, a.ldo, a.m, a.n}, simt, num_sms, stream);
    case LIN_GELU:
      if (ln_any) return linear_ln(a, EpiLinear<ACT_GELU, false>{a.out, a.bias, nullptr, a.ldo, a.m, a.n}, simt, num_sms, stream);
      return linear_kind(a, EpiLinear<ACT_GELU, false>{a.out, a.bias, nullptr, a.ldo, a.m, a.n}, simt, num_sms, stream);
    case LIN_RESIDUAL:
      if (a.stats_acc != nullptr) {
        if (simt || a.fp8) return cudaErrorNotSupported;
        RowStatsAtomic<EpiLinear<ACT_NONE, true>> e{{a.out, a.bias, a.residual, a.ldo, a.m, a.n}, a.stats_acc};
        return linear_dispatch<tc::K_BF16>(a, e, false, num_sms, stream);
      }
      if (a.stats_part != nullptr) {
        if (simt || a.fp8 || a.n % 64 != 0) return cudaErrorNotSupported;
        RowStats<EpiLinear<ACT_NONE, true>> e{{a.out, a.bias, a.residual, a.ldo, a.m, a.n}, a.stats_part, a.n / 32};
        return linear_dispatch<tc::K_BF16>(a, e, false, num_sms, stream);
      }
      return linear_kind(a, EpiLinear<ACT_NONE, true>{a.out, a.bias, a.residual, a.ldo, a.m, a.n}, simt, num_sms, stream);
    case LIN_QKV: {
      if (a.n % 64 != 0 || a.head_rows < a.m) return cudaErrorInvalidValue;
      EpiQkv e{{a.out, a.bias, nullptr, 0, a.m, a.n}, a.head_rows};
      if (ln_any) return linear_ln(a, e, simt, num_sms, stream);
      return linear_kind(a, e, simt, num_sms, stream);
    }
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t gemm_conv(const ConvArgs& a, bool simt, int num_sms, cudaStream_t stream) {
  if (a.n_chunks <= 0) return cudaSuccess;
  if (a.c != 480 || a.hc * a.gt != tc::BLOCK_M) return cudaErrorInvalidValue;
  EpiConv e{a.out, a.bias, a.width, a.hc, a.slots, a.max_w, a.out_pitch, a.out_off, a.n_chunks, a.c};
  const long long g_out = static_cast<long long>(a.n_chunks) * a.slots;  // global output columns
  if (simt) {
    tc::ALoadConv al{a.a, static_cast<int>(a.g_in), a.h_in, a.c, a.hc};
    return launch_simt(al, a.b, 9 * 512, static_cast<int>(g_out * a.hc), a.c, 9 * 512, e, stream);
  }
  tc::GemmShape sh{};
  sh.m_tiles = static_cast<int>((g_out + a.gt - 1) / a.gt);
  sh.n_tiles = 2;
  sh.num_kb = 9 * 8;
  sh.kb_per_tap = 8;
  sh.gt = a.gt;
  return launch_tc<240, tc::A_CONV, tc::K_BF16>(*a.tm_a, *a.tm_b, sh, e, num_sms, stream);
}

namespace {
template <int KIND, class Epi>
cudaError_t conv_out_dispatch(const ConvOutArgs& a, const Epi& e, int num_sms, cudaStream_t stream) {
  tc::GemmShape sh{};
  sh.m_tiles = (a.m + tc::BLOCK_M - 1) / tc::BLOCK_M;
  sh.n_tiles = a.d / a.bn;
  sh.num_kb = a.k / tc::block_k_elems<KIND>();
  sh.kb_per_tap = 1;
  sh.gt = 1;
  switch (a.bn) {
    case 256: return launch_tc<256, tc::A_LINEAR, KIND>(*a.tm_a, *a.tm_b, sh, e, num_sms, stream);
    case 128: return launch_tc<128, tc::A_LINEAR, KIND>(*a.tm_a, *a.tm_b, sh, e, num_sms, stream);
    case 64: return launch_tc<64, tc::A_LINEAR, KIND>(*a.tm_a, *a.tm_b, sh, e, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

cudaError_t gemm_conv_out(const ConvOutArgs& a, bool simt, int num_sms, cudaStream_t stream) {
  if (a.m <= 0) return cudaSuccess;
  const int kb = a.fp8 ? tc::block_k_elems<tc::K_E4M3>() : tc::BLOCK_K;
  if (a.k % kb != 0 || a.d % 16 != 0) return cudaErrorInvalidValue;
  EpiConvOut e{a.out, a.pe, a.row_token, a.tok_per_chunk, a.d, a.m};
  if (simt) {
    if (a.fp8) return cudaErrorNotSupported;
    tc::ALoadLinear al{a.a, a.k, a.m};
    return launch_simt(al, a.b, a.k, a.m, a.d, a.k, e, stream);
  }
  if (a.fp8) {
    if (a.row_scale == nullptr || a.col_scale == nullptr) return cudaErrorInvalidValue;
    Scaled<EpiConvOut> se{e, a.row_scale, a.col_scale};
    return conv_out_dispatch<tc::K_E4M3>(a, se, num_sms, stream);
  }
  if (a.stats_acc != nullptr) {
    RowStatsAtomic<EpiConvOut> re{e, a.stats_acc};
    return conv_out_dispatch<tc::K_BF16>(a, re, num_sms, stream);
  }
  if (a.stats_part != nullptr) {
    RowStats<EpiConvOut> re{e, a.stats_part, a.d / 32};
    return conv_out_dispatch<tc::K_BF16>(a, re, num_sms, stream);
  }
  return conv_out_dispatch<tc::K_BF16>(a, e, num_sms, stream);
}

}  // namespace qasr

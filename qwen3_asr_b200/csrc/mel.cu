// Log-mel kernel: one persistent kernel, work items claimed through a global ticket (see mel.cuh for the math).
//
// Per item of FB = 32 frames (256 threads, 2 CTAs per SM):
//   load     PCM slab -> smem rows of one hop (160 samples, pitch 176), reflect padding by index mirroring
//   stage 1  thread = (frame, j): 25 strided samples x window -> real 25-point DFT in registers -> 13 complex to the
//            exchange buffer E (XOR-swizzled 16-byte chunks: conflict-free for the 8-byte writes and the 16-byte reads)
//   stage 2  thread = (frame, k1): 16 complex from E -> twiddle -> 16-point FFT in registers -> |X|^2 -> P[frame][bin]
//   mel      warp = a group of filters (fully unrolled from the compiled-in structure, weights from constant memory),
//            lane = frame: log10, raw value to HBM (128 B per warp and filter), per-clip max via atomicMax
// The max-8 clamp needs the whole clip's max, so it runs as later work items of the same kernel (kind 1) that wait on
// the clip's done counter and rewrite tiles that are still L2-resident; tickets are handed out in list order, every
// item a clamp waits for has a smaller ticket and is therefore already running or finished -> no deadlock, no second
// kernel, and the log-mel makes one trip to DRAM.
#include <cmath>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace qasr {

namespace {
__constant__ float c_mel_fw[mel::MAX_NNZ];
}

// Host: constants in double, rounded once.  Filter bank follows transformers/audio_utils.py:263-332,
// 356-375, 453-544 (Slaney scale, Slaney norm, 0..8000 Hz, 201 bins, 128 filters), cast f64 -> f32 as
// at feature_extraction_whisper.py:152.  Returns false if the computed sparsity structure differs from
// the compiled-in one (mel_structure.inc).
bool build_mel_tables(mel::Tables* t) {
  using namespace mel;
  std::memset(t, 0, sizeof(Tables));
  const double PI = 3.14159265358979323846;
  for (int k = 0; k < N_FFT; ++k) t->window[k] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * PI * k / N_FFT));
  for (int k1 = 0; k1 < K1; ++k1)
    for (int j = 0; j < 16; ++j) {
      const double a = 2.0 * PI * ((k1 * j) % N_FFT) / N_FFT;
      t->tw[k1][j] = make_float2(static_cast<float>(std::cos(a)), static_cast<float>(-std::sin(a)));
    }
  auto hz_to_mel = [](double f) { return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4)) : 3.0 * f / 200.0; };
  auto mel_to_hz = [](double m) { return m >= 15.0 ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; };
  const double mel_min = hz_to_mel(0.0), mel_max = hz_to_mel(8000.0);
  std::vector<double> ff(N_MELS + 2);
  const double mel_step = (mel_max - mel_min) / (N_MELS + 1);  // numpy.linspace: arange * step + start, last = stop
  for (int i = 0; i < N_MELS + 2; ++i) ff[i] = mel_to_hz(i == N_MELS + 1 ? mel_max : i * mel_step + mel_min);
  int nnz = 0;
  bool ok = true;
  for (int m = 0; m < N_MELS; ++m) {
    int lo = -1, cnt = 0;
    const double enorm = 2.0 / (ff[m + 2] - ff[m]);
    for (int k = 0; k < N_BINS; ++k) {
      const double f = 8000.0 * k / (N_BINS - 1);
      const double down = (f - ff[m]) / (ff[m + 1] - ff[m]);
      const double up = (ff[m + 2] - f) / (ff[m + 2] - ff[m + 1]);
      const double v = std::fmax(0.0, std::fmin(down, up)) * enorm;
      if (v > 0.0) {
        if (lo < 0) lo = k;
        if (k != lo + cnt) ok = false;  // bins of one triangle are contiguous
        if (nnz < MAX_NNZ) t->fw[nnz] = static_cast<float>(v);
        ++nnz;
        ++cnt;
      }
    }
    if (lo != kMelLo[m] || cnt != kMelCnt[m]) ok = false;
  }
  return ok && nnz == kMelNnz;
}

cudaError_t upload_mel_constants(const mel::Tables* host_tables) {
  return cudaMemcpyToSymbol(c_mel_fw, host_tables->fw, sizeof(float) * mel::MAX_NNZ);
}

namespace {

using namespace mel;

constexpr int MAX_DEFERRED_FWD = 16;
struct MelSmem {
  float slab[SLAB_ROWS * SLAB_PITCH];   // 23.9 KB  [hop row][160 (+ alignment shift)]
  float2 E[FB * K1 * E_PITCH];          // 59.9 KB  [frame * 13 + k1][16 j (+2 pad)]
  float P[FB * P_PITCH];                // 26.8 KB  [frame][bin]
  Item desc[2];                         // current / next work item (the next one arrives by cp.async)
  Item deferred[MAX_DEFERRED_FWD];          // clamp items whose clip was not finished when their ticket came up
  int idx[2];
  int n_deferred;
  int run_clamp;                        // broadcast: 0 = nothing to do, 1 = clamp sm.clamp with sm.floor_v
  Item clamp;
  float red[THREADS / 32];
  float floor_v;
  unsigned long long mbar;              // completion of the bulk copies of a slab
};
constexpr int MAX_DEFERRED = MAX_DEFERRED_FWD;
static_assert(sizeof(Item) == 48, "Item is copied as three 16-byte cp.async transfers");
static_assert((SLAB_ROWS - 1) * SLAB_PITCH + ROW_COPY <= SLAB_ROWS * SLAB_PITCH, "bulk row copies stay inside the slab");

// ---- mel phase: filters [M, MEnd) for the lane's frame, fully unrolled from the compiled-in structure --------------
__device__ __forceinline__ float lg2_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int M, int MEnd>
struct MelRun {
  static __device__ __forceinline__ void run(const float* __restrict__ prow, float* __restrict__ optr, long long ld, bool live, float& vmax) {
    constexpr int cnt = kMelCnt[M], ptr = kMelPtr[M], lo = kMelLo[M];
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < cnt; ++j) acc = fmaf(c_mel_fw[ptr + j], prow[lo + j], acc);
    const float v = lg2_fast(fmaxf(acc, 1e-10f)) * 0.30102999566398120f;  // log10
    if (live) *optr = v;
    vmax = fmaxf(vmax, v);
    MelRun<M + 1, MEnd>::run(prow, optr + ld, ld, live, vmax);
  }
};
template <int MEnd>
struct MelRun<MEnd, MEnd> {
  static __device__ __forceinline__ void run(const float*, float*, long long, bool, float&) {}
};
template <int G>
__device__ __forceinline__ void mel_group(const float* __restrict__ prow, float* __restrict__ ocol, long long ld, bool live, float& vmax) {
  constexpr int m0 = kMelGroup[G], m1 = kMelGroup[G + 1];
  MelRun<m0, m1>::run(prow, ocol + m0 * ld, ld, live, vmax);
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// clip-relative sample index of slab position 0
__device__ __forceinline__ int slab_s0(const Item& it) { return it.frame0 * HOP - N_FFT / 2; }

// Warp 0: start the bulk copies of an item's slab (34 hop rows of 164 floats from the 16-byte aligned address at or
// below the row's first sample; stage 1 adds the 0..3 float shift back).  One row per lane, so the 34 copies leave in two
// instructions instead of a 34-trip loop on the thread every barrier waits for.  The expect_tx arrival is the phase's only
// pending arrival, so rows that complete before it is posted cannot flip the phase early.
__device__ __forceinline__ void issue_slab_bulk(MelSmem& sm, const float* __restrict__ pcm, const Item& it, int lane) {
  const float* src = pcm + it.pcm_off + slab_s0(it);
  src -= (reinterpret_cast<uintptr_t>(src) >> 2) & 3;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy accesses of the slab vs the async writes
  if (lane == 0) mbar_expect_tx(&sm.mbar, SLAB_ROWS * ROW_COPY * 4);
  for (int r = lane; r < SLAB_ROWS; r += 32) bulk_g2s(sm.slab + r * SLAB_PITCH, src + r * HOP, ROW_COPY * 4, &sm.mbar);
}
// warp 0, after lane 0 decided `go`: broadcast it and issue
__device__ __forceinline__ void warp0_issue(MelSmem& sm, const float* __restrict__ pcm, const Item& it, int go, int lane) {
  go = __shfl_sync(0xffffffffu, go, 0);
  __syncwarp();  // lane 0's descriptor (cp.async + wait) is visible to the other lanes
  if (go) issue_slab_bulk(sm, pcm, it, lane);
}

__global__ void __launch_bounds__(THREADS, 2) logmel_kernel(const float* __restrict__ pcm, const Item* __restrict__ items, int n_items,
                                                            const Tables* __restrict__ tables, float* __restrict__ out, long long ld,
                                                            unsigned int* __restrict__ ticket, unsigned int* __restrict__ clip_done,
                                                            unsigned int* __restrict__ clip_max) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  MelSmem& sm = *reinterpret_cast<MelSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // per-thread constants, loaded once for all items
  const int j = tid & 15;                 // stage 1: sample phase j of frame (tid >> 4) [+16]
  float win[25];
#pragma unroll
  for (int m = 0; m < 25; ++m) win[m] = __ldg(tables->window + 16 * m + j);
  const int k1 = tid % K1;                // stage 2: output family k1 of frame (tid / 13) [+16], threads 0..207
  float2 tw[16];
#pragma unroll
  for (int jj = 1; jj < 16; ++jj) tw[jj] = __ldg(&tables->tw[k1][jj]);

  // Work distribution: thread 0 keeps two tickets ahead (t_next: descriptor being fetched, t_after: atomic in flight), so
  // neither the atomic nor the descriptor load nor the PCM copy of the next item is ever waited for.
  int t_next = 0, t_after = 0;
  if (tid == 0) {
    mbar_init(&sm.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int t0 = static_cast<int>(atomicAdd(ticket, 1u));
    t_next = static_cast<int>(atomicAdd(ticket, 1u));
    sm.idx[0] = t0;
    sm.n_deferred = 0;
    sm.run_clamp = 0;
    if (t0 < n_items) sm.desc[0] = items[t0];
  }
  __syncthreads();
  if (warp == 0) warp0_issue(sm, pcm, sm.desc[0], lane == 0 && sm.idx[0] < n_items && sm.desc[0].kind == 0 && sm.desc[0].bulk, lane);
  int slot = 0;
  uint32_t parity = 0;

  // clamp + rescale one tile of a finished clip, in place (the tile is still L2-resident); all threads
  auto clamp_tile = [&](const Item& c, float floor_v) {
    if (tid < c.n_frames) {
      float* p = out + c.col0 + c.frame0 + tid;
      // 16 loads in flight per thread before the first store: the compiler cannot hoist a load above a store to `out`
#pragma unroll 1
      for (int m0 = 0; m0 < N_MELS; m0 += 16, p += 16 * ld) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldcg(p + i * ld);
#pragma unroll
        for (int i = 0; i < 16; ++i) __stcg(p + i * ld, (fmaxf(v[i], floor_v) + 4.0f) * 0.25f);
      }
    }
  };
  // thread 0: is the clip of clamp item c finished?  (acquire: its log-mel stores and max are then visible)
  auto clamp_ready = [&](const Item& c) { return ld_acquire_u32(clip_done + c.clip) >= static_cast<unsigned int>(c.need); };
  auto clamp_floor = [&](const Item& c) { return ordered_to_float(ld_acquire_u32(clip_max + c.clip)) - 8.0f; };

  for (;;) {
    const int idx = sm.idx[slot];
    if (idx >= n_items) break;
    const Item it = sm.desc[slot];
    if (tid == 0) {
      t_after = static_cast<int>(atomicAdd(ticket, 1u));
      sm.idx[slot ^ 1] = t_next;
      if (t_next < n_items) {
        const char* src = reinterpret_cast<const char*>(items + t_next);
        char* dst = reinterpret_cast<char*>(&sm.desc[slot ^ 1]);
        cp_async16(dst, src);
        cp_async16(dst + 16, src + 16);
        cp_async16(dst + 32, src + 32);
      }
    }

    if (it.kind == 1) {
      // ---------------- clamp item: run it if its clip is finished, otherwise set it aside (never block) ----------------
      if (tid == 0) {
        if (sm.n_deferred == MAX_DEFERRED) {  // (practically never) no room to defer: wait for the oldest deferred one
          while (!clamp_ready(sm.deferred[0])) __nanosleep(100);
        }
        if (sm.n_deferred > 0 && clamp_ready(sm.deferred[0])) {
          // keep FIFO order: run the oldest deferred item now and queue this one behind the rest
          sm.clamp = sm.deferred[0];
          for (int i = 1; i < sm.n_deferred; ++i) sm.deferred[i - 1] = sm.deferred[i];
          sm.deferred[sm.n_deferred - 1] = it;
          sm.floor_v = clamp_floor(sm.clamp);
          sm.run_clamp = 1;
        } else if (sm.n_deferred == 0 && clamp_ready(it)) {
          sm.clamp = it;
          sm.floor_v = clamp_floor(it);
          sm.run_clamp = 1;
        } else {
          sm.deferred[sm.n_deferred++] = it;
          sm.run_clamp = 0;
        }
      }
      if (warp == 0) {
        // the slab is idle during a clamp item: start the next item's copy right away
        int go = 0;
        if (lane == 0) {
          cp_async_wait_all();
          go = t_next < n_items && sm.desc[slot ^ 1].kind == 0 && sm.desc[slot ^ 1].bulk;
          t_next = t_after;
        }
        warp0_issue(sm, pcm, sm.desc[slot ^ 1], go, lane);
      }
      __syncthreads();
      if (sm.run_clamp) clamp_tile(sm.clamp, sm.floor_v);
      __syncthreads();  // sm.clamp / floor_v / descriptor slots are rewritten by later iterations
      slot ^= 1;
      continue;
    }

    // ---------------- PCM slab in smem ----------------
    int shift = 0;
    if (it.bulk) {
      shift = static_cast<int>((reinterpret_cast<uintptr_t>(pcm + it.pcm_off + slab_s0(it)) >> 2) & 3);
      mbar_wait(&sm.mbar, parity);
      parity ^= 1;
    } else {
      // clip edges: reflect padding resolved by index mirroring; positions no valid frame touches are zero.
      // Loads are issued in batches of 11 before their stores so that their latencies overlap.
      const float* clip = pcm + it.pcm_off;
      const int s0 = slab_s0(it);
      const int need = (it.n_frames - 1) * HOP + N_FFT;
#pragma unroll 1
      for (int base = 0; base < 22; base += 11) {
        float v[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const int u = tid + (base + i) * THREADS;
          v[i] = u < need ? __ldg(clip + reflect_index(s0 + u, it.n_samples)) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const int u = tid + (base + i) * THREADS;
          if (u < SLAB_ROWS * HOP) {
            const int r = u / HOP;
            sm.slab[r * SLAB_PITCH + (u - r * HOP)] = v[i];
          }
        }
      }
      __syncthreads();
    }

    // ---------------- stage 1: 16 real 25-point DFTs per frame ----------------
#pragma unroll 1
    for (int rnd = 0; rnd < 2; ++rnd) {
      const int f = (tid >> 4) + 16 * rnd;
      const float* x = sm.slab + shift + f * SLAB_PITCH + j;
      float v[25];
#pragma unroll
      for (int m = 0; m < 25; ++m) v[m] = x[(m / 10) * SLAB_PITCH + 16 * (m % 10)] * win[m];
      float2 V[K1];
      rdft25(v, V);
      float2* e = sm.E + f * K1 * E_PITCH + j;
#pragma unroll
      for (int q = 0; q < K1; ++q) e[q * E_PITCH] = V[q];
    }
    __syncthreads();  // E complete, slab free

    if (warp == 0) {
      int go = 0;
      if (lane == 0) {
        cp_async_wait_all();
        go = t_next < n_items && sm.desc[slot ^ 1].kind == 0 && sm.desc[slot ^ 1].bulk;
        t_next = t_after;
      }
      warp0_issue(sm, pcm, sm.desc[slot ^ 1], go, lane);
    }

    // ---------------- stage 2: 13 complex 16-point FFTs per frame -> power ----------------
    if (tid < 16 * K1) {
#pragma unroll 1
      for (int rnd = 0; rnd < 2; ++rnd) {
        const int row = tid + 16 * K1 * rnd;      // = frame * 13 + k1
        const int f = row / K1;
        const float4* e = reinterpret_cast<const float4*>(sm.E + row * E_PITCH);
        float2 z[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 q = e[c];
          z[2 * c] = make_float2(q.x, q.y);
          z[2 * c + 1] = make_float2(q.z, q.w);
        }
#pragma unroll
        for (int jj = 1; jj < 16; ++jj) z[jj] = cmul(z[jj], tw[jj]);
        fft16(z);
        float* prow = sm.P + f * P_PITCH;
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float pw = fmaf(z[k2].x, z[k2].x, z[k2].y * z[k2].y);
          if (stage2_unique(k1, k2)) prow[stage2_bin(k1, k2)] = pw;
        }
      }
    }
    __syncthreads();

    // ---------------- mel + log10: warp = filter group, lane = frame ----------------
    {
      const bool live = lane < it.n_frames;
      const float* prow = sm.P + lane * P_PITCH;
      float* ocol = out + it.col0 + it.frame0 + lane;
      float vmax = -INFINITY;
      switch (warp) {
        case 0: mel_group<0>(prow, ocol, ld, live, vmax); break;
        case 1: mel_group<1>(prow, ocol, ld, live, vmax); break;
        case 2: mel_group<2>(prow, ocol, ld, live, vmax); break;
        case 3: mel_group<3>(prow, ocol, ld, live, vmax); break;
        case 4: mel_group<4>(prow, ocol, ld, live, vmax); break;
        case 5: mel_group<5>(prow, ocol, ld, live, vmax); break;
        case 6: mel_group<6>(prow, ocol, ld, live, vmax); break;
        default: mel_group<7>(prow, ocol, ld, live, vmax); break;
      }
      vmax = warp_max(live ? vmax : -INFINITY);
      if (lane == 0) sm.red[warp] = vmax;
    }
    const int n_def = sm.n_deferred;  // read before the barrier: thread 0 only changes it after one, so every thread agrees
    __syncthreads();
    if (tid == 0) {
      float v = sm.red[0];
#pragma unroll
      for (int i = 1; i < THREADS / 32; ++i) v = fmaxf(v, sm.red[i]);
      // the barrier ordered every thread's log-mel stores before this point; the fence makes them (and the max) visible
      // device-wide before the clip's done counter moves
      atomicMax(clip_max + it.clip, float_to_ordered(v));
      __threadfence();
      atomicAdd(clip_done + it.clip, 1u);
    }
    if (n_def > 0) {
      // a clamp item is waiting: run the oldest one if its clip has finished meanwhile
      if (tid == 0) {
        sm.run_clamp = 0;
        if (clamp_ready(sm.deferred[0])) {
          sm.clamp = sm.deferred[0];
          for (int i = 1; i < sm.n_deferred; ++i) sm.deferred[i - 1] = sm.deferred[i];
          --sm.n_deferred;
          sm.floor_v = clamp_floor(sm.clamp);
          sm.run_clamp = 1;
        }
      }
      __syncthreads();
      if (sm.run_clamp) clamp_tile(sm.clamp, sm.floor_v);
      __syncthreads();
    }
    slot ^= 1;
  }

  // no tickets left: every frame item is running or finished, so the remaining deferred clamps can simply be waited for
  for (;;) {
    __syncthreads();
    if (sm.n_deferred == 0) break;
    if (tid == 0) {
      sm.clamp = sm.deferred[sm.n_deferred - 1];
      while (!clamp_ready(sm.clamp)) __nanosleep(100);
      sm.floor_v = clamp_floor(sm.clamp);
    }
    __syncthreads();
    clamp_tile(sm.clamp, sm.floor_v);
    __syncthreads();
    if (tid == 0) --sm.n_deferred;
  }
}

// ===========================================================================================================================
// v2 (round 2, QASR_MEL=v2; NOT the default -- measured no faster): the same arithmetic, re-scheduled to
// get rid of the barriers.  The first kernel synchronises the CTA three times
// per 32-frame item (slab ready, exchange complete, power complete) and ncu showed 4.75 barrier-stall cycles per issued
// instruction at 35 % issue utilisation.  Here both FFT stages are WARP-synchronous: a warp owns 4 of the item's 32 frames, reads
// their samples straight from global memory (25 strided loads per lane, L1-resident: consecutive frames overlap by 60 %), runs
// stage 1 (lane = (frame, j)), exchanges through its private 3.7 KB of shared memory behind a __syncwarp, runs stage 2
// (lane = (frame, k1), 26 of 32 lanes) and leaves the power rows in a CTA-wide buffer.  ONE __syncthreads per item then starts the
// mel phase unchanged (warp = filter group, lane = frame: compile-time unrolled filters, constant-bank weights, 128-byte stores).
// The power buffer is double-buffered, so that barrier is the only one; clip completion is counted per warp (fence + atomic by
// lane 0 after a __syncwarp), not per CTA.  No slab, no bulk copies, no mbarrier.
struct MelSmem2 {
  float2 E[THREADS / 32][2][K1][E_PITCH];   // 29.9 KB  per warp: [frame of the pair][k1][j]
  float P[2][FB][P_PITCH];                  // 53.5 KB  double-buffered power rows of the item's 32 frames
  Item desc[2];
  Item deferred[MAX_DEFERRED_FWD];
  int idx[2];
  int n_deferred;
  int run_clamp;
  Item clamp;
  float floor_v;
};
constexpr int kWarpsPerCta = THREADS / 32;
static_assert(FB == 4 * kWarpsPerCta, "a warp owns four frames of an item");

__global__ void __launch_bounds__(THREADS, 2) logmel_kernel_v2(const float* __restrict__ pcm, const Item* __restrict__ items, int n_items,
                                                               const Tables* __restrict__ tables, float* __restrict__ out, long long ld,
                                                               unsigned int* __restrict__ ticket, unsigned int* __restrict__ clip_done,
                                                               unsigned int* __restrict__ clip_max) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  MelSmem2& sm = *reinterpret_cast<MelSmem2*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // per-lane constants: stage 1 lane = (fr, j), stage 2 lane = (fr2, k1)
  const int fr = lane >> 4, j = lane & 15;
  float win[25];
#pragma unroll
  for (int m = 0; m < 25; ++m) win[m] = __ldg(tables->window + 16 * m + j);
  const int fr2 = lane / K1, k1 = lane - fr2 * K1;     // lanes 26..31: fr2 == 2 -> idle in stage 2
  float2 tw[16];
#pragma unroll
  for (int jj = 1; jj < 16; ++jj) tw[jj] = __ldg(&tables->tw[k1][jj]);

  int t_next = 0, t_after = 0;
  if (tid == 0) {
    const int t0 = static_cast<int>(atomicAdd(ticket, 1u));
    t_next = static_cast<int>(atomicAdd(ticket, 1u));
    sm.idx[0] = t0;
    sm.n_deferred = 0;
    sm.run_clamp = 0;
    if (t0 < n_items) sm.desc[0] = items[t0];
  }
  __syncthreads();
  int slot = 0, pb = 0;

  auto clamp_tile = [&](const Item& c, float floor_v) {
    if (tid < c.n_frames) {
      float* p = out + c.col0 + c.frame0 + tid;
#pragma unroll 1
      for (int m0 = 0; m0 < N_MELS; m0 += 16, p += 16 * ld) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldcg(p + i * ld);
#pragma unroll
        for (int i = 0; i < 16; ++i) __stcg(p + i * ld, (fmaxf(v[i], floor_v) + 4.0f) * 0.25f);
      }
    }
  };
  // a clip is finished when every WARP of every frame item of it has stored its rows (need counts items)
  auto clamp_ready = [&](const Item& c) { return ld_acquire_u32(clip_done + c.clip) >= static_cast<unsigned int>(c.need) * kWarpsPerCta; };
  auto clamp_floor = [&](const Item& c) { return ordered_to_float(ld_acquire_u32(clip_max + c.clip)) - 8.0f; };

  for (;;) {
    const int idx = sm.idx[slot];
    if (idx >= n_items) break;
    const Item it = sm.desc[slot];
    if (tid == 0) {   // keep two tickets ahead: the next descriptor arrives by cp.async while this item is computed
      t_after = static_cast<int>(atomicAdd(ticket, 1u));
      sm.idx[slot ^ 1] = t_next;
      if (t_next < n_items) {
        const char* src = reinterpret_cast<const char*>(items + t_next);
        char* dst = reinterpret_cast<char*>(&sm.desc[slot ^ 1]);
        cp_async16(dst, src);
        cp_async16(dst + 16, src + 16);
        cp_async16(dst + 32, src + 32);
      }
      t_next = t_after;
    }

    if (it.kind == 1) {
      // ---------------- clamp item: run it if its clip is finished, otherwise set it aside (never block) ----------------
      if (tid == 0) {
        if (sm.n_deferred == MAX_DEFERRED) {
          while (!clamp_ready(sm.deferred[0])) __nanosleep(100);
        }
        if (sm.n_deferred > 0 && clamp_ready(sm.deferred[0])) {
          sm.clamp = sm.deferred[0];
          for (int i = 1; i < sm.n_deferred; ++i) sm.deferred[i - 1] = sm.deferred[i];
          sm.deferred[sm.n_deferred - 1] = it;
          sm.floor_v = clamp_floor(sm.clamp);
          sm.run_clamp = 1;
        } else if (sm.n_deferred == 0 && clamp_ready(it)) {
          sm.clamp = it;
          sm.floor_v = clamp_floor(it);
          sm.run_clamp = 1;
        } else {
          sm.deferred[sm.n_deferred++] = it;
          sm.run_clamp = 0;
        }
        cp_async_wait_all();
      }
      __syncthreads();
      if (sm.run_clamp) clamp_tile(sm.clamp, sm.floor_v);
      __syncthreads();
      slot ^= 1;
      continue;
    }

    // ---------------- stages 1 + 2, warp-synchronous: this warp's frames 4 w .. 4 w + 3 of the item, two at a time ----------------
    const float* clip = pcm + it.pcm_off;
    float2* Ew = &sm.E[warp][0][0][0];
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
      {
        const int f = warp * 4 + 2 * r + fr;                       // frame of the item this lane works on in stage 1
        const int s0 = (it.frame0 + f) * HOP - N_FFT / 2 + j;      // clip-relative index of its first sample (m = 0)
        float v[25];
        if (f < it.n_frames) {
          if (s0 - j >= 0 && s0 - j + N_FFT <= it.n_samples) {     // interior frame (uniform over the 16 lanes of a frame)
#pragma unroll
            for (int m = 0; m < 25; ++m) v[m] = __ldg(clip + s0 + 16 * m) * win[m];
          } else {                                                 // clip edge: reflect padding by index mirroring
#pragma unroll
            for (int m = 0; m < 25; ++m) v[m] = __ldg(clip + reflect_index(s0 + 16 * m, it.n_samples)) * win[m];
          }
        } else {
#pragma unroll
          for (int m = 0; m < 25; ++m) v[m] = 0.f;
        }
        float2 V[K1];
        rdft25(v, V);
        float2* e = Ew + fr * K1 * E_PITCH + j;
#pragma unroll
        for (int q = 0; q < K1; ++q) e[q * E_PITCH] = V[q];
      }
      __syncwarp();
      if (fr2 < 2) {
        const int f = warp * 4 + 2 * r + fr2;
        const float4* e = reinterpret_cast<const float4*>(Ew + (fr2 * K1 + k1) * E_PITCH);
        float2 z[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 q = e[c];
          z[2 * c] = make_float2(q.x, q.y);
          z[2 * c + 1] = make_float2(q.z, q.w);
        }
#pragma unroll
        for (int jj = 1; jj < 16; ++jj) z[jj] = cmul(z[jj], tw[jj]);
        fft16(z);
        float* prow = &sm.P[pb][f][0];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const float pw = fmaf(z[k2].x, z[k2].x, z[k2].y * z[k2].y);
          if (stage2_unique(k1, k2)) prow[stage2_bin(k1, k2)] = pw;
        }
      }
      __syncwarp();   // E is rewritten by the next pair
    }
    if (tid == 0) cp_async_wait_all();   // the next descriptor has landed before the barrier publishes it
    __syncthreads();                     // the item's 32 power rows are complete (and P[pb ^ 1] is free: everyone finished the last mel phase)

    // ---------------- mel + log10: warp = filter group, lane = frame ----------------
    {
      const bool live = lane < it.n_frames;
      const float* prow = &sm.P[pb][lane][0];
      float* ocol = out + it.col0 + it.frame0 + lane;
      float vmax = -INFINITY;
      switch (warp) {
        case 0: mel_group<0>(prow, ocol, ld, live, vmax); break;
        case 1: mel_group<1>(prow, ocol, ld, live, vmax); break;
        case 2: mel_group<2>(prow, ocol, ld, live, vmax); break;
        case 3: mel_group<3>(prow, ocol, ld, live, vmax); break;
        case 4: mel_group<4>(prow, ocol, ld, live, vmax); break;
        case 5: mel_group<5>(prow, ocol, ld, live, vmax); break;
        case 6: mel_group<6>(prow, ocol, ld, live, vmax); break;
        default: mel_group<7>(prow, ocol, ld, live, vmax); break;
      }
      vmax = warp_max(live ? vmax : -INFINITY);
      __syncwarp();   // every lane's stores are ordered before lane 0's fence
      if (lane == 0) {
        atomicMax(clip_max + it.clip, float_to_ordered(vmax));
        __threadfence();
        atomicAdd(clip_done + it.clip, 1u);
      }
    }
    slot ^= 1;
    pb ^= 1;
  }

  // no tickets left: every frame item is running or finished, so the remaining deferred clamps can simply be waited for
  for (;;) {
    __syncthreads();
    if (sm.n_deferred == 0) break;
    if (tid == 0) {
      sm.clamp = sm.deferred[sm.n_deferred - 1];
      while (!clamp_ready(sm.clamp)) __nanosleep(100);
      sm.floor_v = clamp_floor(sm.clamp);
    }
    __syncthreads();
    clamp_tile(sm.clamp, sm.floor_v);
    __syncthreads();
    if (tid == 0) --sm.n_deferred;
  }
}

}  // namespace

cudaError_t launch_logmel(const float* pcm, const mel::Item* items, int n_items, const mel::Tables* tables, float* mel_out,
                          long long mel_ld, unsigned int* counters, int n_clips, int num_sms, int variant, cudaStream_t stream) {
  if (n_items == 0) return cudaSuccess;
  // counters: [0] ticket, [1, 1 + n_clips) done, [1 + n_clips, 1 + 2 n_clips) max (ordered-uint encoding; 0 = below every float)
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(unsigned int) * (1 + 2 * static_cast<size_t>(n_clips)), stream);
  if (e != cudaSuccess) return e;
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  if (variant == 2) {
    e = cudaFuncSetAttribute(logmel_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(MelSmem2)));
    if (e != cudaSuccess) return e;
    logmel_kernel_v2<<<grid, mel::THREADS, sizeof(MelSmem2), stream>>>(pcm, items, n_items, tables, mel_out, mel_ld, counters, counters + 1,
                                                                       counters + 1 + n_clips);
    return cudaGetLastError();
  }
  e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(MelSmem)));
  if (e != cudaSuccess) return e;
  logmel_kernel<<<grid, mel::THREADS, sizeof(MelSmem), stream>>>(pcm, items, n_items, tables, mel_out, mel_ld, counters, counters + 1,
                                                                 counters + 1 + n_clips);
  return cudaGetLastError();
}

}  // namespace qasr

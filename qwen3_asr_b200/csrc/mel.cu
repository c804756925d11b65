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
#include <cstdio>
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
// v3 (round 2, the default; QASR_MEL=v1 selects the kernel above): the same transform (results equal to float32 rounding: v3 keeps
// log2 of the mel power until its clamp pass applies log10(2), the 1e-10 clamp and the max-8 floor at once), re-scheduled so that NO warp ever waits
// at a CTA-wide barrier.  ncu on v1: 35 % issue utilisation, 4.75 barrier-stall cycles per issued instruction (three
// __syncthreads per 32-frame item at 16 warps per SM).  Here
//   * items are still claimed through the global ticket (a static i -> CTA i mod grid assignment let CTAs drift apart and spin on
//     clips other CTAs had not finished), but by ONE lane, two items ahead, with the 64-byte descriptor bulk-copied into a ring
//     behind an mbarrier: no broadcast barrier, and every warp walks the ring on its own;
//   * both FFT stages are warp-synchronous: a warp owns 4 of the item's 32 frames (two rounds of two), stage 1 lane = (frame, j)
//     reads its 25 strided samples from global memory (L1-resident: consecutive frames overlap by 60 %) ONE ROUND AHEAD into
//     registers, applies the stage-2 twiddle to its own 12 outputs (32 busy lanes instead of stage 2's 26), and exchanges through
//     the warp's private 3.7 KB of shared memory behind a __syncwarp;
//   * the power rows go to a ring of three 32-frame tiles guarded by full / empty mbarriers (8 arrivals each, one per warp); the
//     mel phase (warp = filter group, lane = frame) of item n runs after the FFT of item n + 1, so a warp only ever waits for
//     the slowest warp of the PREVIOUS item -- a whole item of slack;
//   * the mel phase reads the power row with 16-byte loads (row pitch 204: conflict-free for lane = frame) interleaved with the
//     compile-time unrolled filters, and takes the clip maximum BEFORE the logarithm (lg2 is monotone) with FMNMX3;
//   * the max-8 clamp is no longer a work item of its own: every item carries the 32-frame tile of an item `lag` tickets before
//     it (its clip finished, the tile still in L2) and each warp rewrites 16 of the tile's 128 rows.  The clip's counter is polled
//     with relaxed loads / a 16-byte bulk copy and the tile read with .cg loads: an acquire would invalidate the L1 that holds the
//     re-read PCM, and an ordinary load would share a scoreboard with the sample loads in flight;
//   * a tile's completion signal (clip maximum, then a release increment: ~1 us of round trip for the warp that sends it) is owed
//     by warp (tile mod 8) and leaves from a later FFT, as soon as the tile's `empty` barrier is complete.
// Dependencies point to strictly smaller tickets, a warp runs its pending mel phase and sends what it owes before it waits for a
// clip, and a claimed ticket belongs to a resident CTA: the grid cannot deadlock whatever subset of it is resident.
constexpr int TW_PITCH = 14;               // float2 per twiddle row (13 used): 112 B, same bank pattern as the window rows
constexpr int WIN_PITCH = 28;              // window row pitch: 16-byte loads, conflict-free for the 16 lanes of a frame (28 = -4 mod 32)
constexpr int D3_RING = 8;                 // descriptor ring (items in flight per CTA: the current one, two published ahead, slack)
struct MelSmem3 {
  float2 E[THREADS / 32][2][K1][E_PITCH];   // 29.9 KB  per warp: [frame of the pair][k1][j]
  float P[P3_RING][FB][P3_PITCH];           // 78.3 KB  power rows of three items
  float win[16][WIN_PITCH];                 // window[16 m + j] at [j][m]: read at use, 25 registers less than keeping it
  float2 tw[16][TW_PITCH];                  // W400^(q j) at [j][q]: the stage-2 twiddle, applied by stage 1 to its own outputs
  Item3 ring[D3_RING];                      // descriptors in ticket order, bulk-copied behind ring_full
  uint4 poll[THREADS / 32];                 // per warp: the 16 bytes around the polled clip counter
  unsigned long long full[P3_RING];         // all 8 warps have written their 4 rows
  unsigned long long empty[P3_RING];        // all 8 warps have finished their filters
  unsigned long long ring_full[D3_RING];    // the descriptor has landed (or the slot says "no work left")
  unsigned long long ring_empty[D3_RING];   // all 8 warps are done with the item
  unsigned long long poll_bar[THREADS / 32];// the warp's poll copy has landed
  unsigned int mel_amax[P3_RING];           // largest clamped mel power of the tile (float bits)
};
constexpr int kWarpsPerCta = THREADS / 32;
static_assert(FB == 4 * kWarpsPerCta, "a warp owns four frames of an item");
static_assert(N_MELS == 16 * kWarpsPerCta, "a warp clamps sixteen rows of a tile");
static_assert(sizeof(Item3) == 64, "Item3 is bulk-copied: a multiple of 16 bytes");

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// Bounded wait for the v3 kernel: a protocol bug traps (-> a CUDA error the host reports) instead of hanging the GPU.
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, uint32_t parity);
// (no printf: a call in this kernel costs the register allocator 8 % of the kernel's speed; the trap surfaces as a launch failure)
__device__ __forceinline__ void mel_wait_timeout(int) { __trap(); }
// mbarrier wait of the v3 kernel.  Product build: the three-instruction try_wait loop -- counting the attempts (to trap on a protocol
// bug instead of hanging) costs this issue-bound kernel 3.5 % (A/B on one box: 0.528 vs 0.547 ms), with or without a suspend-time
// hint; -DQASR_MEL_BOUNDED=1 builds the counted version for bring-up.  The one wait that depends on OTHER CTAs (a clip's completion
// counter) is always bounded.
__device__ __forceinline__ void mbar_wait3(unsigned long long* bar, uint32_t parity, int what) {
#ifdef QASR_MEL_BOUNDED
  const uint32_t addr = smem_addr(bar);
  for (unsigned int spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 27)) mel_wait_timeout(what);
  }
#else
  (void)what;
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, uint32_t parity) {   // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// the 394 mel weights travel as a kernel parameter: bank-0 constants are direct FFMA operands, a __constant__ array costs one
// uniform load (LDCU) per multiply-add
struct MelWeights {
  float w[MAX_NNZ - 112];
};
static_assert(kMelNnz <= MAX_NNZ - 112, "mel weights fit the parameter block");
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// filters [M, MEnd) of the lane's frame; NB = first bin of the power row not yet in p[] (rows are loaded 4 bins at a time, just
// before the first filter that needs them, so only a sliding window of the row is live)
template <int M0, int M, int MEnd, int B0, int NB>
struct MelRun3 {
  template <int NP>
  static __device__ __forceinline__ void run(const MelWeights& fw, const float* __restrict__ prow, float (&p)[NP], float* __restrict__ optr, int ld,
                                             float& amax, float carry) {
    constexpr int cnt = kMelCnt[M], ptr = kMelPtr[M], lo = kMelLo[M], hi = lo + cnt;
    constexpr int NB2 = NB < hi ? ((hi + 3) & ~3) : NB;
    // keep the row loads where they are: hoisted to the top of the group they would hold up to 52 registers while 41 prefetched
    // values (the next round's samples, the clamp tile) wait in theirs
    if constexpr (((M - M0) & 3) == 0) asm volatile("" ::: "memory");
#pragma unroll
    for (int b = NB; b < NB2; b += 4) {
      const float4 q = *reinterpret_cast<const float4*>(prow + b);
      p[b - B0] = q.x; p[b - B0 + 1] = q.y; p[b - B0 + 2] = q.z; p[b - B0 + 3] = q.w;
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < cnt; ++j) acc = fmaf(fw.w[ptr + j], p[lo + j - B0], acc);
    const float a = acc;
    // log2 of the mel power, un-clamped and un-scaled (the 1e-10 clamp, the factor log10(2) and the max-8 floor are all applied by the
    // clamp pass, which rewrites every value anyway): two instructions per filter less, and 4 KB less unrolled code;  M * ld < 2^31
    optr[static_cast<long long>(M * ld)] = lg2_fast(a);
    // the clip maximum is taken over the clamped power (lg2 is monotone), two filters per FMNMX3
    if constexpr (((M - M0) & 1) != 0) {
      amax = max3(amax, carry, a);
      MelRun3<M0, M + 1, MEnd, B0, NB2>::run(fw, prow, p, optr, ld, amax, 0.f);
    } else if constexpr (M + 1 == MEnd) {
      amax = fmaxf(amax, a);
    } else {
      MelRun3<M0, M + 1, MEnd, B0, NB2>::run(fw, prow, p, optr, ld, amax, a);
    }
  }
};
template <int M0, int MEnd, int B0, int NB>
struct MelRun3<M0, MEnd, MEnd, B0, NB> {
  template <int NP>
  static __device__ __forceinline__ void run(const MelWeights&, const float*, float (&)[NP], float*, int, float&, float) {}
};
template <int G>
__device__ __forceinline__ void mel_group3(const MelWeights& fw, const float* __restrict__ prow, float* __restrict__ ocol, int ld, float& amax) {
  constexpr int m0 = kMelGroup[G], m1 = kMelGroup[G + 1];
  constexpr int b0 = kMelLo[m0] & ~3, b1 = (kMelLo[m1 - 1] + kMelCnt[m1 - 1] + 3) & ~3;
  static_assert(b1 <= P3_PITCH, "16-byte row loads stay inside the row");
  float p[b1 - b0];
  MelRun3<m0, m0, m1, b0, b0>::run(fw, prow, p, ocol, ld, amax, 0.f);
}

__global__ void __launch_bounds__(THREADS, 2) logmel_kernel_v3(const float* __restrict__ pcm, const Item3* __restrict__ items, int n_items,
                                                               const Tables* __restrict__ tables, float* __restrict__ out, int ld,
                                                               unsigned int* __restrict__ ticket, unsigned int* __restrict__ clip_done,
                                                               unsigned int* __restrict__ clip_max, float* __restrict__ dump,
                                                               const __grid_constant__ MelWeights fw) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  MelSmem3& sm = *reinterpret_cast<MelSmem3*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < P3_RING; ++s) {
      mbar_init(&sm.full[s], kWarpsPerCta);
      mbar_init(&sm.empty[s], kWarpsPerCta);
      sm.mel_amax[s] = 0;
    }
#pragma unroll
    for (int s = 0; s < D3_RING; ++s) {
      mbar_init(&sm.ring_full[s], 1);
      mbar_init(&sm.ring_empty[s], kWarpsPerCta);
    }
#pragma unroll
    for (int s = 0; s < kWarpsPerCta; ++s) mbar_init(&sm.poll_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 16 * 25; i += THREADS) sm.win[i & 15][i >> 4] = __ldg(tables->window + i);
  for (int i = tid; i < 16 * K1; i += THREADS) sm.tw[i & 15][i >> 4] = __ldg(&tables->tw[i >> 4][i & 15]);
  // stage 1 lane = (fr, j), stage 2 lane = (fr2, k1)
  const int fr = lane >> 4, j = lane & 15;
  const int fr2 = lane / K1, k1 = lane - fr2 * K1;     // lanes 26..31: fr2 == 2 -> idle in stage 2
  __syncthreads();                                      // the only CTA-wide barrier: mbarriers and tables are set up

  // ---- work distribution: lane 0 of warp 0 claims tickets two items ahead and has the descriptors bulk-copied into a ring ----
  int t_next = 0, t_after = 0;
  auto publish = [&](int seq) {                         // lane 0 of warp 0 only
    const int slot = seq % D3_RING, use = seq / D3_RING;
    if (use > 0) mbar_wait3(&sm.ring_empty[slot], (use - 1) & 1, 1);   // every warp is done with the slot's previous item
    const int t = t_next;
    t_next = t_after;
    t_after = static_cast<int>(atomicAdd(ticket, 1u));  // consumed by the next call: the round trip is never waited for
    if (t < n_items) {
      mbar_expect_tx(&sm.ring_full[slot], sizeof(Item3));
      bulk_g2s(&sm.ring[slot], items + t, sizeof(Item3), &sm.ring_full[slot]);
    } else {
      sm.ring[slot].n_frames = -1;                      // no work left
      mbar_arrive(&sm.ring_full[slot]);
    }
  };
  if (tid == 0) {
    t_next = static_cast<int>(atomicAdd(ticket, 2u));   // the first two tickets at once; every publish claims one more
    t_after = t_next + 1;
    publish(0);
    publish(1);
  }

  // Completion signal of a tile (the clip maximum, then a RELEASE increment of the clip's counter; it covers the whole CTA's
  // log-mel rows of the tile): owed by warp (tile mod 8), so that the ~1 us the release fence blocks a warp is spread evenly --
  // "the last warp to finish" would make the slowest warp slower still, and the other seven wait for it at the next tile.
  // The owner sends it from a quiet point of a later FFT, as soon as the tile's `empty` barrier is seen complete (all 8 warps have
  // stored their rows and added their maximum to mel_amax; at the latest three tiles on, when this warp has waited for that
  // barrier anyway), or waits for the barrier itself when it must.
  int owe_n = -1, owe_clip = 0;        // tile (position in the CTA's power ring sequence) and its clip; uniform over the warp
  auto send_signal = [&](bool wait) {
    if (owe_n >= 0) {
      const int s = owe_n % P3_RING;
      if (wait) mbar_wait3(&sm.empty[s], (owe_n / P3_RING) & 1, 2);
      if (lane == 0) {
        const float vmax = lg2_fast(fmaxf(__uint_as_float(atomicExch(&sm.mel_amax[s], 0u)), 1e-10f));   // log2 units, >= log2(1e-10)
        atomicMax(clip_max + owe_clip, float_to_ordered(vmax));
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(clip_done + owe_clip) : "memory");
      }
      owe_n = -1;
    }
  };

  float v[25];                         // raw samples of the round about to be transformed, loaded one round ahead
  bool v_ready = false;
  int n_fft = 0;                       // frame items this CTA has started (power-ring position)
  bool pend = false;                   // a frame item whose mel phase is still to run
  int pend_seq = 0;
  uint32_t poll_par = 0;               // phase of this warp's poll barrier
  bool more = true;                    // false: the drain pass after the last item (runs the pending mel phase)

  for (int seq = 0; more || pend; ++seq) {
    const int dslot = seq % D3_RING;
    if (more) {
      if (tid == 0) publish(seq + 2);
      mbar_wait3(&sm.ring_full[dslot], (seq / D3_RING) & 1, 3);
      more = sm.ring[dslot].n_frames >= 0;
    }
    const Item3& cur = sm.ring[dslot];
    const Item3& nxt = sm.ring[(seq + 1) % D3_RING];
    const int cur_frames = more ? cur.n_frames : 0;
    const int cur_clamp = more ? cur.c_n_frames : 0;
    const unsigned int need = cur_clamp > 0 ? static_cast<unsigned int>(cur.c_need) : 0u;
    int slot = 0, use = 0;
    bool polled = false;

    if (cur_frames > 0) {
      // ---------------- stages 1 + 2, warp-synchronous: this warp's frames 4 w .. 4 w + 3 of the item, two at a time ----------------
      slot = n_fft % P3_RING;
      use = n_fft / P3_RING;
      if (use > 0) mbar_wait3(&sm.empty[slot], (use - 1) & 1, 4);
      float2* Ew = &sm.E[warp][0][0][0];
      // round -1 only loads (the first item of the CTA, or after items without frames)
#pragma unroll 1
      for (int r = v_ready ? 0 : -1; r < 2; ++r) {
        if (r >= 0) {
          float u[28];
          const float4* w4 = reinterpret_cast<const float4*>(&sm.win[j][0]);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            const float4 w = w4[c];
            u[4 * c] = w.x; u[4 * c + 1] = w.y; u[4 * c + 2] = w.z; u[4 * c + 3] = w.w;
          }
          float x[25];
#pragma unroll
          for (int m = 0; m < 25; ++m) x[m] = v[m] * u[m];
          const float4* t4 = reinterpret_cast<const float4*>(&sm.tw[j][0]);   // [q = 2 c, 2 c + 1]; q = 0 is 1 and not applied
          float4 t[7];                                                         // read before the transform that hides their latency
#pragma unroll
          for (int c = 0; c < 7; ++c) t[c] = t4[c];
          float2 V[K1];
          rdft25(x, V);
          float2* e = Ew + fr * K1 * E_PITCH + j;
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            if (c == 0) e[0] = V[0]; else e[2 * c * E_PITCH] = cmul(V[2 * c], make_float2(t[c].x, t[c].y));
            if (2 * c + 1 < K1) e[(2 * c + 1) * E_PITCH] = cmul(V[2 * c + 1], make_float2(t[c].z, t[c].w));
          }
        }
        bool nxt_frames = false;
        if (r >= 0) {
          // a quiet point of the item: no load or store of this warp is in flight, so the release of an owed signal waits for
          // nothing but its own round trip; the counter of the clamp tile is fetched into shared memory by a 16-byte bulk copy
          // (read at L2) behind the warp's own mbarrier, i.e. outside the scoreboards the sample loads below occupy -- an ordinary
          // load, or a cp.async group, would be waited for together with them -- and read two stage 2's later
          if (owe_n >= 0 && (n_fft >= owe_n + P3_RING || mbar_test(&sm.empty[owe_n % P3_RING], (owe_n / P3_RING) & 1))) send_signal(false);
          if (r == 0 && cur_clamp > 0) {
            if (lane == 0) {
              const unsigned int* src = clip_done + cur.c_clip;
              mbar_expect_tx(&sm.poll_bar[warp], 16);
              bulk_g2s(&sm.poll[warp], reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(src) & ~static_cast<uintptr_t>(15)), 16,
                       &sm.poll_bar[warp]);
            }
            polled = true;
          }
        }
        if (r == 1) {
          mbar_wait3(&sm.ring_full[(seq + 1) % D3_RING], ((seq + 1) / D3_RING) & 1, 5);   // published an item ago
          nxt_frames = nxt.n_frames > 0;
        }
        __syncwarp();
        {
          // the next round's samples travel while stage 2 (and, across items, the mel phase) runs: 25 strided loads per lane,
          // clip edges by index mirroring
          const Item3& it = r < 1 ? cur : nxt;
          const int rn = (r + 1) & 1;
          const int f = warp * 4 + 2 * rn + fr;
          const int nf = (r < 1 || nxt_frames) ? it.n_frames : 0;
          if (f < nf) {
            const int n_samples = it.n_samples;
            const int s0 = (it.frame0 + f) * HOP - N_FFT / 2;          // clip-relative index of the frame's first sample
            const float* clip = pcm + it.pcm_off;
            if (s0 >= 0 && s0 + N_FFT <= n_samples) {                  // interior frame (uniform over the 16 lanes of a frame)
              const float* x = clip + s0 + j;
#pragma unroll
              for (int m = 0; m < 25; ++m) v[m] = __ldg(x + 16 * m);
            } else {
#pragma unroll
              for (int m = 0; m < 25; ++m) v[m] = __ldg(clip + reflect_index(s0 + j + 16 * m, n_samples));
            }
          } else {
#pragma unroll
            for (int m = 0; m < 25; ++m) v[m] = 0.f;
          }
          if (r == 1) v_ready = nxt_frames;
        }
        if (r >= 0 && fr2 < 2) {
          const int f = warp * 4 + 2 * r + fr2;
          const float4* e = reinterpret_cast<const float4*>(Ew + (fr2 * K1 + k1) * E_PITCH);
          float2 z[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 q = e[c];
            z[2 * c] = make_float2(q.x, q.y);
            z[2 * c + 1] = make_float2(q.z, q.w);
          }
          fft16(z);
          float* prow = &sm.P[slot][f][0];
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2) {
            const float pw = fmaf(z[k2].x, z[k2].x, z[k2].y * z[k2].y);
            if (stage2_unique(k1, k2)) prow[stage2_bin(k1, k2)] = pw;
          }
        }
        __syncwarp();   // E is rewritten by the next round; the P rows are complete before lane 0 arrives
      }
      if (lane == 0) mbar_arrive(&sm.full[slot]);
    } else {
      v_ready = false;
    }

    // ---------------- mel phase of the PREVIOUS frame item, and the clamp tile this item carries ----------------
    // clamp + rescale rows 16 w .. 16 w + 15 of a tile whose clip finished `lag` items ago (still in L2).  The clip's counter is
    // read relaxed and the tile and the maximum are then read from L2 (.cg): those loads are issued only after the counter's
    // value has been tested, L2 is the point of coherence, and the L1 (which holds the PCM the FFT rounds re-read) is not
    // invalidated as an acquire would.  In the common case the tile's 16 loads are issued here and travel during the mel phase.
    float c[16], floor_raw = 0.f;
    float* ctile = nullptr;
    const long long row_stride = ld;
    if (cur_clamp > 0) ctile = out + cur.c_col0 + cur.c_frame0 + lane + static_cast<long long>(16 * warp) * ld;
    const bool clamp_lane = lane < cur_clamp;
    bool ready = false, clamp_loaded = false;
    if (polled) {
      mbar_wait3(&sm.poll_bar[warp], poll_par, 6);
      poll_par ^= 1;
      const unsigned int* w = reinterpret_cast<const unsigned int*>(&sm.poll[warp]);
      ready = w[(reinterpret_cast<uintptr_t>(clip_done + cur.c_clip) >> 2) & 3] >= need;
    }
    bool promoted = false;
    // (further passes only at the tail of a launch: the tile has to wait, and what it waits for may be this warp's own pending item)
#pragma unroll 1
    for (;;) {
      if (ready && !clamp_loaded) {
        floor_raw = __uint_as_float(__ldcg(clip_max + cur.c_clip));
        if (clamp_lane) {
          const float* p = ctile;
#pragma unroll
          for (int i = 0; i < 16; ++i, p += row_stride) c[i] = __ldcg(p);
        }
        clamp_loaded = true;
      }
      if (pend) {
        const Item3& pit = sm.ring[pend_seq % D3_RING];            // (the previous item of the sequence; this one on a second pass)
        const int pend_slot = (n_fft - 1) % P3_RING;
        mbar_wait3(&sm.full[pend_slot], ((n_fft - 1) / P3_RING) & 1, 7);
        const bool live = lane < pit.n_frames;
        const float* prow = &sm.P[pend_slot][lane][0];
        // lanes beyond the clip's last frame store to a dump word (stride 0) instead of branching around 16 stores
        float* ocol = live ? out + pit.col0 + pit.frame0 + lane : dump;
        const int ldl = live ? ld : 0;
        float amax = 0.f;
        switch (warp) {
          case 0: mel_group3<0>(fw, prow, ocol, ldl, amax); break;
          case 1: mel_group3<1>(fw, prow, ocol, ldl, amax); break;
          case 2: mel_group3<2>(fw, prow, ocol, ldl, amax); break;
          case 3: mel_group3<3>(fw, prow, ocol, ldl, amax); break;
          case 4: mel_group3<4>(fw, prow, ocol, ldl, amax); break;
          case 5: mel_group3<5>(fw, prow, ocol, ldl, amax); break;
          case 6: mel_group3<6>(fw, prow, ocol, ldl, amax); break;
          default: mel_group3<7>(fw, prow, ocol, ldl, amax); break;
        }
        amax = warp_max(live ? amax : 0.f);                         // >= 1e-10 for every live lane
        __syncwarp();                                               // every lane's reads of P and stores to `out` are done
        if (lane == 0) {
          atomicMax(&sm.mel_amax[pend_slot], __float_as_uint(amax));   // non-negative floats order like their bit patterns
          mbar_arrive(&sm.empty[pend_slot]);                           // (release: the maximum and, cumulatively, the warp's rows)
        }
        if (((n_fft - 1) & (kWarpsPerCta - 1)) == warp) {
          send_signal(false);                                       // (never owed here: it left at an FFT five tiles ago; complete anyway)
          owe_n = n_fft - 1;
          owe_clip = pit.clip;
        }
        pend = false;
      }
      if (!promoted && cur_frames > 0) {
        pend = true;
        pend_seq = seq;
        ++n_fft;
      }
      promoted = true;
      if (cur_clamp == 0 || clamp_loaded) break;
      if (ld_relaxed_u32(clip_done + cur.c_clip) < need) {
        if (pend) continue;   // run the pending mel phase first
        send_signal(true);
        for (unsigned int spins = 0; ld_relaxed_u32(clip_done + cur.c_clip) < need; ++spins) {
          __nanosleep(2000);
          if (spins > (1u << 22)) mel_wait_timeout(8);   // seconds
        }
      }
      ready = true;
    }
    if (cur_clamp > 0 && clamp_lane) {
      // The raw values are log2 of the mel power.  With c = log10(2): log10 = c L, the reference's max(log10, log10max - 8) and its
      // 1e-10 clamp are max(L, F) with F = max(c Lmax - 8, -10) / c, and (x + 4) / 4 = (max(L, F) - F) * (c / 4) + (floor10 + 4) / 4.
      // Written relative to the floor so that a clamped value is EXACTLY the reference's (a silent clip: -1.5).
      const float vmax10 = ordered_to_float(__float_as_uint(floor_raw)) * 0.30102999566398120f;
      const float floor10 = fmaxf(vmax10 - 8.0f, -10.0f);
      const float floor_v = floor10 * 3.3219280948873623f;
      const float floor_out = fmaf(floor10, 0.25f, 1.0f);
      float* p = ctile;
#pragma unroll
      for (int i = 0; i < 16; ++i, p += row_stride) __stcg(p, fmaf(fmaxf(c[i], floor_v) - floor_v, 0.07525749891599530f, floor_out));
    }
    __syncwarp();   // every lane has read the previous item's descriptor (its mel phase ran in this pass)
    if (seq > 0 && lane == 0) mbar_arrive(&sm.ring_empty[(seq - 1) % D3_RING]);
  }
  send_signal(true);
}

}  // namespace

cudaError_t launch_logmel(const float* pcm, const mel::Item* items, int n_items, const mel::Tables* tables, float* mel_out,
                          long long mel_ld, unsigned int* counters, int n_clips, int num_sms, cudaStream_t stream) {
  if (n_items == 0) return cudaSuccess;
  // counters: [0] ticket, [1, 1 + n_clips) done, [1 + n_clips, 1 + 2 n_clips) max (ordered-uint encoding; 0 = below every float)
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(unsigned int) * (1 + 2 * static_cast<size_t>(n_clips)), stream);
  if (e != cudaSuccess) return e;
  const int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(MelSmem)));
  if (e != cudaSuccess) return e;
  logmel_kernel<<<grid, mel::THREADS, sizeof(MelSmem), stream>>>(pcm, items, n_items, tables, mel_out, mel_ld, counters, counters + 1,
                                                                 counters + 1 + n_clips);
  return cudaGetLastError();
}

cudaError_t launch_logmel_v3(const float* pcm, const mel::Item3* items, int n_items, const mel::Tables* tables, const mel::Tables* host_tables,
                             float* mel_out, long long mel_ld, unsigned int* counters, int n_clips, int num_sms, cudaStream_t stream) {
  if (n_items == 0) return cudaSuccess;
  if (mel_ld >= (1LL << 24)) return cudaErrorInvalidValue;   // 127 * ld is formed in 32 bits (the caller falls back to v1 above this)
  // counters: [0] ticket, [1, 1 + n_clips) done (one count per frame item), [1 + n_clips, 1 + 2 n_clips) max (ordered-uint
  // encoding; 0 = below every float), [1 + 2 n_clips] dump word for the lanes beyond a clip's end; kMelCounterPad words of slack
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(unsigned int) * mel::counter_words(n_clips), stream);
  if (e != cudaSuccess) return e;
  const int cap = mel::v3_grid(num_sms);
  const int grid = n_items < cap ? n_items : cap;
  e = cudaFuncSetAttribute(logmel_kernel_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(MelSmem3)));
  if (e != cudaSuccess) return e;
  MelWeights fw;
  std::memcpy(fw.w, host_tables->fw, sizeof(fw.w));
  logmel_kernel_v3<<<grid, mel::THREADS, sizeof(MelSmem3), stream>>>(pcm, items, n_items, tables, mel_out, static_cast<int>(mel_ld), counters, counters + 1,
                                                                     counters + 1 + n_clips, reinterpret_cast<float*>(counters + 1 + 2 * n_clips), fw);
  return cudaGetLastError();
}

}  // namespace qasr

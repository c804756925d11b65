// qasr_pool_*: one process, several GPUs (SURVEY.md section 8(b) "multi-GPU: qasr_pool_create / submit / collect", 8(e)).
//
// The path shards with no exchange step -- every clip is encoded independently -- so the pool is host code only: one
// handle (weights replicated), one worker thread and one CUDA stream per device.  A submitted batch is cut into at most
// n_devices CONTIGUOUS clip ranges of near-equal work (mel frames), so that every shard reads a contiguous slice of the
// caller's pinned PCM buffer and writes a contiguous slice of the caller's output buffer in clip order: no gather, no
// extra host copy, no collective, no NCCL.  Each worker drives its handle's pipelined entry points
// (qasr_submit_pcm_host / qasr_wait), keeping up to two shards in flight so that the copies of consecutive batches
// overlap the compute, exactly as a single-GPU caller would.
#include "../../include/qasr_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct Batch {  // one qasr_pool_submit
  int pending = 0;       // shards not finished yet
  int rc = 0;            // first failure
  std::string error;
};

struct Shard {
  std::shared_ptr<Batch> batch;
  const float* pcm_host = nullptr;
  std::vector<int64_t> offsets;   // contiguous shard: its clip offsets (absolute sample indices into pcm_host)
  // scattered shard (LPT): per clip its first sample, length and first output row (absolute, into the caller's buffers)
  std::vector<int64_t> begin, len, out_row;
  void* out_host = nullptr;
  int64_t out_capacity_tokens = 0;
  uint64_t handle_ticket = 0;
  double t_enqueue0 = 0.0;        // host clock when the worker started enqueueing it
};

struct WorkerStats {
  double shards = 0, enqueue_ms = 0, wait_ms = 0, h2d_ms = 0, compute_ms = 0, d2h_ms = 0, turnaround_ms = 0;
};

struct Worker {
  int device = 0;
  qasr_handle_t handle = nullptr;
  cudaStream_t stream = nullptr;
  std::thread thread;
  std::deque<Shard> queue;      // submitted, not yet enqueued on the GPU
  std::deque<Shard> inflight;   // enqueued (at most two: the handle double-buffers)
  WorkerStats stats;            // guarded by the pool mutex
};

}  // namespace

struct qasr_pool_s {
  std::vector<std::unique_ptr<Worker>> workers;
  std::mutex mu;
  std::condition_variable cv_work;   // workers: new shard or shutdown
  std::condition_variable cv_done;   // collectors: a batch finished
  bool stop = false;
  bool finalized = false;
  uint64_t next_ticket = 1;
  std::map<uint64_t, std::shared_ptr<Batch>> batches;
  int output_dim = 0;
  int sharding = QASR_SHARD_AUTO;
};

namespace {

// Contiguous partition of the clips into g ranges of near-equal work (mel frames): cut after the clip at which the running total
// crosses k / g of the work (k = 1 .. g - 1), rounding to the nearer side of the boundary clip.  With many clips per GPU this is
// within one clip of the optimum; with fewer clips than GPUs some ranges are empty.  cut has g + 1 entries, cut[0] = 0, cut[g] = n.
void partition_clips(const std::vector<int64_t>& frames, int g, std::vector<int>* cut) {
  const int n = static_cast<int>(frames.size());
  int64_t total = 0;
  for (int64_t f : frames) total += f;
  cut->assign(g + 1, n);
  (*cut)[0] = 0;
  int i = 0;
  int64_t run = 0;
  for (int k = 1; k < g; ++k) {
    const double target = static_cast<double>(total) * k / g;
    while (i < n && static_cast<double>(run) + 0.5 * static_cast<double>(frames[i]) <= target) run += frames[i++];
    (*cut)[k] = std::max(i, (*cut)[k - 1]);
  }
}

// Longest-processing-time-first: clips by mel frames descending (ties: lower index first), each to the
// least-loaded member (ties: lower member first).  Deterministic; shard[i] = member of clip i.
void lpt_assign(const std::vector<int64_t>& frames, int g, std::vector<int>* shard) {
  const int n = static_cast<int>(frames.size());
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return frames[a] > frames[b]; });
  std::vector<int64_t> load(g, 0);
  shard->assign(n, 0);
  for (int i : order) {
    int best = 0;
    for (int k = 1; k < g; ++k)
      if (load[k] < load[best]) best = k;
    (*shard)[i] = best;
    load[best] += frames[i];
  }
}

// QASR_SHARD_AUTO: contiguous ranges (one copy each way per shard) unless the batch has fewer than four clips per member, where a
// range boundary can cost a whole clip of imbalance and LPT's per-clip copies are few
bool use_lpt(int mode, int n_clips, int g) { return mode == QASR_SHARD_LPT || (mode == QASR_SHARD_AUTO && g > 1 && n_clips < 4 * g); }

void finish_shard(qasr_pool_s* p, Shard& s, int rc) {
  std::string err;
  if (rc != 0) {
    const char* e = qasr_last_error();  // thread-local: read it on the worker that failed
    err = e != nullptr ? e : "?";
  }
  std::lock_guard<std::mutex> lk(p->mu);
  if (rc != 0 && s.batch->rc == 0) {
    s.batch->rc = rc;
    s.batch->error = err;
  }
  if (--s.batch->pending == 0) p->cv_done.notify_all();
}

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void worker_main(qasr_pool_s* p, Worker* w) {
  cudaSetDevice(w->device);
  for (;;) {
    Shard s;
    bool enqueue = false;
    {
      std::unique_lock<std::mutex> lk(p->mu);
      p->cv_work.wait(lk, [&] { return p->stop || !w->queue.empty() || !w->inflight.empty(); });
      if (w->queue.empty() && w->inflight.size() == 1 && !p->stop) {
        // One shard in flight, nothing queued: do NOT park in qasr_wait -- the caller's next batch typically arrives a moment after
        // it has collected the previous one, and a worker blocked for the ~12 ms of the running shard would enqueue it (host
        // launches + H2D, ~1.7 ms) only after the GPU had gone idle.  Wait for new work, looking at the shard now and then.
        p->cv_work.wait_for(lk, std::chrono::microseconds(100), [&] { return p->stop || !w->queue.empty(); });
        if (w->queue.empty() && !p->stop) {
          int done = 0;
          const uint64_t t = w->inflight.front().handle_ticket;
          lk.unlock();
          const int rc = qasr_poll(w->handle, t, &done);
          lk.lock();
          if (rc == 0 && done == 0) continue;
        }
      }
      if (!w->queue.empty() && w->inflight.size() < 2) {
        s = std::move(w->queue.front());
        w->queue.pop_front();
        enqueue = true;
      } else if (!w->inflight.empty()) {
        s = std::move(w->inflight.front());
        w->inflight.pop_front();
      } else if (p->stop) {
        return;
      } else {
        continue;
      }
    }
    if (enqueue) {
      s.t_enqueue0 = now_ms();
      // token lengths were already reported by qasr_pool_submit itself: nothing the caller owns besides the PCM and output
      // buffers is touched from this thread
      const int rc = s.begin.empty()
                         ? qasr_submit_pcm_host(w->handle, s.pcm_host, s.offsets.data(), static_cast<int>(s.offsets.size()) - 1, s.out_host,
                                                s.out_capacity_tokens, nullptr, w->stream, &s.handle_ticket)
                         : qasr_submit_clips_host(w->handle, s.pcm_host, s.begin.data(), s.len.data(), s.out_row.data(),
                                                  static_cast<int>(s.begin.size()), s.out_host, s.out_capacity_tokens, nullptr, w->stream,
                                                  &s.handle_ticket);
      if (rc != 0) {
        finish_shard(p, s, rc);
      } else {
        std::lock_guard<std::mutex> lk(p->mu);
        w->stats.enqueue_ms += now_ms() - s.t_enqueue0;
        w->inflight.push_back(std::move(s));
      }
    } else {
      const double t0 = now_ms();
      const int rc = qasr_wait(w->handle, s.handle_ticket);
      const double t1 = now_ms();
      float a = 0.f, b = 0.f, c = 0.f;
      const bool timed = rc == 0 && qasr_pipe_times(w->handle, s.handle_ticket, &a, &b, &c) == 0;
      {
        std::lock_guard<std::mutex> lk(p->mu);
        w->stats.shards += 1;
        w->stats.wait_ms += t1 - t0;
        w->stats.turnaround_ms += t1 - s.t_enqueue0;
        if (timed) { w->stats.h2d_ms += a; w->stats.compute_ms += b; w->stats.d2h_ms += c; }
      }
      finish_shard(p, s, rc);
    }
  }
}

}  // namespace

extern "C" {

int qasr_pool_create(const qasr_config_t* cfg, const int* devices, int n_devices, qasr_pool_t* out) {
  QASR_REQUIRE(cfg != nullptr && devices != nullptr && out != nullptr && n_devices >= 1 && n_devices <= 64, "qasr_pool_create: bad argument");
  std::unique_ptr<qasr_pool_s> p(new qasr_pool_s());
  p->output_dim = cfg->output_dim;
  for (int i = 0; i < n_devices; ++i) {
    std::unique_ptr<Worker> w(new Worker());
    w->device = devices[i];
    const int rc = qasr_create(cfg, devices[i], &w->handle);
    if (rc != 0) {
      for (auto& o : p->workers) qasr_destroy(o->handle);
      return rc;
    }
    p->workers.push_back(std::move(w));
  }
  *out = p.release();
  return 0;
}

int qasr_pool_size(qasr_pool_t p) { return p == nullptr ? 0 : static_cast<int>(p->workers.size()); }

int qasr_pool_set_weight(qasr_pool_t p, const char* name, const void* data, int dtype, const int64_t* shape, int ndim) {
  QASR_REQUIRE(p != nullptr && !p->finalized, "qasr_pool_set_weight: bad pool or already finalized");
  for (auto& w : p->workers) {
    const int rc = qasr_set_weight(w->handle, name, data, dtype, shape, ndim);
    if (rc != 0) return rc;
  }
  return 0;
}

int qasr_pool_finalize(qasr_pool_t p) {
  QASR_REQUIRE(p != nullptr && !p->finalized, "qasr_pool_finalize: bad pool or already finalized");
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& w : p->workers) {
    const int rc = qasr_finalize(w->handle);
    if (rc != 0) return rc;
    QASR_CUDA_CHECK(cudaSetDevice(w->device));
    QASR_CUDA_CHECK(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
  }
  cudaSetDevice(prev);
  for (auto& w : p->workers) w->thread = std::thread(worker_main, p, w.get());
  p->finalized = true;
  return 0;
}

size_t qasr_pool_workspace_bytes(qasr_pool_t p) {
  size_t total = 0;
  if (p != nullptr)
    for (auto& w : p->workers) total += qasr_workspace_bytes(w->handle);
  return total;
}

int qasr_pool_submit(qasr_pool_t p, const float* pcm_host, const int64_t* clip_offsets, int n_clips, void* out_host,
                     int64_t out_capacity_tokens, int64_t* token_lens_out, int32_t* clip_device_out, uint64_t* ticket_out) {
  QASR_REQUIRE(p != nullptr && p->finalized, "qasr_pool_submit before qasr_pool_finalize");
  QASR_REQUIRE(clip_offsets != nullptr && n_clips >= 0 && ticket_out != nullptr && token_lens_out != nullptr, "qasr_pool_submit: bad argument");
  *ticket_out = 0;
  if (n_clips == 0) return 0;
  QASR_REQUIRE(pcm_host != nullptr && out_host != nullptr, "qasr_pool_submit: null buffer");
  // work per clip = mel frames; token counts give every shard its slice of the output (clip order)
  std::vector<int64_t> frames(n_clips), tok_off(n_clips + 1, 0);
  for (int i = 0; i < n_clips; ++i) {
    const int64_t n = clip_offsets[i + 1] - clip_offsets[i];
    QASR_REQUIRE(n >= 0, "qasr_pool_submit: offsets must be non-decreasing");
    frames[i] = n / 160;
    token_lens_out[i] = qasr_token_len(frames[i]);
    tok_off[i + 1] = tok_off[i] + token_lens_out[i];
  }
  QASR_REQUIRE(tok_off[n_clips] <= out_capacity_tokens, "qasr_pool_submit: output buffer too small for " + std::to_string(tok_off[n_clips]) + " tokens");
  const int g = static_cast<int>(p->workers.size());
  auto batch = std::make_shared<Batch>();
  std::vector<std::pair<int, Shard>> shards;
  if (use_lpt(p->sharding, n_clips, g)) {
    std::vector<int> member;
    lpt_assign(frames, g, &member);
    std::vector<Shard> per(g);
    for (int i = 0; i < n_clips; ++i) {
      Shard& s = per[member[i]];
      s.begin.push_back(clip_offsets[i]);
      s.len.push_back(clip_offsets[i + 1] - clip_offsets[i]);
      s.out_row.push_back(tok_off[i]);
      if (clip_device_out != nullptr) clip_device_out[i] = p->workers[member[i]]->device;
    }
    for (int k = 0; k < g; ++k) {
      if (per[k].begin.empty()) continue;
      per[k].batch = batch;
      per[k].pcm_host = pcm_host;
      per[k].out_host = out_host;
      per[k].out_capacity_tokens = out_capacity_tokens;
      shards.emplace_back(k, std::move(per[k]));
    }
  }
  std::vector<int> cut;
  if (shards.empty()) partition_clips(frames, g, &cut);
  for (int k = 0; k < g && !cut.empty(); ++k) {
    const int a = cut[k], b = cut[k + 1];
    if (clip_device_out != nullptr)
      for (int i = a; i < b; ++i) clip_device_out[i] = p->workers[k]->device;
    if (b <= a) continue;
    Shard s;
    s.batch = batch;
    s.pcm_host = pcm_host;
    s.offsets.assign(clip_offsets + a, clip_offsets + b + 1);
    s.out_host = static_cast<char*>(out_host) + static_cast<size_t>(tok_off[a]) * p->output_dim * 2;
    s.out_capacity_tokens = tok_off[b] - tok_off[a];
    shards.emplace_back(k, std::move(s));
  }
  batch->pending = static_cast<int>(shards.size());
  {
    std::lock_guard<std::mutex> lk(p->mu);
    const uint64_t t = p->next_ticket++;
    p->batches[t] = batch;
    for (auto& ks : shards) p->workers[ks.first]->queue.push_back(std::move(ks.second));
    *ticket_out = t;
  }
  p->cv_work.notify_all();
  return 0;
}

int qasr_pool_set_sharding(qasr_pool_t p, int mode) {
  QASR_REQUIRE(p != nullptr, "qasr_pool_set_sharding: null pool");
  QASR_REQUIRE(mode == QASR_SHARD_AUTO || mode == QASR_SHARD_CONTIGUOUS || mode == QASR_SHARD_LPT, "qasr_pool_set_sharding: unknown mode");
  std::lock_guard<std::mutex> lk(p->mu);
  p->sharding = mode;
  return 0;
}

int qasr_pool_plan_mode(const int64_t* clip_offsets, int n_clips, int n_devices, int mode, int32_t* clip_shard_out) {
  QASR_REQUIRE(clip_offsets != nullptr && clip_shard_out != nullptr && n_clips >= 0 && n_devices >= 1, "qasr_pool_plan: bad argument");
  QASR_REQUIRE(mode == QASR_SHARD_AUTO || mode == QASR_SHARD_CONTIGUOUS || mode == QASR_SHARD_LPT, "qasr_pool_plan: unknown mode");
  std::vector<int64_t> frames(n_clips);
  for (int i = 0; i < n_clips; ++i) {
    QASR_REQUIRE(clip_offsets[i + 1] >= clip_offsets[i], "qasr_pool_plan: offsets must be non-decreasing");
    frames[i] = (clip_offsets[i + 1] - clip_offsets[i]) / 160;
  }
  if (use_lpt(mode, n_clips, n_devices)) {
    std::vector<int> member;
    lpt_assign(frames, n_devices, &member);
    for (int i = 0; i < n_clips; ++i) clip_shard_out[i] = member[i];
    return 0;
  }
  std::vector<int> cut;
  partition_clips(frames, n_devices, &cut);
  for (int k = 0; k < n_devices; ++k)
    for (int i = cut[k]; i < cut[k + 1]; ++i) clip_shard_out[i] = k;
  return 0;
}

int qasr_pool_plan(const int64_t* clip_offsets, int n_clips, int n_devices, int32_t* clip_shard_out) {
  return qasr_pool_plan_mode(clip_offsets, n_clips, n_devices, QASR_SHARD_CONTIGUOUS, clip_shard_out);
}

int qasr_pool_collect(qasr_pool_t p, uint64_t ticket) {
  QASR_REQUIRE(p != nullptr, "qasr_pool_collect: null pool");
  if (ticket == 0) return 0;
  std::shared_ptr<Batch> b;
  {
    std::unique_lock<std::mutex> lk(p->mu);
    auto it = p->batches.find(ticket);
    QASR_REQUIRE(it != p->batches.end(), "qasr_pool_collect: unknown (or already collected) ticket");
    b = it->second;
    p->cv_done.wait(lk, [&] { return b->pending == 0; });
    p->batches.erase(it);
  }
  if (b->rc != 0) {
    qasr::set_last_error("qasr_pool_collect: a shard failed: " + b->error);
    return b->rc;
  }
  return 0;
}

int qasr_pool_stats(qasr_pool_t p, double* out, int capacity_members, int reset) {
  QASR_REQUIRE(p != nullptr && out != nullptr && capacity_members >= static_cast<int>(p->workers.size()), "qasr_pool_stats: bad argument");
  std::lock_guard<std::mutex> lk(p->mu);
  for (size_t i = 0; i < p->workers.size(); ++i) {
    WorkerStats& st = p->workers[i]->stats;
    const double n = st.shards > 0 ? st.shards : 1.0;
    double* o = out + i * QASR_POOL_STAT_FIELDS;
    o[0] = st.shards; o[1] = st.enqueue_ms / n; o[2] = st.wait_ms / n; o[3] = st.h2d_ms / n; o[4] = st.compute_ms / n; o[5] = st.d2h_ms / n;
    o[6] = st.turnaround_ms / n;
    if (reset != 0) st = WorkerStats();
  }
  return 0;
}

void qasr_pool_destroy(qasr_pool_t p) {
  if (p == nullptr) return;
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->stop = true;
  }
  p->cv_work.notify_all();
  for (auto& w : p->workers)
    if (w->thread.joinable()) w->thread.join();   // workers drain what is queued / in flight before they leave
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& w : p->workers) {
    if (w->stream != nullptr) {
      cudaSetDevice(w->device);
      cudaStreamSynchronize(w->stream);
      cudaStreamDestroy(w->stream);
    }
    qasr_destroy(w->handle);
  }
  cudaSetDevice(prev);
  delete p;
}

}  // extern "C"

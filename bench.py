#!/usr/bin/env python
"""Benchmark of the hot path: PCM -> log-mel -> Qwen3-ASR audio-encoder hidden states.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c4|c5] [--quantize fp8|fp8_per_row]

One step = one pass of the hot path over one batch of synthetic input.  The default workload is BASELINE.json configs[1] (c2:
Qwen3-ASR-1.7B dims, 32 x 30 s 16 kHz clips per GPU, weak scaling); the other BASELINE.json configs are selectable and print the
same JSON schema:
    c1  configs[0]  0.6B, one 5 s clip, batch 1 (the latency-bound end; replicas at N > 1)
    c2  configs[1]  1.7B, 32 x 30 s per GPU                                                       [default; the headline]
    c3  configs[2]  1.7B, 128 WebSocket windows (0.45-6 s) as int16 PCM: /32768, band-pass, flush pad, log-mel, encoder, one ragged batch
    c4  configs[3]  1.7B, one hour of audio in ~230 segments of 1-30 s, LPT-sharded by clip over the N GPUs (strong scaling)
    c5  configs[4]  (ii) the c3 windows with BOTH models resident: partials -> 0.6B, flushes -> 1.7B (src/server.py:1351);
                    (i) is `--config c2 --quantize fp8`
Clips are independent: no data-path collective at any N; NCCL is used only for the timing barrier / max-over-ranks.  Prints ONE JSON
line on rank 0.  `--impl reference` times the reference's CPU torch path (the oracle restatement of the transformers classes the
reference calls) on the host cores instead, on a bounded sample of the same workload.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# DRAM traffic measured once with ncu --set full (profiles/): the bench itself never runs under a profiler
# (read + write GB of conv2, conv3, conv_out, then per layer qkv, out_proj, fc1, fc2: profiles/r02p_gemm_raw_summary.txt, final build)
NCU_GEMM_DRAM_BYTES_PER_STEP = (3.752991 + 0.727449 + 0.779130 + 0.177139 + 0.241481 + 0.020581 + 24 * (
    0.032237 + 0.028606 + 0.053577 + 0.000636 + 0.034357 + 0.051031 + 0.138782 + 0.013685)) * 1e9
NCU_MEL_DRAM_BYTES_PER_LAUNCH = 516.7e6 + 408.5e6   # dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02o_mel_v3_final_summary.txt

METRIC = "audio-sec encoded/sec (mel+encoder, 1.7B)"
UNIT = "audio-s/s"
SR = 16000
N_CLIPS = 32
CLIP_SECONDS = 30.0
DEFAULT_STEPS = {"c1": 200, "c2": 20, "c3": 50, "c4": 5, "c5": 50}
REF_STEPS = {"c1": 5, "c2": 2, "c3": 2, "c4": 2, "c5": 2}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops_sustained"], "bf16_tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples = []
        self.windows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def median_sm_mhz(self, t0, t1):
        v = []
        for ts, line in list(self.samples):
            if t0 <= ts <= t1 + 0.05:
                try:
                    v.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return float(np.median(v)) if v else None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.samples:
            if not any(a - 0.05 <= ts <= b + 0.05 for a, b in self.windows):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ the CPU arm (oracle port)
def oracle_step(weights, cfg, clips, threads):
    """The reference's CPU path (oracle restatement): torch.stft log-mel + fp32 encoder with the window mask."""
    from oracle import encoder_forward, logmel_torch_f32
    from oracle.encoder import EncoderConfig

    ecfg = EncoderConfig(cfg["d_model"], cfg["encoder_layers"], cfg["encoder_attention_heads"], cfg["encoder_ffn_dim"],
                         cfg["output_dim"], name=cfg["name"])
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    mels = [logmel_torch_f32(c) for c in clips]
    mels = [torch.from_numpy(m).to(torch.bfloat16).float().numpy() for m in mels]
    out, toks = encoder_forward(weights, ecfg, mels)
    dt = time.perf_counter() - t0
    return dt, out, toks


def ws_prefilter_cpu(win16: np.ndarray, flush: bool) -> np.ndarray:
    """The reference's own CPU prologue of a WS window (src/server.py:1321-1338 + the SDK's 0.5 s minimum)."""
    from scipy.signal import butter, sosfilt

    x = np.concatenate([win16, np.zeros(9600, np.int16)]) if flush else win16
    f = x.astype(np.float32) / 32768.0
    y = sosfilt(butter(4, [300, 3400], btype="bandpass", fs=SR, output="sos"), f).astype(np.float32)
    return np.concatenate([y, np.zeros(8000 - len(y), np.float32)]) if len(y) < 8000 else y


def cpu_sample(config: str, bounded: bool):
    """[(model name, clips, description)] -- the bounded sample of the workload the CPU arm is timed on (about 10-30 s of CPU work)."""
    from qwen3_asr_b200 import synth

    if config == "c1":
        return [("0.6B", synth.workload_c1(), "the C1 clip itself (5 s, batch 1)")]
    if config == "c2":
        n = 4 if bounded else 2   # one batched call over the sample, as the SDK batches the chunks of a request
        return [("1.7B", [synth.speech_like(int(CLIP_SECONDS * SR), i) for i in range(n)], f"first {n} of the {N_CLIPS} clips of the C2 batch (one batched call)")]
    if config in ("c3", "c5"):
        wins, flush = synth.workload_c3()
        idx = list(range(8))
        clips = [ws_prefilter_cpu(wins[i], flush[i]) for i in idx]
        if config == "c3":
            return [("1.7B", clips, "first 8 of the 128 WS windows (scipy band-pass included)")]
        return [("0.6B", [c for c, i in zip(clips, idx) if not flush[i]], "the partial windows among the first 8 WS windows on 0.6B"),
                ("1.7B", [c for c, i in zip(clips, idx) if flush[i]], "the flush windows among the first 8 WS windows on 1.7B")]
    lens = synth.workload_c4_lengths()
    return [("1.7B", synth.workload_c4_clips(lens, [0, 1]), "first 2 of the hour's segments")]


def run_cpu_arm(config: str, steps: int, warmup: int, bounded: bool):
    from qwen3_asr_b200.synth import model_config, random_weights

    cores = os.cpu_count() or 1
    parts = cpu_sample(config, bounded)
    prepared = []
    for name, clips, _ in parts:
        cfg = model_config(name)
        prepared.append((cfg, random_weights(cfg, seed=0), clips))
    for _ in range(max(1, min(warmup, 1))):
        for cfg, w, clips in prepared:
            oracle_step(w, cfg, clips[:1], cores)
    t, outs = 0.0, []
    for _ in range(steps):
        outs = []
        for cfg, w, clips in prepared:
            dt, out, toks = oracle_step(w, cfg, clips, cores)
            t += dt
            outs.append((out, toks))
    audio_s = sum(sum(len(c) for c in clips) for _, _, clips in prepared) / SR * steps
    sample = "; ".join(d for _, _, d in parts) + f" per step, torch fp32, {cores} threads"
    return audio_s / t, 1e3 * t / steps, cores, sample, outs


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.steps if args.steps is not None else REF_STEPS[args.config]
    value, ms, cores, sample, _ = run_cpu_arm(args.config, steps, args.warmup, bounded=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.config == "c4" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(config: str = "c2"):
    common = {"parallelism": "clip-sharded replicas, no collective"}
    if config == "c1":
        return {"workload": "C1: Qwen3-ASR-0.6B log-mel + audio-encoder forward, one 5 s 16 kHz clip, batch 1 (BASELINE.json configs[0]) on the "
                            "GPU; random-init weights, 0.1 * N(0, 1) noise", "clips_per_gpu": 1, "clip_seconds": 5.0,
                "cache": "inputs larger than L2: every step streams the 373 MB of encoder weights through a 126 MB L2", **common}
    if config == "c3":
        return {"workload": "C3: Qwen3-ASR-1.7B, WebSocket sliding-window re-encode: 128 concurrent streams, windows of 0.45-6 s as int16 PCM, "
                            "every 4th a flush (+600 ms silence) (BASELINE.json configs[2]); int16 -> /32768 -> 300-3400 Hz band-pass -> "
                            "log-mel -> encoder as ONE ragged batch; random-init weights, reference 'speech-like' audio",
                "clips_per_gpu": 128, "cache": "inputs larger than L2: 635 MB of weights per step", **common}
    if config == "c5":
        return {"workload": "C5(ii): the C3 windows with BOTH models resident (BASELINE.json configs[4]): the 96 partial windows on Qwen3-ASR-0.6B, "
                            "the 32 flush windows on 1.7B (src/server.py:1351), one ragged batch per model per step",
                "clips_per_gpu": 128, "cache": "inputs larger than L2: 373 + 635 MB of weights per step", **common}
    if config == "c4":
        return {"workload": "C4: Qwen3-ASR-1.7B, one hour of synthetic audio silence-split into ~230 segments of 1-30 s, clip-sharded over the GPUs "
                            "by LPT on mel frames, micro-batches <= 13 312 tokens (BASELINE.json configs[3]); random-init weights",
                "total_seconds": 3600, "cache": "inputs larger than L2: 230 MB PCM, 635 MB weights, ~14 GB of activations per step",
                "parallelism": "clip-sharded by LPT, no collective"}
    return {
        "workload": f"C2: Qwen3-ASR-1.7B log-mel + audio-encoder forward, {N_CLIPS} x {CLIP_SECONDS:.0f} s 16 kHz clips per GPU "
                    "(BASELINE.json configs[1]); random-init weights, reference 'speech-like' synthetic audio",
        "clips_per_gpu": N_CLIPS, "clip_seconds": CLIP_SECONDS,
        "cache": "inputs larger than L2: each step streams 61 MB PCM, 635 MB weights and ~4 GB of activations through a 126 MB L2",
        **common,
    }


# ------------------------------------------------------------------------------------------------ the workloads on the GPU
class FloatWork:
    """c1 / c2 / c4: float32 PCM clips through qasr_encode_pcm (device-resident) and qasr_submit_pcm_host / qasr_wait (end to end)."""

    def __init__(self, enc, clips, dev):
        self.enc = enc
        self.encs = [enc]
        lens = [len(c) for c in clips]
        self.audio_s = sum(lens) / SR
        self.offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        self.pcm_host = torch.empty(int(self.offs[-1]), dtype=torch.float32, pin_memory=True)
        for i, c in enumerate(clips):
            self.pcm_host[int(self.offs[i]):int(self.offs[i + 1])] = torch.from_numpy(c)
        self.pcm_dev = self.pcm_host.to(dev)
        n_tok = int(sum(enc.token_len(n // 160) for n in lens))
        self.out_dev = torch.empty((n_tok, enc.output_dim), dtype=torch.bfloat16, device=dev)
        self.out_host = [torch.empty((n_tok, enc.output_dim), dtype=torch.bfloat16, pin_memory=True) for _ in range(2)]
        self.pending, self.n = [], 0
        self.h2d, self.d2h = int(self.pcm_host.numel() * 4), int(self.out_host[0].numel() * 2)
        self.api = ("qasr_submit_pcm_host + qasr_wait (C ABI, pinned host buffers; batch i+1 submitted before batch i is awaited; "
                    "time = max(device events, host wall clock) over K submits and the final drain)")

    def step_dev(self):
        self.enc.encode_pcm_packed(self.pcm_dev, self.offs, self.out_dev)

    def step_e2e(self):
        # the serving loop: batch i+1 is submitted (H2D copy + compute enqueued) before batch i's result is awaited, so the
        # copies of one batch overlap the compute of its neighbours; every step still moves its own input and output
        buf = self.out_host[self.n % 2]
        self.n += 1
        t, _ = self.enc.submit_pcm_host(self.pcm_host, self.offs, buf)
        self.pending.append(t)
        if len(self.pending) > 1:
            self.enc.wait(self.pending.pop(0))

    def drain(self):
        while self.pending:
            self.enc.wait(self.pending.pop(0))

    def result_host(self):
        self.enc.encode_pcm_host(self.pcm_host, self.offs, self.out_host[0])
        return [self.out_host[0]]


class WindowWork:
    """c3 / c5: int16 WS windows through qasr_ws_window + qasr_encode_pcm; `groups` = [(encoder, window indices)] (one per model)."""

    def __init__(self, groups, wins, flush, dev):
        from qwen3_asr_b200 import B200PreFrontend

        self.groups = []
        self.encs = [g[0] for g in groups]
        self.audio_s = sum(len(w) + (9600 if f else 0) for w, f in zip(wins, flush)) / SR
        self.h2d = self.d2h = 0
        for enc, idx in groups:
            pre = B200PreFrontend(enc)
            arrs = [wins[i] for i in idx]
            fl = [flush[i] for i in idx]
            offs = np.concatenate([[0], np.cumsum([len(a) for a in arrs])]).astype(np.int64)
            host = torch.empty(int(offs[-1]), dtype=torch.int16, pin_memory=True)
            for i, a in enumerate(arrs):
                host[int(offs[i]):int(offs[i + 1])] = torch.from_numpy(a)
            lens = [max(len(a) + (9600 if f else 0), 8000) for a, f in zip(arrs, fl)]
            n_tok = int(sum(enc.token_len(n // 160) for n in lens))
            out_dev = torch.empty((n_tok, enc.output_dim), dtype=torch.bfloat16, device=dev)
            out_host = torch.empty((n_tok, enc.output_dim), dtype=torch.bfloat16, pin_memory=True)
            self.groups.append({"enc": enc, "pre": pre, "offs": offs, "flush": fl, "host": host, "dev": host.to(dev), "out_dev": out_dev,
                                "out_host": out_host, "n_tok": n_tok})
            self.h2d += int(host.numel() * 2)
            self.d2h += int(out_host.numel() * 2)
        self.api = ("B200PreFrontend.encode_windows path: pinned int16 -> H2D -> qasr_ws_window -> qasr_encode_pcm -> D2H into pinned memory, "
                    "synchronised every step (C ABI; time = max(device events, host wall clock))")

    def _run(self, g, dev16):
        pcm, offs = g["pre"].ws_window_packed(dev16, g["offs"], g["flush"])
        g["enc"].encode_pcm_packed(pcm, offs, g["out_dev"])

    def step_dev(self):
        for g in self.groups:
            self._run(g, g["dev"])

    def step_e2e(self):
        for g in self.groups:
            self._run(g, g["host"].to(g["dev"].device, non_blocking=True))
            g["out_host"].copy_(g["out_dev"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def drain(self):
        pass

    def result_host(self):
        self.step_e2e()
        return [g["out_host"] for g in self.groups]


def build_work(args, rank, world, dev):
    from qwen3_asr_b200 import B200AudioEncoder, synth

    def make(name, **kw):
        cfg = synth.model_config(name)
        return B200AudioEncoder(cfg, synth.random_weights(cfg, seed=0), device=dev.index, quantize=args.quantize, **kw)

    c = args.config
    if c == "c1":
        return FloatWork(make("0.6B", max_chunks=64), synth.workload_c1(seed=rank), dev), "weak"
    if c == "c2":
        n = int(CLIP_SECONDS * SR)
        return FloatWork(make("1.7B"), [synth.speech_like(n, rank * N_CLIPS + i) for i in range(N_CLIPS)], dev), "weak"
    if c == "c3":
        wins, flush = synth.workload_c3(seed0=1000 + 128 * rank)
        return WindowWork([(make("1.7B"), list(range(len(wins))))], wins, flush, dev), "weak"
    if c == "c5":
        wins, flush = synth.workload_c3(seed0=1000 + 128 * rank)
        partial = [i for i, f in enumerate(flush) if not f]
        final = [i for i, f in enumerate(flush) if f]
        return WindowWork([(make("0.6B"), partial), (make("1.7B"), final)], wins, flush, dev), "weak"
    lens = synth.workload_c4_lengths()
    mine = synth.lpt_assign([n // 160 for n in lens], world)[rank]
    work = FloatWork(make("1.7B"), synth.workload_c4_clips(lens, mine), dev)
    work.total_audio_s = sum(lens) / SR
    work.imbalance = None
    return work, "strong"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="which BASELINE.json config to run (default c2 = configs[1], the one the metric is quoted on)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quantize", default=None, choices=["fp8", "fp8_per_row"],
                    help="BASELINE.json configs[4](i): e4m3 Linears with dynamic activation scales (the reference's QUANTIZE=fp8); default bf16")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    if args.steps is None:
        args.steps = DEFAULT_STEPS[args.config]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this backend has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if local_rank == 0:
        from qwen3_asr_b200.build import build as build_lib

        build_lib()  # no-op when the in-tree libqasr_b200.so is newer than its sources
    if dist is not None:
        dist.barrier()  # nobody loads the library while local rank 0 may be (re)linking it

    peaks = load_peaks()
    work, scaling = build_work(args, rank, world, dev)
    encs = work.encs
    enc = encs[-1]
    # whole-job audio per step: every rank's own batch under weak scaling, the fixed hour under strong scaling
    audio_s_job = work.total_audio_s if scaling == "strong" else world * work.audio_s

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, after=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        timed.host_ms = (time.time() - w0) * 1e3 / steps   # host time to ENQUEUE one step (what the calling thread is busy for)
        if after is not None:
            after()
        e1.record()
        barrier()
        w1 = time.time()
        return max_over_ranks(e0.elapsed_time(e1)), max_over_ranks((w1 - w0) * 1e3), (w0, w1)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # log-mel kernel alone on a working set >> L2 (8 x the C2 batch: 491 MB PCM in, 393 MB log-mel out), timed BEFORE the encoder
    # steps: the HBM peak it is divided by (MEASURED_PEAKS.json) is itself the figure of a kernel timed alone, and after a second of
    # GEMMs at the board's power cap the SM clock this (instruction-issue-bound) kernel then runs at is 10-25 % lower
    mel_prof, mel_clocks = None, None
    if rank == 0 and args.config == "c2":
        reps = 8
        n = int(CLIP_SECONDS * SR)
        big = work.pcm_dev.repeat(reps)
        big_offs = np.arange(N_CLIPS * reps + 1, dtype=np.int64) * n
        for _ in range(3):
            enc.logmel_packed(big, big_offs)
        torch.cuda.synchronize()
        enc.profile(True)
        t0 = time.time()
        for _ in range(400):
            enc.logmel_packed(big, big_offs)
        torch.cuda.synchronize()
        t1 = time.time()
        mel_prof = enc.profile_read()
        enc.profile(False)
        mel_clocks = sampler.median_sm_mhz(t0, t1) if sampler is not None else None
        del big
        torch.cuda.empty_cache()
    for _ in range(args.warmup):
        work.step_dev()
    torch.cuda.synchronize()
    l0 = sum(e.launch_count for e in encs)
    ms_dev, _, win_dev = timed(work.step_dev, args.steps)
    host_ms_dev = timed.host_ms
    launches = sum(e.launch_count for e in encs) - l0
    graph_stats = [e.graph_stats() for e in encs]

    # same K steps again with per-launch CUDA events on the launching stream -> per-kernel durations
    for e in encs:
        e.profile(True)
    ms_prof, _, win_prof = timed(work.step_dev, args.steps)
    prof = {}
    for e in encs:
        for k, v in e.profile_read().items():
            acc = prof.setdefault(k, {"ms": 0.0, "work": 0.0, "launches": 0})
            acc["ms"] += v["ms"]
            acc["work"] += v["work"]
            acc["launches"] += v["launches"]
        e.profile(False)

    for _ in range(2):
        work.step_e2e()
    work.drain()
    # timed region = K end-to-end steps + the final drain (all K results on the host), events on the compute stream + wall clock
    ms_e2e_ev, ms_e2e_wall, win_e2e = timed(work.step_e2e, args.steps, after=work.drain)
    ms_e2e = max(ms_e2e_ev, ms_e2e_wall)  # the D2H copies run on a side stream: the wall clock bounds them too
    results_host = work.result_host()     # leave the reference output of this batch on the host for the CPU check

    clocks = None
    if sampler is not None:
        for w in (win_dev, win_prof, win_e2e):
            sampler.window(*w)
        clocks = sampler.stop()

    if rank == 0:
        value = audio_s_job * args.steps / (ms_dev / 1e3)
        e2e_value = audio_s_job * args.steps / (ms_e2e / 1e3)
        gemm_names = [k for k in prof if k.endswith("_gemm")]
        gemm_ms = sum(prof[k]["ms"] for k in gemm_names)
        gemm_flops = sum(prof[k]["work"] for k in gemm_names)
        gemm_launches = sum(prof[k]["launches"] for k in gemm_names)
        total_prof_ms = sum(v["ms"] for v in prof.values())
        achieved_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        roofline = {
            "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM family: conv2/conv3 implicit GEMM, conv_out, qkv, out_proj, fc1, fc2, proj1, proj2)",
            "achieved": achieved_tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["bf16_tflops"],
            "peak_source": f"{peaks['source']} (cuBLAS bf16 sustained; burst {peaks['bf16_tflops_burst']})",
            "algorithmic_flops_per_step": gemm_flops / args.steps, "launches_per_step": gemm_launches / args.steps,
            "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None,
            "timing": "per-launch CUDA events on the launching stream over a second pass of the same K steps",
            # dram__bytes_read.sum + dram__bytes_write.sum of the family's 101 launches of one C2 step (conv2 4.59 GB, conv3 0.95 GB,
            # conv_out 0.26 GB, per layer qkv 60 MB + out_proj 54 MB + fc1 85 MB + fc2 150 MB), ncu --set full, divided by 101
            "traffic": NCU_GEMM_DRAM_BYTES_PER_STEP / 101.0 if args.config == "c2" and args.quantize is None else None,
            "traffic_source": "profiles/r02p_gemm_raw_summary.txt (one ncu --set full capture of each kernel of the family, C2 batch, final build)",
        }
        kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                       "tflops": (v["work"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and k != "logmel" and v["work"] > 0 else None}
                   for k, v in prof.items()}
        roofline_mel = None
        if mel_prof and "logmel" in mel_prof:
            m = mel_prof["logmel"]
            fin = mel_prof.get("logmel_finish", {"ms": 0.0})
            gbs = m["work"] / ((m["ms"] + fin["ms"]) / 1e3) / 1e9
            roofline_mel = {"bound": "hbm", "kernel": "logmel_kernel_v3 (persistent, ticketed, warp-autonomous FFT stages, mbarrier ring of power tiles, clamp tiles riding on later items)", "achieved": gbs, "peak": peaks["hbm_gbs"],
                            "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
                            "algorithmic_bytes_per_launch": m["work"] / m["launches"], "avg_launch_ms": (m["ms"] + fin["ms"]) / m["launches"],
                            "workload": "8 x C2 batch (256 x 30 s): 491 MB PCM in, 393 MB log-mel out, >> L2",
                            # dram__bytes_read.sum 516.7 MB + dram__bytes_write.sum 408.5 MB of one launch on this workload
                            "traffic": NCU_MEL_DRAM_BYTES_PER_LAUNCH, "traffic_source": "profiles/r02o_mel_v3_final_summary.txt",
                            "timing": "per-launch CUDA events on the launching stream, 400 launches back to back before the encoder steps",
                            "sm_mhz": mel_clocks,
                            "audio_s_per_s": 8 * N_CLIPS * CLIP_SECONDS * m["launches"] / ((m["ms"] + fin["ms"]) / 1e3)}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            v, _, cores, sample, outs = run_cpu_arm(args.config, 1, 1, bounded=False)
            # the same clips are the first rows of every model group's output: check the GPU result against the CPU one
            err = 0.0
            for (ref_out, ref_toks), got in zip(outs, results_host):
                g = got[: sum(ref_toks)].float()
                err = max(err, float((g - ref_out).abs().max() / ref_out.abs().max()))
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + ", 1 warm-up",
                            "max_rel_err_gpu_vs_cpu": err}
        weight_bytes = sum(e.weight_bytes for e in encs)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16" if args.quantize is None else f"e4m3 Linears ({args.quantize}) + bf16 convs/attention", "data": "synthetic",
            "config": workload_config(args.config),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": work.h2d, "d2h_bytes_per_step": work.d2h, "api": work.api},
            "gpu_launches": int(launches),
            # time the calling thread spends enqueueing one step; with the graph cache a repeated shape is one cudaGraphLaunch
            "host_enqueue_ms_per_step": host_ms_dev, "graph_cache": graph_stats,
            "clocks": clocks,
            "roofline": roofline, "roofline_mel": roofline_mel, "kernels": kernels,
            "ms_per_step_profiled": ms_prof / args.steps,
            "cpu_baseline": cpu_baseline,
            "device_bytes": sum(e.device_bytes for e in encs),
            "audio_s_per_step": audio_s_job,
            # batch-1 / small-batch regime: the floor is reading every weight once from HBM
            "weight_floor_ms": weight_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3,
        }
        print(json.dumps(line), flush=True)
    for e in encs:
        e.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

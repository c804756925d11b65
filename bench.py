#!/usr/bin/env python
"""Benchmark of the hot path: PCM -> log-mel -> Qwen3-ASR audio-encoder hidden states.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = one pass of the hot path over one batch of synthetic input.  Workload = BASELINE.json
configs[1]: Qwen3-ASR-1.7B dims, 32 x 30 s 16 kHz clips per GPU (weak scaling: clips are independent,
no data-path collective; NCCL is used only for the timing barrier / max-over-ranks).  Prints ONE JSON
line on rank 0.  `--impl reference` times the reference's CPU torch path (the oracle restatement of the
transformers classes the reference calls) on the host cores instead.
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
NCU_GEMM_DRAM_BYTES_PER_STEP = (3.192950 + 0.724703 + 0.782233 + 0.177196 + 0.241311 + 0.023039 + 24 * (
    0.032128 + 0.027506 + 0.053372 + 0.000543 + 0.034240 + 0.053047 + 0.137746 + 0.013343)) * 1e9
NCU_MEL_DRAM_BYTES_PER_LAUNCH = (497.989120 + 389.929216) * 1e6

METRIC = "audio-sec encoded/sec (mel+encoder, 1.7B)"
UNIT = "audio-s/s"
MODEL = "1.7B"
N_CLIPS = 32
CLIP_SECONDS = 30.0
SR = 16000


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


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qwen3_asr_b200.synth import model_config, random_weights, speech_like

    cfg = model_config(MODEL)
    weights = random_weights(cfg, seed=0)
    cores = os.cpu_count() or 1
    sample_clips = 1
    clips = [speech_like(int(CLIP_SECONDS * SR), i) for i in range(sample_clips)]
    for _ in range(max(1, min(args.warmup, 1))):
        oracle_step(weights, cfg, clips, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _, _ = oracle_step(weights, cfg, clips, cores)
        t += dt
    audio_s = sample_clips * CLIP_SECONDS * args.steps
    value = audio_s / t
    sample = f"{sample_clips} x {CLIP_SECONDS:.0f} s clip(s) of the C2 batch per step, torch fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config():
    return {
        "workload": f"C2: Qwen3-ASR-{MODEL} log-mel + audio-encoder forward, {N_CLIPS} x {CLIP_SECONDS:.0f} s 16 kHz clips per GPU "
                    "(BASELINE.json configs[1]); random-init weights, reference 'speech-like' synthetic audio",
        "clips_per_gpu": N_CLIPS, "clip_seconds": CLIP_SECONDS,
        "cache": "inputs larger than L2: each step streams 61 MB PCM, 635 MB weights and ~4 GB of activations through a 126 MB L2",
        "parallelism": "clip-sharded replicas, no collective",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quantize", default=None, choices=["fp8", "fp8_per_row"],
                    help="BASELINE.json configs[4](i): e4m3 Linears with dynamic activation scales (the reference's QUANTIZE=fp8); default bf16")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

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

    from qwen3_asr_b200 import B200AudioEncoder
    from qwen3_asr_b200.synth import model_config, random_weights, speech_like

    peaks = load_peaks()
    cfg = model_config(MODEL)
    weights = random_weights(cfg, seed=0)
    enc = B200AudioEncoder(cfg, weights, device=local_rank, quantize=args.quantize)
    n = int(CLIP_SECONDS * SR)
    clips = [speech_like(n, rank * N_CLIPS + i) for i in range(N_CLIPS)]
    audio_s_per_step = N_CLIPS * CLIP_SECONDS

    offs = np.arange(N_CLIPS + 1, dtype=np.int64) * n
    pcm_host = torch.empty(N_CLIPS * n, dtype=torch.float32, pin_memory=True)
    for i, c in enumerate(clips):
        pcm_host[i * n:(i + 1) * n] = torch.from_numpy(c)
    pcm_dev = pcm_host.to(dev)
    n_tok = N_CLIPS * enc.token_len(n // 160)
    out_dev = torch.empty((n_tok, enc.output_dim), dtype=torch.bfloat16, device=dev)
    out_host = torch.empty((n_tok, enc.output_dim), dtype=torch.bfloat16, pin_memory=True)

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

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        w1 = time.time()
        return max_over_ranks(e0.elapsed_time(e1)), (w0, w1)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    step_dev = lambda: enc.encode_pcm_packed(pcm_dev, offs, out_dev)
    out_host2 = torch.empty_like(out_host).pin_memory()
    pending = []

    def step_e2e():
        # the serving loop: batch i+1 is submitted (H2D copy + compute enqueued) before batch i's result is awaited, so the
        # copies of one batch overlap the compute of its neighbours; every step still moves its own input and output
        buf = out_host if step_e2e.n % 2 == 0 else out_host2
        step_e2e.n += 1
        t, _ = enc.submit_pcm_host(pcm_host, offs, buf)
        pending.append(t)
        if len(pending) > 1:
            enc.wait(pending.pop(0))

    step_e2e.n = 0

    def drain_e2e():
        while pending:
            enc.wait(pending.pop(0))

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    l0 = enc.launch_count
    ms_dev, win_dev = timed(step_dev, args.steps)
    launches = enc.launch_count - l0

    # same K steps again with per-launch CUDA events on the launching stream -> per-kernel durations
    enc.profile(True)
    ms_prof, win_prof = timed(step_dev, args.steps)
    prof = enc.profile_read()
    enc.profile(False)

    for _ in range(2):
        step_e2e()
    drain_e2e()

    # timed region = K submits + the final drain (all K results on the host), events on the compute stream + wall clock
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    drain_e2e()
    e1.record()
    barrier()
    w1 = time.time()
    ms_e2e, win_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0)), (w0, w1)
    ms_e2e_wall = max_over_ranks((w1 - w0) * 1e3)
    ms_e2e = max(ms_e2e, ms_e2e_wall)  # the D2H copies run on a side stream: the wall clock bounds them too
    enc.encode_pcm_host(pcm_host, offs, out_host)  # leave the reference output of this batch in out_host for the CPU check

    # mel kernel alone on a working set >> L2 (8 x the C2 batch: 491 MB PCM in, 393 MB log-mel out)
    mel_prof = None
    if rank == 0:
        reps = 8
        big = pcm_dev.repeat(reps)
        big_offs = np.arange(N_CLIPS * reps + 1, dtype=np.int64) * n
        for _ in range(2):
            enc.logmel_packed(big, big_offs)
        enc.profile(True)
        for _ in range(5):
            enc.logmel_packed(big, big_offs)
        mel_prof = enc.profile_read()
        enc.profile(False)
        del big
    clocks = None
    if sampler is not None:
        for w in (win_dev, win_prof, win_e2e):
            sampler.window(*w)
        clocks = sampler.stop()

    if rank == 0:
        value = world * audio_s_per_step * args.steps / (ms_dev / 1e3)
        e2e_value = world * audio_s_per_step * args.steps / (ms_e2e / 1e3)
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
            # dram__bytes_read.sum + dram__bytes_write.sum of the family's 101 launches of one step (conv2 3.80 GB, conv3 0.95 GB,
            # conv_out 0.24 GB, per layer qkv 55 MB + out_proj 54 MB + fc1 81 MB + fc2 146 MB), ncu --set full, divided by 101
            "traffic": NCU_GEMM_DRAM_BYTES_PER_STEP / 101.0,
            "traffic_source": "profiles/r01p_gemm_raw_summary.txt (one ncu --set full capture of each kernel of the family, C2 batch)",
        }
        kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                       "tflops": (v["work"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 and k != "logmel" and v["work"] > 0 else None}
                   for k, v in prof.items()}
        roofline_mel = None
        if mel_prof and "logmel" in mel_prof:
            m = mel_prof["logmel"]
            fin = mel_prof.get("logmel_finish", {"ms": 0.0})
            gbs = m["work"] / ((m["ms"] + fin["ms"]) / 1e3) / 1e9
            roofline_mel = {"bound": "hbm", "kernel": "logmel_kernel (persistent, ticketed frame + clamp items)", "achieved": gbs, "peak": peaks["hbm_gbs"],
                            "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
                            "algorithmic_bytes_per_launch": m["work"] / m["launches"], "avg_launch_ms": (m["ms"] + fin["ms"]) / m["launches"],
                            "workload": "8 x C2 batch (256 x 30 s): 491 MB PCM in, 393 MB log-mel out, >> L2",
                            # dram__bytes_read.sum 498.0 MB + dram__bytes_write.sum 389.9 MB of one launch on this workload
                            "traffic": NCU_MEL_DRAM_BYTES_PER_LAUNCH, "traffic_source": "profiles/r01c_mel_ticketed_summary.txt",
                            "audio_s_per_s": 8 * audio_s_per_step * m["launches"] / ((m["ms"] + fin["ms"]) / 1e3)}
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample_n = 2
            oracle_step(weights, cfg, clips[:1], cores)  # warm-up
            dt, ref_out, ref_toks = oracle_step(weights, cfg, clips[:sample_n], cores)
            got = out_host[: sum(ref_toks)].float()
            err = float((got - ref_out).abs().max() / ref_out.abs().max())
            cpu_baseline = {"value": sample_n * CLIP_SECONDS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"first {sample_n} of the {N_CLIPS} clips of the same batch, torch fp32 ({cores} threads), 1 warm-up",
                            "max_rel_err_gpu_vs_cpu": err}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.quantize is None else f"e4m3 Linears ({args.quantize}) + bf16 convs/attention", "data": "synthetic",
            "config": workload_config(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(pcm_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 2),
                    "api": "qasr_submit_pcm_host + qasr_wait (C ABI, pinned host buffers; batch i+1 submitted before batch i is awaited; "
                           "time = max(device events, host wall clock) over K submits and the final drain)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline, "roofline_mel": roofline_mel, "kernels": kernels,
            "ms_per_step_profiled": ms_prof / args.steps,
            "cpu_baseline": cpu_baseline,
            "device_bytes": enc.device_bytes,
        }
        print(json.dumps(line), flush=True)
    enc.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

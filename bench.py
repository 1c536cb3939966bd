#!/usr/bin/env python
"""bench.py -- BEATs embedding-extraction throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (waveform -> fbank -> BEATs-base encoder -> mean-pooled 768-d embedding) over
one batch of synthetic 10 s clips PER GPU (BASELINE.json configs[1]: 256 x 10 s, bf16; clip-sharded data parallel,
weak scaling), followed for N > 1 by the one collective of the path: an NCCL all-gather of the pooled embeddings.

Prints ONE JSON line.  `value` = audio-hours/s with inputs resident in HBM; `e2e` = the same metric through the plugin
API (`Model.forward`, classifier-free pooled mode) with pinned HOST inputs, H2D and D2H inside the timed region;
`roofline` = the dominant kernel (tcgen05 GEMM) against the measured bf16 peak, from CUDA events recorded around
its launches; `cpu_baseline` = the torch-CPU oracle port timed on this box's host cores on a bounded sample.
`--impl reference` times the CPU oracle port alone (the reference is pure Python/torch and cannot travel to the box).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 10
SAMPLE_RATE = 16000
BATCH_PER_GPU = 256
FLOPS_PER_CLIP_10S = 98.60e9  # BASELINE.md section 3 (2*MAC; patch-embed, pos-conv, QKV, QK^T, PV, out, FFN, gate)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


def gemm_traffic_from_profile():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum), averaged over the GEMM
    launches of one bench step, from the committed ncu launch list of this same command (profiles/launches_*.csv)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "launches_r*.csv")))
    if not files:
        return None, None
    n_tot, bytes_tot = 0, 0.0
    for ln in open(files[-1]):
        if ln.startswith("gemm_bf16_kernel"):
            f = ln.rstrip("\n").rsplit(",", 6)
            try:
                n, rd, wr = int(f[1]), float(f[5]), float(f[6])
            except (ValueError, IndexError):
                continue
            n_tot += n
            bytes_tot += n * (rd + wr) * 1e6
    return (bytes_tot / n_tot, os.path.basename(files[-1])) if n_tot else (None, None)


# ------------------------------------------------------------------------------------------------------------
# CPU baseline: the numpy oracle port (oracle/), all host threads (OpenBLAS), bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_time(n_clips: int, reps: int = 1):
    """Seconds for one pass of the torch-CPU oracle port (oracle/beats_torch.py: the reference's own op mix on
    MKL / oneDNN with every host thread) over `n_clips` 10 s clips."""
    import numpy as np
    import torch

    from oracle import beats_encoder as OE
    from oracle import beats_torch as OT
    from oracle.weights import make_beats_weights

    torch.set_num_threads(os.cpu_count() or 1)
    dims = OE.BeatsDims()
    W = OT.to_torch(make_beats_weights(dims, seed=0, init="reference"))
    wav = torch.randn(n_clips, CLIP_SECONDS * SAMPLE_RATE, generator=torch.Generator().manual_seed(1234)) * 0.1
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        chunks = []
        for i in range(0, n_clips, 8):  # 8-clip chunks keep the materialised [B,12,N,N] mask (94 MB each) cache-friendly
            chunks.append(OT.beats_forward(W, wav[i : i + 8], None, dims)["x"].mean(dim=1))
        pooled = torch.cat(chunks)
        best = min(best, time.perf_counter() - t0)
    assert pooled.shape == (n_clips, 768)
    return best


CPU_SAMPLE_CLIPS = 32


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_clips = CPU_SAMPLE_CLIPS
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_oracle_time(n_clips)
    steps = max(1, min(args.steps, 5))
    ts = [cpu_oracle_time(n_clips) for _ in range(steps)]
    sec = sum(ts) / len(ts)
    value = n_clips * CLIP_SECONDS / 3600.0 / sec
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference",
        "metric": "beats_embed_throughput", "value": value, "unit": "audio-hours/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BEATs-base embedding extraction, 10 s clips @16 kHz, mean-pooled 768-d (BASELINE configs[1])",
                   "sample": f"{n_clips} clips per step on host CPU", "weights": "random-init (reference distributions)"},
        "cpu_baseline": {"value": value, "unit": "audio-hours/s", "cores": cores, "kind": "port",
                         "sample": f"{n_clips} x 10 s clips per step, torch-CPU oracle port of the reference path (oracle/beats_torch.py, {cores} threads)"},
        "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
            for ln in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.rows.append([c.strip() for c in ln.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from avex_b200 import _lib, plugin
    from avex_b200.plugin import beats_model  # noqa: F401

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; avex_b200 has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # model through the plugin surface, random-init weights of the named architecture (esp_aves2_sl_beats_all.yml)
    init_config = dict(encoder_layers=12, encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12,
                       deep_norm=True, dropout=0.0, attention_dropout=0.0, finetuned_model=False,
                       layer_wise_gradient_decay_ratio=0.6)  # fmt: skip
    plugin.register_model("bench_beats_base", plugin.ModelSpec(name="beats", device="cuda", init_config=init_config))
    torch.manual_seed(0)
    model = plugin.load_model("bench_beats_base", device="cuda", return_features_only=True).eval()
    bk = model.backbone
    bk.precision = args.precision  # "bf16" (the metric) or "fp32" (validation mode; informational)

    B, T = args.batch, CLIP_SECONDS * SAMPLE_RATE
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    wav_dev = torch.randn(B, T, device=dev, generator=g) * 0.1
    host = [torch.empty(B, T, dtype=torch.float32).pin_memory() for _ in range(2)]
    host[0].copy_(wav_dev.cpu())
    host[1].copy_(host[0])
    pooled_host = torch.empty(B, 768, dtype=torch.float32).pin_memory()
    gathered = torch.empty(world * B, 768, device=dev) if world > 1 else None

    def step_device():
        res = bk.run(wav_dev, None, want_features=False, want_pooled=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, res["pooled"])
        return res["pooled"]

    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [torch.empty(B, T, device=dev) for _ in range(2)]

    ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def step_e2e(i, ev_ready):
        """pinned host batch -> H2D (copy stream, double-buffered) -> forward -> D2H of the pooled embeddings."""
        cur = i & 1
        main = torch.cuda.current_stream()
        # the next batch's H2D is enqueued BEFORE this batch's 69 kernels, so that it overlaps them (enqueued after, the copy
        # started late and 2.9 ms of the 3.0 ms transfer showed up in the step: tools/e2e_probe.py)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[cur ^ 1])  # the forward that last read that buffer has finished
            dev_in[cur ^ 1].copy_(host[cur ^ 1], non_blocking=True)
            ev_ready[cur ^ 1].record(copy_stream)
        main.wait_event(ev_ready[cur])  # this batch has landed in HBM
        res = bk.run(dev_in[cur], None, want_features=False, want_pooled=True)
        ev_consumed[cur].record(main)
        if world > 1:
            dist.all_gather_into_tensor(gathered, res["pooled"])
        pooled_host.copy_(res["pooled"], non_blocking=True)
        return res["pooled"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    W_, K_ = max(3, args.warmup), args.steps
    for _ in range(W_):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = _lib.launch_count()
    ms = timed(lambda i: step_device(), K_)
    launches = _lib.launch_count() - l0
    ms_per_step = ms / K_
    clips_per_s = world * B / (ms_per_step / 1e3)
    value = clips_per_s * CLIP_SECONDS / 3600.0

    # ---- e2e: host buffers, copies inside the timed region --------------------------------------------------------
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    with torch.cuda.stream(copy_stream):
        dev_in[0].copy_(host[0], non_blocking=True)
        ev_ready[0].record(copy_stream)
    for i in range(2):
        step_e2e(i, ev_ready)
    torch.cuda.synchronize()
    # re-arm buffer 0 for the timed loop
    with torch.cuda.stream(copy_stream):
        dev_in[0].copy_(host[0], non_blocking=True)
        ev_ready[0].record(copy_stream)
    ms_e2e = timed(lambda i: step_e2e(i, ev_ready), K_)
    e2e_value = world * B / (ms_e2e / K_ / 1e3) * CLIP_SECONDS / 3600.0
    if rank == 0:
        sampler.stop()

    # ---- per-kernel CUDA-event profile of the same step (second pass; not part of `value`) ---------------------------
    lib.avexk_profile_enable(1)
    prof_steps = min(K_, 3)
    for _ in range(prof_steps):
        step_device()
    torch.cuda.synchronize()
    prof = {}
    for kid, name in enumerate(["fbank", "gemm", "attention", "layernorm", "posconv"]):
        n, t, w = C.c_longlong(), C.c_double(), C.c_double()
        lib.avexk_profile_read(kid, C.byref(n), C.byref(t), C.byref(w))
        prof[name] = {"launches": n.value // prof_steps, "ms_per_step": t.value / prof_steps, "work_per_step": w.value / prof_steps}
    lib.avexk_profile_enable(0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    gm = prof["gemm"]
    tf = gm["work_per_step"] / (gm["ms_per_step"] * 1e-3) / 1e12 if gm["ms_per_step"] > 0 else 0.0
    roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": tf / pk["tf_sustained"], "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
            "traffic": None, "launches_per_step": gm["launches"], "avg_launch_ms": gm["ms_per_step"] / max(1, gm["launches"]),
            "share_of_step": gm["ms_per_step"] / ms_per_step}  # fmt: skip
    traffic, src = gemm_traffic_from_profile()
    if traffic is not None:
        roof["traffic"] = traffic
        roof["traffic_source"] = f"profiles/{src}: mean dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (ncu, same command)"
        # algorithmic HBM bytes per launch for comparison: operands read once, outputs written once (per-step totals / launches)
        roof["algorithmic_bytes_per_launch"] = B * 496 * (12 * (768 * 2 + 2304 * 2) + 12 * (768 * 2 + 768 * 4 + 768 * 6) + 12 * (768 * 2 + 3072 * 2)
                                                          + 12 * (3072 * 2 + 768 * 4 + 768 * 6) + (768 * 2 + 512 * 4) + (1536 * 2 + 768 * 4)) / max(1, gm["launches"])
    fbk = prof["fbank"]
    fb_gbs = fbk["work_per_step"] / (fbk["ms_per_step"] * 1e-3) / 1e9 if fbk["ms_per_step"] > 0 else 0.0
    kernels = {k: {"ms_per_step": round(v["ms_per_step"], 4), "launches": v["launches"]} for k, v in prof.items()}
    kernels["fbank"]["achieved_GBps"] = fb_gbs
    kernels["fbank"]["hbm_frac"] = fb_gbs / pk["hbm_gbs"]
    for k in ("attention", "posconv"):
        v = prof[k]
        kernels[k]["TFLOPs"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12 if v["ms_per_step"] > 0 else 0.0

    # ---- CPU baseline on this box's host cores (bounded sample) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = CPU_SAMPLE_CLIPS
        cpu_oracle_time(8)  # warm-up
        sec = cpu_oracle_time(n_cpu)
        cores = os.cpu_count() or 1
        cpu = {"value": n_cpu * CLIP_SECONDS / 3600.0 / sec, "unit": "audio-hours/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} x 10 s clips, one pass, torch-CPU oracle port of the reference path (oracle/beats_torch.py, {cores} threads), {sec:.2f} s"}  # fmt: skip

    line = {
        "metric": "beats_embed_throughput", "value": value, "unit": "audio-hours/s", "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32 (3-term split-bf16 GEMMs, fp32 attention / pos-conv)",
        "data": "synthetic",
        "config": {"workload": "BEATs-base embedding extraction, 256 x 10 s clips @16 kHz per GPU, mean-pooled 768-d (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": world * B, "clip_seconds": CLIP_SECONDS, "tokens_per_clip": 496,
                   "weights": "random-init, reference init distributions, seed 0", "parallelism": f"clip-sharded dp{world}",
                   "l2": "inputs and activations (0.16-3.9 GB per step) exceed the 126 MB L2; no flush needed",
                   "clips_per_s": clips_per_s, "model_tflops": clips_per_s * FLOPS_PER_CLIP_10S / 1e12},
        "e2e": {"value": e2e_value, "unit": "audio-hours/s", "h2d_bytes_per_step": world * B * T * 4, "d2h_bytes_per_step": world * B * 768 * 4,
                "ms_per_step": ms_e2e / K_, "api": "plugin Model backbone.run(want_pooled) with pinned host input, double-buffered H2D"},
        "gpu_launches": launches,
        "roofline": roof,
        "kernels": kernels,
        "encoder_tensor_frac": clips_per_s * FLOPS_PER_CLIP_10S / 1e12 / world / pk["tf_sustained"],
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
# secondary workload: EfficientNet-B0 mel-spectrogram feature extractor (BASELINE configs[2])
# ------------------------------------------------------------------------------------------------------------
EFF_CLIP_SECONDS = 5
EFF_ACT_BYTES_PER_CLIP = 35.0e6      # ~35 MB of bf16 activation traffic per 5 s clip (unfused NHWC path)
EFF_MEL_BYTES_PER_CLIP = 4 * 80000 + 4 * 128 * 501  # waveform in, single-channel log-mel out


def run_effnet(args):
    """`--workload effnet`: 5 s clips -> mel -> EfficientNet-B0 features [B,1280,4,16] through the plugin Model.  Prints one
    JSON line of the same shape as the main workload (metric effnet_b0_feature_throughput, clips/s)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from avex_b200 import _lib, plugin
    from avex_b200.plugin import efficientnet_model  # noqa: F401
    from oracle.weights import make_effnet_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; avex_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = plugin.ModelSpec(name="efficientnet", device="cuda", efficientnet_variant="b0",
                            audio_config=dict(sample_rate=16000, n_fft=800, hop_length=160, win_length=800, window="hann", n_mels=128,
                                              representation="mel_spectrogram", normalize=True, target_length_seconds=10,
                                              window_selection="random"))  # fmt: skip
    plugin.register_model("bench_effnet_b0", spec)
    model = plugin.build_model_from_spec(spec, "cuda", pretrained=False, return_features_only=True).eval()
    stats_path = os.path.join(ROOT, "tests", "golden", "effnet_bn_stats.npz")
    W = make_effnet_weights(seed=3, num_classes=0, bn_stats=dict(np.load(stats_path)))  # calibrated BatchNorm statistics
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=False)
    B = args.batch if args.batch != BATCH_PER_GPU else 512 // max(1, world) if world > 1 else 512
    T = EFF_CLIP_SECONDS * SAMPLE_RATE
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    wav_dev = torch.randn(B, T, device=dev, generator=g) * 0.1
    host = torch.empty(B, T, dtype=torch.float32).pin_memory()
    host.copy_(wav_dev.cpu())
    feat_host = torch.empty(B, 1280, dtype=torch.float32).pin_memory()

    def step_device(_i=0):
        with torch.no_grad():
            return model(wav_dev)

    def step_e2e(_i=0):
        x = host.to(dev, non_blocking=True)
        with torch.no_grad():
            f = model(x)
        feat_host.copy_(f.mean(dim=(2, 3)), non_blocking=True)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    W_, K_ = max(3, args.warmup), args.steps
    for _ in range(W_):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = _lib.launch_count()
    ms = timed(step_device, K_) / K_
    launches = _lib.launch_count() - l0
    ms_e2e = timed(step_e2e, K_) / K_
    mel = model._engine.mel
    ms_mel = timed(lambda i: mel.run(wav_dev, normalize=True), K_) / K_
    if rank == 0:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    clips_per_s = world * B / (ms * 1e-3)
    act_gbs = B * EFF_ACT_BYTES_PER_CLIP / (ms * 1e-3) / 1e9
    mel_gbs = B * EFF_MEL_BYTES_PER_CLIP / (ms_mel * 1e-3) / 1e9
    line = {
        "metric": "effnet_b0_feature_throughput", "value": clips_per_s, "unit": "clips/s", "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "EfficientNet-B0 mel-spectrogram feature extractor, 512 x 5 s clips @16 kHz (BASELINE configs[2]), features [B,1280,4,16]",
                   "batch_per_gpu": B, "global_batch": world * B, "clip_seconds": EFF_CLIP_SECONDS, "weights": "random-init, calibrated BatchNorm statistics",
                   "audio_hours_per_s": clips_per_s * EFF_CLIP_SECONDS / 3600.0,
                   "l2": "activations (up to 1.6 GB per layer) exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": world * B * T * 4, "d2h_bytes_per_step": world * B * 1280 * 4,
                "ms_per_step": ms_e2e, "api": "plugin Model.forward with pinned host input; pooled [B,1280] features read back"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "whole forward (NHWC bf16 activations)", "achieved": act_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": act_gbs / pk["hbm_gbs"], "peak_source": f"{pk['src']} hbm_gbs", "traffic": None,
                     "algorithmic_bytes": "35 MB of bf16 activation traffic per 5 s clip (SURVEY 8d)"},
        "kernels": {"melspec": {"ms_per_step": round(ms_mel, 4), "achieved_GBps": mel_gbs, "hbm_frac": mel_gbs / pk["hbm_gbs"],
                                "algorithmic_bytes_per_clip": EFF_MEL_BYTES_PER_CLIP}},
        "cpu_baseline": None,
        "clocks": sampler.summary(),
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="fp32: the <= 1e-3 validation mode (informational)")
    ap.add_argument("--workload", default="beats", choices=["beats", "effnet"],
                    help="beats (default): the BASELINE.json metric; effnet: secondary line for the EfficientNet-B0 path")
    args = ap.parse_args()
    if args.workload == "effnet" and args.impl == "ours":
        run_effnet(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- BEATs embedding-extraction throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (waveform -> fbank -> BEATs-base encoder -> mean-pooled 768-d embedding) over
one batch of synthetic 10 s clips PER GPU (BASELINE.json configs[1]: 256 x 10 s, bf16; clip-sharded data parallel,
weak scaling), followed for N > 1 by the one collective of the path: an NCCL all-gather of the pooled embeddings.

Prints ONE JSON line.
  value        audio-hours/s with inputs resident in HBM (device-timed, max over ranks)
  e2e          the same metric through the reference-facing plugin call -- `model.extract_embeddings(x, aggregation="mean")`
               with a hook on the last layer -- from pinned HOST buffers, H2D and D2H inside the timed region
  roofline     the dominant kernel (tcgen05 GEMM): ALGORITHMIC FLOPs / its CUDA-event time inside the step vs the measured
               sustained bf16 peak (`executed_tflops` also counts the 3x K of the split-bf16 front-end GEMMs)
  strong       global batch 256 split over the N ranks (32 clips per rank at N = 8), same step
  secondary    BASELINE configs[2] (EfficientNet-B0, 512 x 5 s; with its own e2e, roofline, cpu_baseline and gpu_eager_baseline)
               and configs[4] shape (BEATs, 64 x 60 s, 13 pooled hooks)
  gpu_eager_baseline   the reference's own torch modules on the SAME GPU (fp32 and autocast-bf16, batch 32): the incumbent
  cpu_baseline the reference's CPU path on this box's host cores, bounded sample (rank 0, N = 1 only)
`--impl reference` times the UNMODIFIED reference (earthspecies/avex installed under baseline/_ref) on the host cores through its
own public API (`load_model` -> `register_hooks_for_layers` -> `extract_embeddings`); falls back to the torch oracle port when
baseline/_ref is absent.
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
FLOPS_PER_CLIP_60S = 870.07e9
WORKLOAD = "BEATs-base embedding extraction, 256 x 10 s clips @16 kHz per GPU, mean-pooled 768-d (BASELINE configs[1])"
INIT_CONFIG = dict(encoder_layers=12, encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12,
                   deep_norm=True, dropout=0.0, attention_dropout=0.0, finetuned_model=False,
                   layer_wise_gradient_decay_ratio=0.6)  # fmt: skip


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


def gemm_flops_algorithmic(M: int) -> float:
    """2*M*N*K of every nn.Linear / patch-embed GEMM of one forward over M token rows (K NOT tripled for the split front end)."""
    per_layer = 2.0 * M * (2304 * 768 + 768 * 768 + 3072 * 768 + 768 * 3072)
    return 12 * per_layer + 2.0 * M * (512 * 256 + 768 * 512)


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
# the reference itself (unmodified earthspecies/avex from baseline/_ref) and the oracle port as fall-back
# ------------------------------------------------------------------------------------------------------------
def import_reference():
    """Import the real `avex` package from baseline/_ref (pip-installed there, git-ignored, travels with the snapshot).  The three
    third-party modules it imports at top level but never reaches on this path (gcsfs, s3fs, h5py) are absent from the image
    and stubbed.  Returns the module or None."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "avex")):
        return None
    import types

    for name in ("gcsfs", "s3fs", "h5py"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                m = types.ModuleType(name)
                m.GCSFileSystem = type("GCSFileSystem", (), {})
                m.S3FileSystem = type("S3FileSystem", (), {})
                sys.modules[name] = m
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import avex

        return avex
    except Exception as e:  # pragma: no cover
        print(f"bench.py: baseline/_ref present but not importable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        return None


def reference_model(avex, device: str):
    """esp_aves2_sl_beats_all's architecture, random init, through the reference's own registry / loader."""
    import torch
    from avex.models.utils.registry import get_model_spec

    spec = get_model_spec("esp_aves2_sl_beats_all").model_copy(deep=True)
    avex.register_model("bench_reference_beats", spec)
    torch.manual_seed(0)
    model = avex.load_model("bench_reference_beats", device=device, return_features_only=True).eval()
    model.register_hooks_for_layers(["last_layer"])
    return model


class CpuArm:
    """One pass = `n_clips` 10 s clips -> pooled [n, 768] on the host cores, all threads."""

    def __init__(self):
        import torch

        torch.set_num_threads(os.cpu_count() or 1)
        self.torch = torch
        self.avex = import_reference()
        if self.avex is not None:
            self.kind = "reference"
            self.model = reference_model(self.avex, "cpu")
            self.what = "unmodified earthspecies/avex 1.2.0 from baseline/_ref: load_model -> register_hooks_for_layers(['last_layer']) -> extract_embeddings(aggregation='mean')"
        else:
            from oracle import beats_encoder as OE
            from oracle import beats_torch as OT
            from oracle.weights import make_beats_weights

            self.kind = "port"
            self.dims = OE.BeatsDims()
            self.OT = OT
            self.W = OT.to_torch(make_beats_weights(self.dims, seed=0, init="reference"))
            self.what = "torch-CPU oracle port of the reference path (oracle/beats_torch.py)"
        self.wav = None

    def run(self, n_clips: int) -> float:
        torch = self.torch
        if self.wav is None or self.wav.shape[0] != n_clips:
            self.wav = torch.randn(n_clips, CLIP_SECONDS * SAMPLE_RATE, generator=torch.Generator().manual_seed(1234)) * 0.1
        t0 = time.perf_counter()
        chunks = []
        with torch.no_grad():
            for i in range(0, n_clips, 8):  # 8-clip chunks keep the reference's materialised [B,12,N,N] mask (94 MB each) cache-friendly
                x = self.wav[i : i + 8]
                if self.kind == "reference":
                    chunks.append(self.model.extract_embeddings(x, aggregation="mean"))
                else:
                    chunks.append(self.OT.beats_forward(self.W, x, None, self.dims)["x"].mean(dim=1))
        pooled = torch.cat(chunks)
        dt = time.perf_counter() - t0
        assert pooled.shape == (n_clips, 768)
        return dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(0, args.warmup)
    t_probe = arm.run(8)  # also the first-touch warm-up
    # bounded sample: size the per-step batch so that the whole --steps/--warmup run ends within ~3 minutes
    budget = 180.0
    n_clips = int(budget / ((steps + warm) * (t_probe / 8.0)))
    n_clips = max(8, min(BATCH_PER_GPU, (n_clips // 8) * 8))
    for _ in range(warm):
        arm.run(n_clips)
    ts = [arm.run(n_clips) for _ in range(steps)]
    sec = sum(ts) / len(ts)
    value = n_clips * CLIP_SECONDS / 3600.0 / sec
    sample = f"{n_clips} x 10 s clips per step (bounded sample of the 256-clip batch), {arm.what}, {cores} threads"
    line = {
        "impl": "reference",
        "metric": "beats_embed_throughput", "value": value, "unit": "audio-hours/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{n_clips} clips per step on host CPU", "clips_per_step": n_clips,
                   "weights": "random-init (reference init)"},
        "cpu_baseline": {"value": value, "unit": "audio-hours/s", "cores": cores, "kind": arm.kind, "sample": sample},
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


class Dist:
    """torchrun environment + the barrier / max-over-ranks timing contract."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; avex_b200 has no CPU fallback (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize on both sides, CUDA events, max over ranks (ms total)."""
        torch = self.torch
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        self.barrier()
        ms = a.elapsed_time(b)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_ours(args):
    import torch

    from avex_b200 import _lib, plugin
    from avex_b200.plugin import beats_model  # noqa: F401

    D = Dist()
    world, rank, dev = D.world, D.rank, D.dev
    dist = D.dist
    lib = _lib.load()

    # model through the plugin surface, random-init weights of the named architecture (esp_aves2_sl_beats_all.yml)
    plugin.register_model("bench_beats_base", plugin.ModelSpec(name="beats", device="cuda", init_config=INIT_CONFIG))
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

    def step_device(_i=0, x=None, out=None):
        res = bk.run(wav_dev if x is None else x, None, want_features=False, want_pooled=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered if out is None else out, res["pooled"])
        return res["pooled"]

    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [torch.empty(B, T, device=dev) for _ in range(2)]
    ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]
    model.register_hooks_for_layers(["last_layer"])  # the reference's way to a pooled 768-d embedding per clip

    def step_e2e(i, ev_ready):
        """pinned host batch -> H2D (copy stream, double-buffered) -> model.extract_embeddings(aggregation="mean") -> D2H."""
        cur = i & 1
        main = torch.cuda.current_stream()
        # the next batch's H2D is enqueued BEFORE this batch's kernels, so that it overlaps them (tools/e2e_probe.py)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[cur ^ 1])  # the forward that last read that buffer has finished
            dev_in[cur ^ 1].copy_(host[cur ^ 1], non_blocking=True)
            ev_ready[cur ^ 1].record(copy_stream)
        main.wait_event(ev_ready[cur])  # this batch has landed in HBM
        emb = model.extract_embeddings(dev_in[cur], aggregation="mean")
        ev_consumed[cur].record(main)
        if world > 1:
            dist.all_gather_into_tensor(gathered, emb)
        pooled_host.copy_(emb, non_blocking=True)
        return emb

    W_, K_ = max(3, args.warmup), args.steps
    for _ in range(W_):
        step_device()
    sampler = ClockSampler(D.local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = _lib.launch_count()
    ms = D.timed(step_device, K_)
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
    with torch.cuda.stream(copy_stream):  # re-arm buffer 0 for the timed loop
        dev_in[0].copy_(host[0], non_blocking=True)
        ev_ready[0].record(copy_stream)
    ms_e2e = D.timed(lambda i: step_e2e(i, ev_ready), K_)
    e2e_value = world * B / (ms_e2e / K_ / 1e3) * CLIP_SECONDS / 3600.0
    if rank == 0:
        sampler.stop()

    # ---- strong scaling: global batch 256 split over the ranks ---------------------------------------------------------
    strong = None
    if not args.no_extras:
        Bs = max(1, BATCH_PER_GPU // world)
        xs = wav_dev[:Bs]
        gs = torch.empty(world * Bs, 768, device=dev) if world > 1 else None
        for _ in range(3):
            step_device(0, xs, gs)
        ks = max(K_, 10)
        ms_s = D.timed(lambda i: step_device(i, xs, gs), ks) / ks
        strong = {"global_batch": Bs * world, "clips_per_rank": Bs, "ms_per_step": ms_s, "steps": ks,
                  "value": Bs * world / (ms_s * 1e-3) * CLIP_SECONDS / 3600.0, "unit": "audio-hours/s",
                  "note": "same step (forward + all-gather of pooled embeddings) with the fixed global batch of BASELINE configs[1]"}  # fmt: skip

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

    # ---- secondary workloads + the incumbent GPU implementation (every rank takes part; bounded to a few seconds) -------
    secondary, eager = {}, None
    if not args.no_extras:
        secondary["beats_64x60s"] = bench_beats_long(D, model)
        del wav_dev, dev_in, host
        bk._ws = None
        torch.cuda.empty_cache()
        secondary["effnet_512x5s"] = bench_effnet(D, steps=5, warmup=3)
        torch.cuda.empty_cache()
        if rank == 0:
            eager = bench_gpu_eager(dev)
        D.barrier()

    if rank != 0:
        D.close()
        return
    pk = peaks()
    gm = prof["gemm"]
    M = B * 496
    tf_alg = gemm_flops_algorithmic(M) / (gm["ms_per_step"] * 1e-3) / 1e12 if gm["ms_per_step"] > 0 else 0.0
    tf_exec = gm["work_per_step"] / (gm["ms_per_step"] * 1e-3) / 1e12 if gm["ms_per_step"] > 0 else 0.0
    roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": tf_alg, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": tf_alg / pk["tf_sustained"], "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
            "algorithmic_flops_per_step": gemm_flops_algorithmic(M), "executed_tflops": tf_exec,
            "executed_note": "executed also counts the 3x K of the 3-term split-bf16 patch-embed / projection GEMMs",
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
    kernels["fbank"]["algorithmic_bytes_per_clip"] = 4 * T + 4 * 998 * 128
    for k in ("attention", "posconv"):
        v = prof[k]
        kernels[k]["TFLOPs"] = v["work_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12 if v["ms_per_step"] > 0 else 0.0

    # ---- CPU baseline on this box's host cores (bounded sample) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        arm = CpuArm()
        n_cpu = 32
        arm.run(8)  # warm-up
        sec = arm.run(n_cpu)
        cores = os.cpu_count() or 1
        cpu = {"value": n_cpu * CLIP_SECONDS / 3600.0 / sec, "unit": "audio-hours/s", "cores": cores, "kind": arm.kind,
               "sample": f"{n_cpu} x 10 s clips, one pass, {arm.what}, {cores} threads, {sec:.2f} s"}  # fmt: skip

    line = {
        "metric": "beats_embed_throughput", "value": value, "unit": "audio-hours/s", "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32 (3-term split-bf16 GEMMs, fp32 attention / pos-conv)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "batch_per_gpu": B, "global_batch": world * B, "clip_seconds": CLIP_SECONDS, "tokens_per_clip": 496,
                   "weights": "random-init, reference init distributions, seed 0", "parallelism": f"clip-sharded dp{world}",
                   "l2": "inputs and activations (0.16-3.9 GB per step) exceed the 126 MB L2; no flush needed",
                   "clips_per_s": clips_per_s, "model_tflops": clips_per_s * FLOPS_PER_CLIP_10S / 1e12},
        "e2e": {"value": e2e_value, "unit": "audio-hours/s", "h2d_bytes_per_step": world * B * T * 4, "d2h_bytes_per_step": world * B * 768 * 4,
                "ms_per_step": ms_e2e / K_,
                "api": "plugin Model.register_hooks_for_layers(['last_layer']) + Model.extract_embeddings(x, aggregation='mean') (the reference's "
                       "public call), pinned host input, double-buffered H2D, pooled [B,768] read back"},
        "gpu_launches": launches,
        "roofline": roof,
        "kernels": kernels,
        "encoder_tensor_frac": clips_per_s * FLOPS_PER_CLIP_10S / 1e12 / world / pk["tf_sustained"],
        "strong": strong,
        "secondary": secondary,
        "gpu_eager_baseline": eager,
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    D.close()


# ------------------------------------------------------------------------------------------------------------
# secondary: BEATs at the configs[4] shape (64 x 60 s clips, layer-wise features: all 13 hooked layers, mean-pooled)
# ------------------------------------------------------------------------------------------------------------
def bench_beats_long(D: Dist, model, steps: int = 3):
    import torch

    world, dev = D.world, D.dev
    Bl = max(1, 64 // world)
    T = 60 * SAMPLE_RATE
    g = torch.Generator(device=dev).manual_seed(99 + D.rank)
    wav = torch.randn(Bl, T, device=dev, generator=g) * 0.1
    model.register_hooks_for_layers(["all"])
    try:
        def step(_i=0):
            return model.extract_embeddings(wav, aggregation="mean")  # [Bl, 13 * 768]

        for _ in range(2):
            emb = step()
        assert emb.shape == (Bl, 13 * 768)
        ms = D.timed(step, steps) / steps
        # attention share from the event profiler
        from avex_b200 import _lib

        lib = _lib.load()
        lib.avexk_profile_enable(1)
        step()
        torch.cuda.synchronize()
        n, t, w = C.c_longlong(), C.c_double(), C.c_double()
        lib.avexk_profile_read(2, C.byref(n), C.byref(t), C.byref(w))
        att_ms, att_tf = t.value, (w.value / (t.value * 1e-3) / 1e12 if t.value > 0 else 0.0)
        lib.avexk_profile_enable(0)
    finally:
        model.deregister_all_hooks()
    clips_per_s = world * Bl / (ms * 1e-3)
    pk = peaks()
    return {"workload": "BEATs-base, 64 x 60 s clips (N = 2992 tokens), all 13 hooked layers mean-pooled -> [B, 9984] (BASELINE configs[4] shape; "
                        "strong scaling: 64 / n_gpus clips per rank)",
            "clips_per_rank": Bl, "ms_per_step": ms, "steps": steps, "value": clips_per_s * 60 / 3600.0, "unit": "audio-hours/s",
            "model_tflops": clips_per_s * FLOPS_PER_CLIP_60S / 1e12, "encoder_tensor_frac": clips_per_s * FLOPS_PER_CLIP_60S / 1e12 / world / pk["tf_sustained"],
            "attention_ms_per_step": att_ms, "attention_share": att_ms / ms, "attention_TFLOPs": att_tf}  # fmt: skip


# ------------------------------------------------------------------------------------------------------------
# secondary workload: EfficientNet-B0 mel-spectrogram feature extractor (BASELINE configs[2])
# ------------------------------------------------------------------------------------------------------------
EFF_CLIP_SECONDS = 5
EFF_ACT_BYTES_PER_CLIP = 35.0e6      # ~35 MB of bf16 activation traffic per 5 s clip (unfused NHWC path)
EFF_MEL_BYTES_PER_CLIP = 4 * 80000 + 4 * 128 * 501  # waveform in, single-channel log-mel out


def bench_effnet(D: Dist, steps: int, warmup: int, batch: int | None = None, cpu_baseline: bool = True):
    """5 s clips -> mel -> EfficientNet-B0 features [B,1280,4,16] through the plugin Model; returns the result dict."""
    import numpy as np
    import torch

    from avex_b200 import _lib, plugin
    from avex_b200.plugin import efficientnet_model  # noqa: F401
    from oracle.weights import make_effnet_weights  # seeded weight generation only (outside every timed region)

    world, rank, dev = D.world, D.rank, D.dev
    spec = plugin.ModelSpec(name="efficientnet", device="cuda", efficientnet_variant="b0",
                            audio_config=dict(sample_rate=16000, n_fft=800, hop_length=160, win_length=800, window="hann", n_mels=128,
                                              representation="mel_spectrogram", normalize=True, target_length_seconds=10,
                                              window_selection="random"))  # fmt: skip
    plugin.register_model("bench_effnet_b0", spec)
    model = plugin.build_model_from_spec(spec, "cuda", pretrained=False, return_features_only=True).eval()
    stats_path = os.path.join(ROOT, "tests", "golden", "effnet_bn_stats.npz")
    W = make_effnet_weights(seed=3, num_classes=0, bn_stats=dict(np.load(stats_path)))  # calibrated BatchNorm statistics
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=False)
    B = batch if batch is not None else max(1, 512 // world)
    T = EFF_CLIP_SECONDS * SAMPLE_RATE
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    wav_dev = torch.randn(B, T, device=dev, generator=g) * 0.1
    host = [torch.empty(B, T, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in host:
        h.copy_(wav_dev.cpu())
    feat_host = torch.empty(B, 1280, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [torch.empty(B, T, device=dev) for _ in range(2)]
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def step_device(_i=0):
        with torch.no_grad():
            return model(wav_dev)

    def step_e2e(i=0):
        """pinned host batch -> H2D (copy stream, double-buffered as in the BEATs line) -> Model.forward -> pooled features D2H."""
        cur = i & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):  # the next batch's H2D is enqueued before this batch's kernels: it overlaps them
            copy_stream.wait_event(ev_consumed[cur ^ 1])
            dev_in[cur ^ 1].copy_(host[cur ^ 1], non_blocking=True)
            ev_ready[cur ^ 1].record(copy_stream)
        main.wait_event(ev_ready[cur])
        with torch.no_grad():
            f = model(dev_in[cur])
        ev_consumed[cur].record(main)
        feat_host.copy_(f.mean(dim=(2, 3)), non_blocking=True)

    for _ in range(max(3, warmup)):
        step_device()
    l0 = _lib.launch_count()
    ms = D.timed(step_device, steps) / steps
    launches = _lib.launch_count() - l0
    with torch.cuda.stream(copy_stream):  # arm buffer 0 for the timed loop
        dev_in[0].copy_(host[0], non_blocking=True)
        ev_ready[0].record(copy_stream)
    ev_consumed[1].record(torch.cuda.current_stream())
    ms_e2e = D.timed(step_e2e, steps) / steps
    mel = model._engine.mel
    ms_mel = D.timed(lambda i: mel.run(wav_dev, normalize=True), steps) / steps
    pk = peaks()
    clips_per_s = world * B / (ms * 1e-3)
    act_gbs = B * EFF_ACT_BYTES_PER_CLIP / (ms * 1e-3) / 1e9
    mel_gbs = B * EFF_MEL_BYTES_PER_CLIP / (ms_mel * 1e-3) / 1e9
    cpu = eager = None
    del model
    torch.cuda.empty_cache()
    if cpu_baseline and rank == 0 and world == 1:
        cpu = effnet_cpu_baseline()
        eager = effnet_gpu_eager(dev)
    return {
        "workload": "EfficientNet-B0 mel-spectrogram feature extractor, 512 x 5 s clips @16 kHz (BASELINE configs[2]; 512 / n_gpus clips per rank), features [B,1280,4,16]",
        "value": clips_per_s, "unit": "clips/s", "audio_hours_per_s": clips_per_s * EFF_CLIP_SECONDS / 3600.0, "ms_per_step": ms, "steps": steps,
        "clips_per_rank": B, "gpu_launches": launches,
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": world * B * T * 4, "d2h_bytes_per_step": world * B * 1280 * 4,
                "ms_per_step": ms_e2e, "api": "plugin Model.forward; every step copies its batch from pinned host memory (copy stream, double-buffered) and reads the pooled [B,1280] features back"},
        "roofline": {"bound": "hbm", "kernel": "whole forward (NHWC fp16 activations)", "achieved": act_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": act_gbs / pk["hbm_gbs"], "peak_source": f"{pk['src']} hbm_gbs", "traffic": None,
                     "algorithmic_bytes": "35 MB of bf16 activation traffic per 5 s clip (SURVEY 8d)"},
        "kernels": {"melspec": {"ms_per_step": round(ms_mel, 4), "achieved_GBps": mel_gbs, "hbm_frac": mel_gbs / pk["hbm_gbs"],
                                "algorithmic_bytes_per_clip": EFF_MEL_BYTES_PER_CLIP}},
        "cpu_baseline": cpu, "gpu_eager_baseline": eager,
    }  # fmt: skip


def effnet_cpu_baseline(n_clips: int = 32):
    """The reference's own EfficientNet path (avex audio processor + torchvision efficientnet_b0) on the host cores."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    avex = import_reference()
    wav = torch.randn(n_clips, EFF_CLIP_SECONDS * SAMPLE_RATE, generator=torch.Generator().manual_seed(4321)) * 0.1
    try:
        if avex is None:
            raise RuntimeError("baseline/_ref absent")
        from avex.models.utils.factory import build_model_from_spec
        from avex.models.utils.registry import get_model_spec

        model = build_model_from_spec(get_model_spec("esp_aves2_effnetb0_all"), "cpu", pretrained=False, return_features_only=True).eval()
        kind, what = "reference", "unmodified avex efficientnet Model (AudioProcessor mel + torchvision efficientnet_b0), random init"

        def run(x):
            with torch.no_grad():
                return model(x)
    except Exception as e:
        return {"value": None, "unavailable": f"{type(e).__name__}: {e}"}
    run(wav[:4])
    t0 = time.perf_counter()
    for i in range(0, n_clips, 8):
        run(wav[i : i + 8])
    sec = time.perf_counter() - t0
    return {"value": n_clips / sec, "unit": "clips/s", "cores": cores, "kind": kind, "sample": f"{n_clips} x 5 s clips, one pass, {what}, {cores} threads, {sec:.2f} s"}


def effnet_gpu_eager(dev, batch: int = 128, steps: int = 3):
    """The incumbent on the same GPU for the EfficientNet line: the unmodified reference Model (torchaudio / torchvision ops in
    eager mode) on cuda, fp32 and under bf16 autocast."""
    import torch

    try:
        avex = import_reference()
        if avex is None:
            raise RuntimeError("baseline/_ref absent")
        from avex.models.utils.factory import build_model_from_spec
        from avex.models.utils.registry import get_model_spec

        model = build_model_from_spec(get_model_spec("esp_aves2_effnetb0_all"), "cuda", pretrained=False, return_features_only=True).eval()
        wav = torch.randn(batch, EFF_CLIP_SECONDS * SAMPLE_RATE, device=dev, generator=torch.Generator(device=dev).manual_seed(4321)) * 0.1

        def timeit():
            with torch.no_grad():
                for _ in range(2):
                    model(wav)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(steps):
                    model(wav)
                b.record()
                torch.cuda.synchronize()
            return a.elapsed_time(b) / steps

        ms32 = timeit()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms16 = timeit()
        out = {"kind": "reference", "what": "unmodified avex efficientnet Model on cuda (AudioProcessor mel + torchvision efficientnet_b0, eager), random init",
               "batch": batch, "fp32": {"ms_per_batch": ms32, "value": batch / (ms32 * 1e-3)},
               "autocast_bf16": {"ms_per_batch": ms16, "value": batch / (ms16 * 1e-3)}, "unit": "clips/s"}
    except Exception as e:  # an OOM or an import problem must not cost the line
        out = {"unavailable": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------
# the incumbent on the same GPU: the reference's own torch modules in eager mode
# ------------------------------------------------------------------------------------------------------------
def bench_gpu_eager(dev, batch: int = 32, steps: int = 3):
    import torch

    avex = import_reference()
    wav = torch.randn(batch, CLIP_SECONDS * SAMPLE_RATE, device=dev, generator=torch.Generator(device=dev).manual_seed(1234)) * 0.1
    try:
        if avex is not None:
            model = reference_model(avex, "cuda")
            kind = "reference"
            what = "unmodified avex beats Model on cuda, extract_embeddings(aggregation='mean'), hook on the last layer"

            def run():
                with torch.no_grad():
                    return model.extract_embeddings(wav, aggregation="mean")
        else:
            from oracle import beats_encoder as OE
            from oracle import beats_torch as OT
            from oracle.weights import make_beats_weights

            dims = OE.BeatsDims(layers=12)
            Wt = {k: v.to(dev) for k, v in OT.to_torch(make_beats_weights(dims, seed=0, init="reference")).items()}
            kind, what = "port", "oracle/beats_torch.py (the reference's op mix in stock torch) on cuda"

            def run():
                return OT.beats_forward(Wt, wav, None, dims)["x"].mean(dim=1)

        def timeit():
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                run()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / steps

        ms32 = timeit()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms16 = timeit()
        ah = lambda ms: batch * CLIP_SECONDS / 3600.0 / (ms * 1e-3)  # noqa: E731
        out = {"kind": kind, "what": what, "batch": batch, "note": "the materialised [B,12,N,N] mask limits the eager batch; per-clip cost is flat in B",
               "fp32": {"ms_per_batch": ms32, "value": ah(ms32)}, "autocast_bf16": {"ms_per_batch": ms16, "value": ah(ms16)}, "unit": "audio-hours/s"}
    except Exception as e:  # an OOM or an import problem must not cost the main line
        out = {"unavailable": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    return out


def run_effnet(args):
    """`--workload effnet`: the EfficientNet line on its own (same dict as secondary.effnet_512x5s, promoted to a full line)."""
    D = Dist()
    sampler = ClockSampler(D.local)
    if D.rank == 0:
        sampler.start()
    r = bench_effnet(D, steps=args.steps, warmup=args.warmup, batch=None if args.batch == BATCH_PER_GPU else args.batch,
                     cpu_baseline=not args.no_cpu_baseline)
    if D.rank == 0:
        sampler.stop()
        line = {"metric": "effnet_b0_feature_throughput", "value": r["value"], "unit": "clips/s", "n_gpus": D.world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if D.world > 1 else "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": r["workload"], "batch_per_gpu": r["clips_per_rank"], "weights": "random-init, calibrated BatchNorm statistics",
                           "l2": "activations (up to 1.6 GB per layer) exceed the 126 MB L2; no flush needed"},
                "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": r["roofline"], "kernels": r["kernels"],
                "cpu_baseline": r["cpu_baseline"], "gpu_eager_baseline": r["gpu_eager_baseline"], "clocks": sampler.summary()}  # fmt: skip
        print(json.dumps(line), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong / secondary / gpu_eager_baseline (kernel iteration runs)")
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

#!/usr/bin/env python
"""bench.py -- headline benchmark of the RemFx hot path on B200 (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--legs 3,4,5|none]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
  Open-Unmix distortion-removal forward = `OpenUnmixModel.sample`, batch 32 x 262144 samples, 48 kHz mono,
  synthetic audio (clamp(0.1 N(0,1))) and seeded random-init weights (no network for data / checkpoints).
A "step" is one pass of the hot path over one batch.  Metric: audio-seconds per second (whole job).

  value         device-timed (CUDA events on the launching stream), inputs already resident in HBM
  e2e           same call through the public API on pinned HOST buffers (H2D + kernels + D2H inside the timed region)
  roofline      dominant kernel (BiLSTM recurrence): latency-bound; algorithmic bytes / live per-launch duration vs
                MEASURED_PEAKS.json, us per dependent step, plus the whole step against both rooflines
  cpu_baseline  the UNCHANGED reference `remfx.models.OpenUnmixModel.sample` (oracle/_ref, see oracle/make_ref.py) on the
                host cores, full batch (rank 0, N=1); falls back to the oracle port when no reference copy travelled
  gpu_eager     the same unchanged reference module `.cuda()` on this GPU under stock PyTorch (cuDNN-RNN / cuBLAS / cuFFT),
                TF32 off and on -- SURVEY 8(d)(ii)'s "real bar"
  other_configs BASELINE.json configs 3, 4, 5 measured in the same run (each with its own roofline): Hybrid-Demucs
                forward + MR-STFT/L1 loss 32x262144; RemFX-detect chain 16x262144 (global batch, sharded by item; Hybrid
                Demucs for distortion / compressor as cfg/exp/remfx_detect.yaml:63-68 ships, Open-Unmix for the rest);
                data-parallel training step (forward + loss + backward + ONE NCCL all-reduce of the flat gradient
                bucket + clip + AdamW) with the all-reduce timed on the device.

`--impl reference` times the reference's own CPU implementation of the headline path (the unchanged module from
oracle/_ref) on the host cores, full batch per step, and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
T = 262144
BATCH = 32
CHUNK_S = T / SR
METRIC = "audio-seconds/sec @48kHz 262144-sample chunks"
WORKLOAD = "Open-Unmix (umx) distortion-removal forward (OpenUnmixModel.sample), batch 32x262144, fp32-parity mode"
NBUF = 5  # distinct input batches rotated through the timed loop: 5 x 33.5 MB = 168 MB > 126 MB L2


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def _reference_module(device="cpu"):
    """The unchanged reference wrapper (oracle/_ref or /root/reference) in eval mode, or None when no copy is available."""
    import torch

    from oracle import refshim

    if not refshim.available():
        return None
    mods = refshim.ref_modules()
    torch.manual_seed(0)
    m = mods.models.OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=SR)
    return m.to(device).eval()


def _cpu_reference(steps: int, warmup: int, items: int):
    """Reference CPU path on all host threads: (times, cores, threads, kind, what)."""
    import torch

    torch.set_flush_denormal(True)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from remfx_b200.synth import synth_audio

    x = synth_audio(12345, items, T)
    ref = _reference_module("cpu")
    if ref is not None:
        kind, what = "reference", "unchanged remfx.models.OpenUnmixModel.sample (oracle/_ref copy of the reference, torch-CPU)"
        fn = lambda: ref.sample(x)  # noqa: E731
    else:
        from oracle import umx as oumx
        from oracle import weights

        sd = weights.umx_state(0)
        kind, what = "port", "oracle port of OpenUnmixModel.sample (torch-CPU, fused LSTM); no oracle/_ref copy travelled"
        fn = lambda: oumx.sample(x, sd, fast_lstm=True, wiener_trig=True)  # noqa: E731
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, cores, torch.get_num_threads(), kind, what


def run_reference(args):
    rank, _, world = _dist_env()
    if rank != 0:
        return
    # the whole batch-32 workload per step (about a second of CPU work per step); only a very long --steps run falls back to a
    # bounded sample of the batch so that the arm still ends within a few minutes
    items = BATCH if args.steps <= 60 else max(1, min(BATCH, (60 * BATCH) // args.steps))
    times, cores, threads, kind, what = _cpu_reference(args.steps, max(1, args.warmup), items)
    tot = sum(times)
    value = items * CHUNK_S * len(times) / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU", "global_batch": items, "chunk_samples": T, "sample_rate": SR,
                   "step": f"{items} of the {BATCH} chunks per step" + ("" if items == BATCH else " (bounded sample: long --steps run)")},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": threads, "kind": kind,
                         "sample": f"{what}, {items}x262144 per step, {len(times)} steps, os.cpu_count()={cores}"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _gpu_eager(dev, x):
    """the unchanged reference module on this GPU under stock PyTorch eager, TF32 off / on."""
    import torch

    ref = _reference_module(dev)
    if ref is None:
        return {"unavailable": "no oracle/_ref copy of the reference travelled with this snapshot"}
    out = {"api": "unchanged remfx.models.OpenUnmixModel.sample, .cuda().eval(), torch.no_grad (cuDNN RNN, cuBLAS, cuFFT, "
                  "Python Wiener loop)", "batch": int(x.shape[0])}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.no_grad():
            for name, tf32 in (("tf32_off", False), ("tf32_on", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                for _ in range(2):
                    ref.sample(x)
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ref.sample(x)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = statistics.median(ts)
                out[name] = {"ms_per_step": ms, "ms_best": min(ts), "value": x.shape[0] * CHUNK_S / (ms / 1e3), "unit": "audio-s/s"}
    except Exception as exc:  # the bar must never take the product line down with it
        out["error"] = f"{type(exc).__name__}: {exc}"[:300]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


# ======================================================================================================
# BASELINE.json configs 3, 4, 5 in the same run
# ======================================================================================================
def _time_calls(fn, n, warm, barrier, reduce_max):
    import torch

    for _ in range(warm):
        fn()
    barrier()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    barrier()
    return reduce_max(statistics.median(ts)), reduce_max(min(ts))


def leg_demucs_forward(dev, rank, world, barrier, reduce_max, peaks):
    """Config 3: DemucsModel.forward((x, y)) -> (loss, out), 32 x 262144 per GPU (weak scaling), no collective."""
    import torch

    from remfx_b200.models import DemucsModel
    from remfx_b200.synth import synth_audio

    B = 32
    torch.manual_seed(0)
    m = DemucsModel(sample_rate=SR, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).to(dev).eval()
    x, y = synth_audio(31 + rank, B, T).to(dev), synth_audio(32 + rank, B, T).to(dev)
    with torch.no_grad():
        med, best = _time_calls(lambda: m((x, y)), 5, 2, barrier, reduce_max)
    flops = 117.0e9 * B   # 110.62 conv/linear/bmm + 6.40 LSTM GFLOP per chunk
    ach = flops / (med / 1e3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    out = {"config": "3: Hybrid-Demucs forward + MR-STFT/L1 loss, batch 32x262144 per GPU", "api": "DemucsModel.forward((x, target))",
           "ms_per_step": med, "ms_best": best, "value": world * B * CHUNK_S / (med / 1e3), "unit": "audio-s/s", "scaling": "weak",
           "launches": m.launches_per_call(B, T),
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "algorithmic_flops": flops,
                        "note": "whole forward, algorithmic fp32 FLOPs; the bf16x3 split issues 3x as many MMA FLOPs (frac x3 of the "
                                "pipe is busy with them); sustained cuBLAS bf16 peak"}}
    del m
    torch.cuda.empty_cache()
    return out


def leg_chain(dev, rank, world, barrier, reduce_max, peaks):
    """Config 4: classifier + cascade, GLOBAL batch 16 x 262144 sharded by item over the ranks (strong scaling), no collective.
    Members as cfg/exp/remfx_detect.yaml ships where an oracle exists: Hybrid Demucs for distortion and compressor (:63-68);
    its DCUNet members (reverb / chorus / delay, asteroid: no oracle here) are stood in for by Open-Unmix."""
    import torch

    from remfx_b200.chain import ALL_EFFECTS, RemFXChainInference
    from remfx_b200.classifier import Cnn14
    from remfx_b200.models import DemucsModel, OpenUnmixModel
    from remfx_b200.parallel import shard_range
    from remfx_b200.synth import synth_audio

    Bg = 16
    lo, hi = shard_range(Bg, rank, world)
    order = ["RandomPedalboardDistortion", "RandomPedalboardCompressor", "RandomPedalboardReverb", "RandomPedalboardChorus",
             "RandomPedalboardDelay"]  # cfg/exp/remfx_detect.yaml:80-85
    members = {}
    for i, e in enumerate(ALL_EFFECTS):
        torch.manual_seed(100 + i)
        if e in ("RandomPedalboardDistortion", "RandomPedalboardCompressor"):
            members[e] = DemucsModel(sample_rate=SR, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).to(dev).eval()
        else:
            members[e] = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=SR).to(dev).eval()
    torch.manual_seed(7)
    clf = Cnn14(num_classes=5, sample_rate=SR, model_sample_rate=SR, n_fft=2048, hop_length=512, n_mels=128, specaugment=True).to(dev).eval()
    xg, yg = synth_audio(41, Bg, T), synth_audio(42, Bg, T)
    x, y = xg[lo:hi].to(dev), yg[lo:hi].to(dev)
    res = {}
    for name, use_all in (("all_effect_models", True), ("classifier_decisions", False)):
        chain = RemFXChainInference(members, SR, 1025, order, classifier=clf, use_all_effect_models=use_all)
        fn = (lambda: chain((x, y, None, None), 0)) if hi > lo else (lambda: None)
        med, best = _time_calls(fn, 5, 2, barrier, reduce_max)
        res[name] = {"ms_per_step": med, "ms_best": best, "value": Bg * CHUNK_S / (med / 1e3), "unit": "audio-s/s"}
        if not use_all and hi > lo:
            res[name]["effects_per_item_rank0"] = float(chain.last_labels.sum(1).mean())
    # algorithmic FLOPs of the all-effects pass per item: Cnn14 41.3 + 2 x HDemucs 117.0 + 3 x Open-Unmix 6.57 GFLOP
    flops = (41.3e9 + 2 * 117.0e9 + 3 * 6.57e9) * Bg
    med = res["all_effect_models"]["ms_per_step"]
    ach = flops / (med / 1e3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) * world
    out = {"config": "4: RemFX-detect chain (Cnn14 + Hybrid-Demucs x2 [distortion, compressor] + Open-Unmix x3), global batch 16x262144",
           "api": "RemFXChainInference.forward((x, y, None, None))", "scaling": "strong", "items_per_gpu": hi - lo,
           "ms_per_step": med, "value": res["all_effect_models"]["value"], "unit": "audio-s/s", "modes": res,
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "algorithmic_flops": flops, "note": "all-effects pass, whole job over n_gpus; see config 3 for the bf16x3 caveat"}}
    del members, clf
    torch.cuda.empty_cache()
    return out


def leg_train(dev, rank, world, barrier, reduce_max, peaks, use_dist):
    """Config 5: data-parallel training step = forward (kept activations) + MR-STFT/100 L1 + backward + ONE all-reduce of the
    flat fp32 gradient bucket (NCCL over NVLink, inside FusedAdamW.step) + clip-by-global-norm 10 + AdamW.  Weak scaling."""
    import torch

    from remfx_b200.models import DemucsModel, TCNModel
    from remfx_b200.synth import synth_audio
    from remfx_b200.train import RemFX

    torch.manual_seed(0)
    if getattr(DemucsModel, "supports_training", False):
        net = DemucsModel(sample_rate=SR, sources=["mixture"], audio_channels=1, nfft=4096, channels=48)
        B, name, fwd_flops = 16, "Hybrid Demucs (cfg/exp/5-5_full.yaml:3), 16x262144 per GPU (cfg/exp/5-5_full.yaml:27)", 117.0e9
    else:
        net = TCNModel(sample_rate=SR, num_bins=1025, ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7,
                       stack_size=10, dilation_growth=2, condition=False, latent_dim=2, norm_type="identity", causal=False,
                       estimate_loudness=False)
        B, name, fwd_flops = 1, "TCN (cfg/model/tcn.yaml), 1x262144 per GPU", 5135.5e9
    mod = RemFX(lr=1e-4, lr_beta1=0.95, lr_beta2=0.999, lr_eps=1e-6, lr_weight_decay=1e-3, sample_rate=SR, network=net.to(dev), max_steps=50000)
    mod.compute_metrics = False  # the no_grad metric block (remfx/models.py:227-255) is reported separately below
    x, y = synth_audio(51 + rank, B, T).to(dev), synth_audio(52 + rank, B, T).to(dev)
    batch = (x, y, None, None)
    nparam = sum(p.numel() for p in net.parameters())
    for i in range(2):
        mod.fit_step(batch, i)
    opt = mod._optim
    opt.set_timing(True)
    barrier()
    torch.cuda.reset_peak_memory_stats(dev)
    ts, n = [], 4
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mod.fit_step(batch, i)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    barrier()
    med = reduce_max(statistics.median(ts))
    tm = opt.timing_ms()
    ar_local = statistics.median([a for a, _ in tm])
    ar = reduce_max(ar_local)        # the rank that waited longest: collective + the skew of the ranks' backward passes
    ar_min = -reduce_max(-ar_local)  # the last rank to arrive: the collective itself
    upd = reduce_max(statistics.median([b for _, b in tm]))
    opt.set_timing(False)
    mod.compute_metrics = True
    tsm = []
    for i in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mod.fit_step(batch, i)
        e1.record()
        torch.cuda.synchronize()
        tsm.append(e0.elapsed_time(e1))
    with_metrics = reduce_max(min(tsm))
    grad_bytes = nparam * 4
    ideal_ar_ms = (2.0 * (world - 1) / world * grad_bytes / 900e9 * 1e3) if world > 1 else 0.0
    flops = 3 * fwd_flops * B  # forward + input gradient + weight gradient
    ach = flops / (med / 1e3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    out = {"config": f"5: data-parallel training step, {name}", "api": "remfx_b200.train.RemFX.fit_step (what Lightning runs around training_step)",
           "scaling": "weak", "ms_per_step": med, "value": world * B * CHUNK_S / (med / 1e3), "unit": "audio-s/s",
           "ms_per_step_with_metric_block": with_metrics,
           "all_reduce": {"collective": "one NCCL SUM all-reduce of the flat fp32 gradient bucket" if use_dist else "none (1 GPU)",
                          "bytes": grad_bytes, "ms": ar, "ms_last_rank_to_arrive": ar_min, "ideal_ring_ms_at_900GBs": ideal_ar_ms,
                          "busbw_GBs": (2.0 * (world - 1) / world * grad_bytes / (ar_min / 1e3) / 1e9) if (use_dist and ar_min > 0) else None,
                          "note": "ms = max over ranks of the time between the all-reduce's launch and its end on that rank's stream: it contains "
                                  "the wait for the slowest rank's backward (rank skew); ms_last_rank_to_arrive = min over ranks = the collective "
                                  "itself, which is what busbw is computed from",
                          "overlap_with_backward": False},
           "clip_adamw_ms": upd, "parameters": nparam, "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30,
           "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "algorithmic_flops": flops, "note": "3 x forward FLOPs per chunk (forward, input gradient, weight gradient); per GPU"}}
    del mod, net, opt
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch

    from remfx_b200 import _lib
    from remfx_b200.models import OpenUnmixModel
    from remfx_b200.synth import synth_audio

    rank, local, world = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist

        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    _lib.lib()  # fail loudly if the CUDA library is missing

    torch.manual_seed(0)
    model = OpenUnmixModel(n_fft=2048, hop_length=512, n_channels=1, alpha=0.3, sample_rate=SR)  # random-init weights
    model = model.to(dev).eval()
    xs_host = [synth_audio(12345 + rank * 100 + i, BATCH, T).pin_memory() for i in range(NBUF)]
    xs = [x.to(dev) for x in xs_host]
    out_host = torch.empty(BATCH, 1, T, dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()

    def reduce_max(v: float) -> float:
        if not use_dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: the multi-lane pipeline (model.pipeline(): rfx_umx_pipe_*) ------------
    # Inputs already in HBM.  K steps are pushed back to back and the pipeline is flushed inside the timed region, so the
    # number includes the fill and drain of the pipeline (one step's latency, ~5 ms, against ~1.35 ms per step in steady state).  CUDA events on the launching (current) stream: the first
    # lane waits for everything enqueued before the push, and flush() makes the current stream wait for every output.
    pipe = model.pipeline(dev)
    depth = pipe.depth
    nout = 2 * depth + 2
    outs_dev = [torch.empty(BATCH, 1, T, dtype=torch.float32, device=dev) for _ in range(nout)]
    nwarm = max(3, args.warmup)
    for i in range(nwarm):
        pipe.push(xs[i % NBUF], outs_dev[i % nout])
    pipe.flush()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pipe.set_profiling(min(4096, args.steps * model.model.nb_layers))
    barrier()
    ev0.record()
    for k in range(args.steps):
        pipe.push(xs[k % NBUF], outs_dev[k % nout])
    pipe.flush()
    ev1.record()
    barrier()
    ms_total = reduce_max(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = world * BATCH * CHUNK_S / (ms_step / 1e3)
    lstm_ms = pipe.recurrence_times_ms()   # every recurrence launch of the timed region (cudaEvents on its own stream)
    pipe.set_profiling(0)

    # ---- bf16-fast mode (BASELINE.json configs[1] says "bf16"): single-pass products in gemm2 and the tcgen05 recurrence ------
    # Same pipeline, same timed region as `value`; reported beside the fp32-parity headline with its measured error against it.
    import remfx_b200

    bf16_fast = None
    try:
        ref_out = outs_dev[(args.steps - 1) % nout].clone()          # parity-mode output of the last timed step
        ref_in = xs[(args.steps - 1) % NBUF]
        remfx_b200.set_precision("bf16")
        for i in range(nwarm):
            pipe.push(xs[i % NBUF], outs_dev[i % nout])
        pipe.flush()
        barrier()
        fv0, fv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fv0.record()
        for k in range(args.steps):
            pipe.push(xs[k % NBUF], outs_dev[k % nout])
        pipe.flush()
        fv1.record()
        barrier()
        fast_ms = reduce_max(fv0.elapsed_time(fv1)) / args.steps
        seq = pipe.push(ref_in, outs_dev[0])
        pipe.flush()
        fast_out = pipe.wait(seq)
        torch.cuda.synchronize()
        err = float((fast_out.double() - ref_out.double()).norm() / ref_out.double().norm())
        bf16_fast = {"value": world * BATCH * CHUNK_S / (fast_ms / 1e3), "unit": "audio-s/s", "ms_per_step": fast_ms,
                     "rel_rms_vs_fp32_parity_mode": err,
                     "what": "remfx_b200.set_precision('bf16'): hi*hi pass only in gemm2 and the tcgen05 recurrence (3x fewer MMAs); "
                             "activations still travel as hi/lo planes"}
    except Exception as exc:
        bf16_fast = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    finally:
        remfx_b200.set_precision("fp32")
        barrier()

    # ---- the same workload as one blocking call per step (latency form), with per-stage timing -----------
    model.set_profiling(True, dev)
    for i in range(3):
        model.sample(xs[i % NBUF])
    barrier()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nser = min(args.steps, 10)
    sv0.record()
    for k in range(nser):
        model.sample(xs[k % NBUF])
    sv1.record()
    barrier()
    serial_ms = reduce_max(sv0.elapsed_time(sv1)) / nser
    stage_acc = model.stage_times_ms()  # stages of the last serial step
    model.set_profiling(False, dev)

    # ---- end-to-end through the public API on host buffers ---------------------------------------------
    # Headline e2e = the pipeline on pinned HOST tensors: every step does its own H2D (33.5 MB) and D2H (33.5 MB) inside
    # the timed region and the loop consumes every result (wait() on step k - 6 before pushing k, as a serving loop
    # would); host wall clock, stopped when the last result is in host memory.  Beside it: one blocking call per step.
    outs_host = [torch.empty(BATCH, 1, T, dtype=torch.float32).pin_memory() for _ in range(nout)]
    out_host = outs_host[0]
    for i in range(2):
        model.sample_host(xs_host[i % NBUF], out_host, dev)
    barrier()
    t0 = time.perf_counter()
    for k in range(nser):
        model.sample_host(xs_host[k % NBUF], out_host, dev)
    torch.cuda.synchronize()
    sync_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    sync_ms = reduce_max(sync_ms) / nser

    for i in range(nwarm):
        pipe.push(xs_host[i % NBUF], outs_host[i % nout])
    pipe.flush()
    barrier()
    seqs = []
    lag = 2 * depth   # results are consumed `lag` steps behind the newest push (lag + 1 host buffers in flight <= nout)
    t0 = time.perf_counter()
    for k in range(args.steps):
        if k >= lag:
            pipe.wait(seqs[k - lag])
        seqs.append(pipe.push(xs_host[k % NBUF], outs_host[k % nout]))
    pipe.flush()
    for sq in seqs[-lag:]:
        pipe.wait(sq)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)   # host wall clock: every result is in host memory when it stops
    barrier()
    e2e_ms = reduce_max(e2e_ms) / args.steps
    clocks = sampler.stop()
    e2e_value = world * BATCH * CHUNK_S / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel (BiLSTM recurrence, one launch = one layer, both directions) ----
    peaks, peak_src = _peaks()
    F = T // 512 + 1
    M = BATCH * F
    H = 256
    lstm_bytes = M * 8 * H * 4 + M * 2 * H * 4 + 2 * 4 * H * H * 4  # read G, write h, read W_hh once
    lstm_t = statistics.mean(lstm_ms) / 1e3   # average over the recurrence launches of the timed region
    achieved = lstm_bytes / lstm_t / 1e9
    pinfo = pipe.info()
    steps_per_launch = F  # 513 strictly dependent time steps per launch (one launch = one BiLSTM layer, both directions)
    step_bytes = 92.5e6    # algorithmic HBM bytes of one whole step (x in, out, weights once, nothing in between)
    step_flops = 6.57e9 * BATCH  # FC 1.61 + LSTM 4.84 + FFT 0.12 GFLOP per chunk
    tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    roofline = {
        "kernel": "lstm_rec_tc_kernel<16,2> (BiLSTM recurrence on tcgen05, W_hh as the TMEM A operand, two phase-locked 16-slot groups per CTA; 1 launch per layer)",
        "bound": "latency",
        "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": 157.8e6,  # dram read+write bytes per launch, ncu --set full (profiles/r2/ncu_lstm_rec_tc_dual_locked.txt)
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": lstm_bytes, "ms_per_launch": lstm_t * 1e3,
        "launches_timed": len(lstm_ms),
        "us_per_dependent_step": lstm_t * 1e6 / steps_per_launch, "dependent_steps_per_launch": steps_per_launch,
        "tensor_pipe_active_frac_ncu": 0.28,  # sm__pipe_tensor_subunit active on the SMs the launch holds (same ncu capture)
        "note": "513 strictly dependent steps per launch: neither HBM- nor tensor-bound but latency-bound (MMA -> TMEM epilogue -> "
                "cluster exchange per step), so the HBM figure above is reported for the contract and the per-step latency is the "
                "number to improve; in the pipeline a launch holds "
                f"{8 * 2 * ((BATCH + pinfo['slots_per_cluster'] - 1) // max(1, pinfo['slots_per_cluster']))} SMs and {pinfo['recurrence_streams']} launches "
                f"(one per LSTM layer) run side by side on a {pinfo['recurrence_sms']}-SM partition while the other kernels of the steps in "
                f"flight use the remaining {pinfo['other_sms']} SMs",
        "whole_step": {"ms": ms_step, "algorithmic_bytes": step_bytes, "hbm_GBs": step_bytes / (ms_step / 1e3) / 1e9,
                       "hbm_frac": step_bytes / (ms_step / 1e3) / 1e9 / peaks["hbm_gbs"],
                       "algorithmic_flops": step_flops, "tflops": step_flops / (ms_step / 1e3) / 1e12,
                       "tensor_frac": step_flops / (ms_step / 1e3) / 1e12 / tpeak},
        "serial_stage_ms": {k: round(v, 4) for k, v in stage_acc.items()},
    }
    # every kernel family of the step against ITS roofline, from the live serial stage times (one blocking `sample` call, 148 SMs):
    # dense layers on the tensor pipe (fp32-equivalent FLOPs; the bf16x3 split issues 3x as many), STFT / iSTFT on HBM bytes
    bins, hid = 1025, 512
    lda1, ldm, ldz = (bins + 7) // 8 * 8, (bins + 3) // 4 * 4, (bins + 1) // 2 * 2
    fam = []
    gemm_flops = {"fc1": 2.0 * M * hid * bins, "fc2": 2.0 * M * hid * 2 * hid, "fc3": 2.0 * M * bins * hid}
    for l in range(3):
        gemm_flops[f"wih{l}"] = 2.0 * M * 8 * H * hid
    for k, fl_ in gemm_flops.items():
        if k in stage_acc and stage_acc[k] > 0:
            tf = fl_ / (stage_acc[k] / 1e3) / 1e12
            fam.append({"kernel": f"gemm2_kernel ({k})", "bound": "tensor", "ms": round(stage_acc[k], 4), "achieved": tf, "unit": "TFLOP/s fp32-equivalent",
                        "peak": tpeak, "frac": tf / tpeak, "frac_counting_the_3_bf16_passes": 3 * tf / tpeak})
    byt = {"stft": BATCH * T * 4 + M * ldz * 8 + M * lda1 * 4, "istft": M * ldz * 8 + M * ldm * 4 + BATCH * T * 4}
    for k, b_ in byt.items():
        if k in stage_acc and stage_acc[k] > 0:
            gb = b_ / (stage_acc[k] / 1e3) / 1e9
            fam.append({"kernel": f"{k}2048_tma_kernel", "bound": "hbm", "ms": round(stage_acc[k], 4), "achieved": gb, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                        "frac": gb / peaks["hbm_gbs"], "algorithmic_bytes": b_})
    roofline["kernel_families_serial"] = fam

    line = {
        "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * BATCH, "chunk_samples": T, "sample_rate": SR,
                   "parallelism": f"dp{world} (items sharded, no data-path collective)",
                   "gemm": "tcgen05 bf16x3 (fp32-grade)", "l2": f"inputs rotate over {NBUF} buffers (168 MB > 126 MB L2); "
                   "~575 MB of intermediates stream through HBM every step",
                   "schedule": f"{depth}-lane pipeline (OpenUnmixModel.pipeline / rfx_umx_pipe_push), fill + drain inside the timed region",
                   "single_call_ms": serial_ms, "pipeline": pinfo},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": BATCH * T * 4,
                "d2h_bytes_per_step": BATCH * T * 4,
                "api": "OpenUnmixModel.pipeline().push/wait on pinned host tensors (rfx_umx_pipe_push), every result consumed",
                "blocking_call": {"value": world * BATCH * CHUNK_S / (sync_ms / 1e3), "ms_per_step": sync_ms,
                                  "api": "OpenUnmixModel.sample_host (rfx_umx_sample_host), one blocking call per step"}},
        "bf16_fast": bf16_fast,
        "gpu_launches": args.steps * model.launches_per_call(),
        "clocks": clocks,
        "roofline": roofline,
    }

    # release the headline model's pipeline workspace before the other legs
    del pipe, outs_dev
    torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, cores, threads, kind, what = _cpu_reference(steps=3, warmup=1, items=BATCH)
        best = min(times)
        line["cpu_baseline"] = {"value": BATCH * CHUNK_S / best, "unit": "audio-s/s", "cores": threads, "kind": kind,
                                "sample": f"{what} on the full batch (32x262144), best of 3, "
                                          f"os.cpu_count()={cores}, median {statistics.median(times):.2f}s"}
    if world == 1 and not args.no_gpu_eager:
        line["gpu_eager"] = _gpu_eager(dev, xs[0])

    legs = [] if args.legs in ("none", "") else [t.strip() for t in args.legs.split(",")]
    others = []
    for tag, fn in (("3", leg_demucs_forward), ("4", leg_chain), ("5", leg_train)):
        if tag not in legs:
            continue
        try:
            if tag == "5":
                others.append(fn(dev, rank, world, barrier, reduce_max, peaks, use_dist))
            else:
                others.append(fn(dev, rank, world, barrier, reduce_max, peaks))
        except Exception as exc:  # a secondary leg must not take the headline line down; it is reported as failed
            others.append({"config": tag, "error": f"{type(exc).__name__}: {exc}"[:400]})
            try:
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
            except Exception:
                pass
    if others:
        line["other_configs"] = others
    if rank == 0:
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true")
    ap.add_argument("--legs", default="3,4,5", help="BASELINE.json configs measured beside the headline (comma list of 3,4,5 or 'none')")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

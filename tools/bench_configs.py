"""Measured numbers for every single-GPU BASELINE.json config (secondary to bench.py, which measures the headline config):

  config 1  TCN forward, 1 x 262144                      (+ the CPU oracle on the same chunk)
  config 2  Open-Unmix sample, 32 x 262144               -> bench.py
  config 3  Hybrid-Demucs forward + MR-STFT/L1 loss, 32 x 262144
  config 4  RemFX-detect chain: Cnn14 classifier + 5 effect-specific removers, per-GPU share 2 x 262144 (and 16 x 262144)

One JSON line per config: GPU ms/call (CUDA events, median of n), audio-s/s, fp32-equivalent TFLOP/s, and the CPU oracle
(torch-CPU restatement of the reference path, all host threads) on a bounded sample.  Usage: python tools/bench_configs.py [--no-cpu]
"""
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remfx_b200.synth import synth_audio  # noqa: E402

T, SR = 262144, 48000
CHUNK_S = T / SR
CPU = "--no-cpu" not in sys.argv


def gpu_time(fn, n=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def cpu_time(fn, n=2):
    torch.set_num_threads(os.cpu_count() or 1)
    torch.set_flush_denormal(True)
    ts = []
    with torch.no_grad():
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
    return min(ts)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def config1_tcn():
    from oracle import tcn as otcn
    from oracle import weights
    from remfx_b200.models import TCNModel

    sd = weights.tcn_state(0)
    m = TCNModel(sample_rate=SR, num_bins=1025, ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256, kernel_size=7,
                 stack_size=10, dilation_growth=2, causal=False)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    for B in (1, 4):
        x = synth_audio(1, B, T).cuda()
        med, best = gpu_time(lambda: m.sample(x), n=5)
        line = dict(config=f"TCN forward {B}x262144", ms=med, ms_best=best, audio_s_per_s=B * CHUNK_S / (med / 1e3),
                    tflops_fp32_equiv=5135.5e9 * B / (med / 1e3) / 1e12, launches=m.launches_per_call() if hasattr(m, "launches_per_call") else None)
        if CPU and B == 1:
            Ts = 65536  # bounded sample: a quarter chunk (5.1 TFLOP per full chunk would take minutes on the host)
            xs = synth_audio(1, 1, Ts)
            dt = cpu_time(lambda: otcn.sample(xs, sd, fused=True), n=1)
            line["cpu_oracle"] = dict(audio_s_per_s=Ts / SR / dt, seconds=dt, sample=f"1x{Ts}", cores=torch.get_num_threads())
        emit(**line)


def config3_demucs():
    from oracle import hdemucs as ohd
    from oracle import loss as oloss
    from remfx_b200.models import DemucsModel

    torch.manual_seed(0)
    m = DemucsModel(sample_rate=SR, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).cuda().eval()
    for B in (1, 32):
        x, y = synth_audio(1, B, T).cuda(), synth_audio(2, B, T).cuda()
        med, best = gpu_time(lambda: m((x, y)), n=5)
        line = dict(config=f"Hybrid-Demucs forward + MRSTFT/L1 loss {B}x262144", ms=med, ms_best=best, audio_s_per_s=B * CHUNK_S / (med / 1e3),
                    tflops_fp32_equiv=117.0e9 * B / (med / 1e3) / 1e12, launches=m.launches_per_call(B, T))
        if CPU and B == 1:
            ref = ohd.build(0)
            xc, yc = synth_audio(1, 1, T), synth_audio(2, 1, T)

            def run():
                out = ohd.sample(xc, ref)
                return oloss.remfx_loss(out, yc)
            dt = cpu_time(run, n=2)
            line["cpu_oracle"] = dict(audio_s_per_s=CHUNK_S / dt, seconds=dt, sample="1x262144 (torchaudio HDemucs + loss restatement)",
                                      cores=torch.get_num_threads())
        emit(**line)


def config4_chain():
    from oracle import chain as ochain
    from oracle import cnn14 as ocnn
    from oracle import umx as oumx
    from oracle import weights
    from remfx_b200.chain import RemFXChainInference
    from remfx_b200.classifier import Cnn14
    from remfx_b200.models import OpenUnmixModel

    order = ["RandomPedalboardDistortion", "RandomPedalboardCompressor", "RandomPedalboardReverb", "RandomPedalboardChorus",
             "RandomPedalboardDelay"]
    sds = {e: weights.umx_state(50 + i) for i, e in enumerate(ochain.ALL_EFFECTS)}
    members = {}
    for e, sd in sds.items():
        mm = OpenUnmixModel(sample_rate=SR)
        mm.load_state_dict(sd)
        members[e] = mm.cuda().eval()
    csd = weights.cnn14_state(0)
    clf = Cnn14(num_classes=5, sample_rate=SR, model_sample_rate=SR, n_fft=2048, hop_length=512, n_mels=128, specaugment=True)
    clf.load_state_dict(csd)
    clf = clf.cuda().eval()
    for B, use_all in ((2, True), (16, True), (16, False)):
        x, y = weights.synth_diverse(77, B, T).cuda(), synth_audio(78, B, T).cuda()
        chain = RemFXChainInference(members, SR, 1025, order, classifier=clf, use_all_effect_models=use_all)
        med, best = gpu_time(lambda: chain((x, y, None, None), 0), n=5)
        line = dict(config=f"RemFX-detect chain (Cnn14 + 5 Open-Unmix removers) {B}x262144, "
                           + ("all effect models" if use_all else "effects chosen by the classifier's decisions"),
                    ms=med, ms_best=best, audio_s_per_s=B * CHUNK_S / (med / 1e3))
        if not use_all:
            line["effects_per_item"] = float(chain.last_labels.sum(1).mean())
        if CPU and B == 2:
            xc, yc = x.cpu(), y.cpu()
            omem = {e: (lambda sd: (lambda z: oumx.sample(z, sd)))(sd) for e, sd in sds.items()}
            dt = cpu_time(lambda: ochain.forward(xc, yc, None, omem, order, classify=lambda z: torch.hstack(ocnn.forward(z, csd)), use_all=True), n=1)
            line["cpu_oracle"] = dict(audio_s_per_s=B * CHUNK_S / dt, seconds=dt, sample="2x262144, all effect models", cores=torch.get_num_threads())
        emit(**line)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["1", "3", "4"]
    if "1" in which:
        config1_tcn()
    if "3" in which:
        config3_demucs()
    if "4" in which:
        config4_chain()

"""Hybrid-Demucs training step at the benchmark size (cfg/exp/5-5_full.yaml: HDemucs, 16 chunks of 262144 per GPU):
forward (kept pre-activations) + MRSTFT/L1 loss + backward + [NCCL all-reduce] + clip + AdamW.

    python tools/hd_train_bench.py [--batch 16] [--steps 4] [--warmup 2] [--T 262144]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/hd_train_bench.py ...

One JSON line: ms per stage (CUDA events on the launching stream), audio-seconds/s of the whole step, algorithmic TFLOP/s
(3 x the forward's 117 GFLOP per chunk), peak memory, losses.  Synthetic data, seeded random-init weights."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--T", type=int, default=262144)
    a = ap.parse_args()
    from remfx_b200.losses import remfx_loss
    from remfx_b200.models import DemucsModel
    from remfx_b200.optim import configure_optimizers
    from remfx_b200.synth import synth_audio

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    torch.manual_seed(0)
    m = DemucsModel(sample_rate=48000, sources=["mixture"], audio_channels=1, nfft=4096, channels=48).cuda()
    opt = configure_optimizers(m, max_steps=50000)["optimizer"]
    opt.set_timing(True)
    x = synth_audio(12345 + rank, a.batch, a.T).cuda()
    t = synth_audio(54321 + rank, a.batch, a.T).cuda()
    acc = [0.0] * 3
    losses = []
    torch.cuda.reset_peak_memory_stats()
    for it in range(a.warmup + a.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        out = m._sample_train(x)
        ev[1].record()
        loss = remfx_loss(out, t)
        loss.backward()
        ev[2].record()
        opt.step()
        ev[3].record()
        torch.cuda.synchronize()
        losses.append(float(loss.detach()))
        if it >= a.warmup:
            for i in range(3):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
    ms = [v / a.steps for v in acc]
    tm = opt.timing_ms()[a.warmup:]
    ar = sum(v for v, _ in tm) / max(1, len(tm))
    if world > 1:
        import torch.distributed as dist

        tms = torch.tensor(ms + [ar], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms, ar = [float(v) for v in tms[:3]], float(tms[3])
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    total = sum(ms)
    nparam = sum(p.numel() for p in m.parameters())
    print(json.dumps({
        "workload": f"Hybrid-Demucs training step (forward_train + MRSTFT/100 L1 + backward + clip 10 + AdamW), batch {a.batch}x{a.T} per GPU",
        "n_gpus": world, "scaling": "weak", "ms_per_step": total, "stage_ms": dict(zip(["forward", "loss+backward", "optimizer(all-reduce+clip+adamw)"], ms)),
        "all_reduce_ms": ar, "all_reduce_bytes": nparam * 4,
        "audio_s_per_s": world * a.batch * a.T / 48000.0 / (total * 1e-3),
        "algorithmic_tflops_per_gpu": 3 * 117.0e9 * a.batch * (a.T / 262144.0) / (total * 1e-3) / 1e12,
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "losses": losses, "grad_norm_last": float(opt.total_norm),
        "steps": a.steps, "warmup": a.warmup}))


if __name__ == "__main__":
    main()

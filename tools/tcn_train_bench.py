"""TCN training step at the benchmark size: forward (kept activations) + MRSTFT/L1 loss + backward + clip + AdamW.

    python tools/tcn_train_bench.py [--batch 1] [--steps 5] [--warmup 2] [--T 262144]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/tcn_train_bench.py ...

Prints one JSON line: ms per stage (CUDA events on the launching stream), audio-seconds/s of the whole step, algorithmic
TFLOP/s (3 x the forward's 5.136 TFLOP per chunk: forward, input gradient, weight gradient; the recomputed pre-activations
are extra work and not counted), and the loss of every step.  Synthetic data, seeded random weights (oracle.weights)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--T", type=int, default=262144)
    ap.add_argument("--cpu-baseline", action="store_true",
                    help="time the same step done by torch on the host cores instead (oracle network under autograd + oracle loss + "
                         "clip_grad_norm_ + torch.optim.AdamW), on --cpu-T samples per item; no GPU needed")
    ap.add_argument("--cpu-T", type=int, default=32768)
    a = ap.parse_args()
    if a.cpu_baseline:
        return cpu_baseline(a)
    from oracle import weights  # seeded synthetic weights / audio only; nothing is computed by the oracle here
    from remfx_b200.models import TCNModel
    from remfx_b200.optim import configure_optimizers

    # data parallel: one process per GPU under torch.distributed.run; every rank trains on its own shard (weak scaling) and
    # FusedAdamW.step() runs the one gradient all-reduce of the step over NCCL
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    m = TCNModel(sample_rate=48000, num_bins=1025, ninputs=1, noutputs=1, nblocks=20, channel_growth=0, channel_width=256,
                 kernel_size=7, stack_size=10, dilation_growth=2, condition=False, latent_dim=2, norm_type="identity", causal=False,
                 estimate_loudness=False)
    m.load_state_dict(weights.tcn_state(0), strict=True)
    m = m.cuda()
    opt = configure_optimizers(m, max_steps=1000)["optimizer"]
    x = weights.synth_audio(12345 + rank, a.batch, a.T).cuda()
    t = weights.synth_audio(54321 + rank, a.batch, a.T).cuda()
    names = ["forward", "loss+backward", "optimizer"]
    acc = [0.0] * 3
    losses = []
    for it in range(a.warmup + a.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        out = m._sample_train(x)
        ev[1].record()
        from remfx_b200.losses import remfx_loss
        from remfx_b200.ops import causal_crop

        loss = remfx_loss(out, causal_crop(t, out.shape[-1]))
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
    if world > 1:
        import torch.distributed as dist

        tms = torch.tensor(ms, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)   # the step is as slow as its slowest rank
        ms = [float(v) for v in tms]
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    total = sum(ms)
    audio_s = world * a.batch * a.T / 48000.0
    L = m.out_length(a.T)
    fwd_tflop = 5.1355 * world * a.batch * (a.T / 262144.0)
    print(json.dumps({
        "workload": f"TCN training step (forward_train + MRSTFT/100 L1 + backward + clip 10 + AdamW), batch {a.batch}x{a.T} per GPU",
        "n_gpus": world, "scaling": "weak", "collective": "one NCCL all-reduce of the flat fp32 gradient bucket per step" if world > 1 else "none",
        "ms_per_step": total, "stage_ms": dict(zip(names, ms)), "audio_s_per_s": audio_s / (total * 1e-3),
        "algorithmic_tflops": 3 * fwd_tflop / (total * 1e-3), "out_length": L, "losses": losses,
        "grad_norm_last": float(opt.total_norm), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
        "steps": a.steps, "warmup": a.warmup}))


def cpu_baseline(a):
    """The reference's way of doing the step (torch autograd on the CPU; oracle/tcn.py issues the same torch ops as
    remfx/tcn.py), on a bounded sample: audio-seconds of OUTPUT-producing input per second of step time."""
    import time

    from oracle import tcn as otcn
    from oracle import weights

    torch.set_flush_denormal(True)
    sd = {k: v.clone().requires_grad_(True) for k, v in weights.tcn_state(0).items()}
    params = list(sd.values())
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.95, 0.999), eps=1e-6, weight_decay=1e-3)
    x = weights.synth_audio(12345, a.batch, a.cpu_T)
    t = weights.synth_audio(54321, a.batch, a.cpu_T)
    times, losses = [], []
    for _ in range(max(1, min(a.steps, 3))):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _ = otcn.forward((x, t), sd)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 10.0)
        opt.step()
        times.append(time.perf_counter() - t0)
        losses.append(float(loss.detach()))
    best = min(times)
    print(json.dumps({
        "workload": f"TCN training step on the host CPU (torch autograd through the oracle), batch {a.batch}x{a.cpu_T}",
        "seconds_per_step": best, "audio_s_per_s": a.batch * a.cpu_T / 48000.0 / best, "cores": torch.get_num_threads(),
        "cpu_count": os.cpu_count(), "losses": losses, "kind": "port",
        "sample": f"{a.batch}x{a.cpu_T} samples per step (receptive field 12277), best of {len(times)}"}))


if __name__ == "__main__":
    main()

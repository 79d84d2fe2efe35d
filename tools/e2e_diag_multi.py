"""Host <-> device copy ceiling of the box, per rank and in aggregate (e2e scaling 0.40 at N = 8).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/e2e_diag_multi.py

Every rank copies 33.5 MB pinned buffers (the bench's per-step input and output) H2D and D2H on two streams at once -- exactly
the traffic of one e2e step, no kernels -- first with ALL ranks copying at the same time, then rank by rank alone.  Prints one
JSON line: per-rank and aggregate GB/s in both modes, the e2e step time those rates allow, CPU / NUMA facts."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NBYTES = 32 * 262144 * 4
REPS = 60


def copy_rate(xh, xd, od, oh, s1, s2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(REPS):
        with torch.cuda.stream(s1):
            xd.copy_(xh, non_blocking=True)
        with torch.cuda.stream(s2):
            oh.copy_(od, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / REPS
    return 2 * NBYTES / dt / 1e9, dt * 1e3   # GB/s (both directions together), ms per (H2D + D2H) pair


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    xh = torch.empty(NBYTES // 4).pin_memory()
    oh = torch.empty(NBYTES // 4).pin_memory()
    xh.normal_()
    xd = torch.empty(NBYTES // 4, device="cuda")
    od = torch.randn(NBYTES // 4, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    copy_rate(xh, xd, od, oh, s1, s2)  # warm-up
    if world > 1:
        dist.barrier()
    together = copy_rate(xh, xd, od, oh, s1, s2)
    alone = (0.0, 0.0)
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            alone = copy_rate(xh, xd, od, oh, s1, s2)
    res = torch.tensor([together[0], together[1], alone[0], alone[1]], device="cuda", dtype=torch.float64)
    if world > 1:
        allr = [torch.zeros_like(res) for _ in range(world)]
        dist.all_gather(allr, res)
    else:
        allr = [res]
    if rank == 0:
        tg = [float(t[0]) for t in allr]
        al = [float(t[2]) for t in allr]
        numa = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) if os.path.isdir("/sys/devices/system/node") else []
        print(json.dumps({
            "n_gpus": world, "bytes_each_way_per_step": NBYTES,
            "all_ranks_at_once": {"per_rank_GBs": [round(v, 1) for v in tg], "aggregate_GBs": round(sum(tg), 1),
                                  "ms_per_step_pair_max": round(max(float(t[1]) for t in allr), 3)},
            "one_rank_at_a_time": {"per_rank_GBs": [round(v, 1) for v in al], "ms_per_step_pair_max": round(max(float(t[3]) for t in allr), 3)},
            "copy_only_e2e_efficiency_bound": round((sum(tg) / world) / (sum(al) / world), 3),
            "cpu_count": os.cpu_count(), "numa_nodes": numa,
            "note": "H2D and D2H of 33.5 MB each on two streams per rank, pinned memory, no kernels: the floor of an e2e step"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

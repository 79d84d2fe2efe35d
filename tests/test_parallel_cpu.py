"""`not gpu`: the N>1 path -- item sharding and gathering -- on the gloo backend with world_size 2 (and 3)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from remfx_b200.parallel import run_sharded, shard_range, shard_sizes


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 7, 16, 32, 33):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import stft as ostft  # a real per-item op of the path (CPU stand-in for model.sample)

        g = torch.Generator().manual_seed(0)
        x = torch.randn(n_items, 1, 4096, generator=g)
        calls = []

        def fn(xs):
            calls.append(xs.shape[0])
            return ostft.spectrogram(xs, torch.hann_window(512), 512, 128, 0.3)

        full = run_sharded(fn, x, gather=True)
        ref = ostft.spectrogram(x, torch.hann_window(512), 512, 128, 0.3)
        lo, hi = shard_range(n_items, rank, world)
        ok = full.shape == ref.shape and torch.allclose(full, ref, rtol=1e-5, atol=1e-7) and sum(calls) == hi - lo
        local = run_sharded(fn, x, gather=False)
        ok = ok and (local is None if hi == lo else torch.allclose(local, ref[lo:hi], rtol=1e-5, atol=1e-7))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 5), (2, 4), (3, 2)])
def test_run_sharded_gloo(world, n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True) for r in range(world)]

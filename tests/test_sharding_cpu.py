"""N>1 path on CPU: world_size-2 gloo run of the shard + gather logic bench.py / a multi-GPU eval uses
(no data-path collective; only the 12 B/joint results are exchanged)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rsgnet_b200.sharding import gather_results, shard_bounds, shard_images


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 256, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    off = np.array([0, 3, 3, 10, 12, 20])
    assert shard_images(off, 0, 2) == (0, 3, 0, 10) and shard_images(off, 1, 2) == (3, 5, 10, 20)


def _fake_results(lo, hi, K):
    idx = torch.arange(lo, hi, dtype=torch.float32)
    preds = idx[:, None, None] * 10 + torch.arange(K, dtype=torch.float32)[None, :, None] + torch.tensor([0.25, 0.5])
    return preds, (idx[:, None, None] / 100 + torch.zeros(1, K, 1))


def _worker(rank, world, port, n, K, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_bounds(n, rank, world)
    preds, mv = _fake_results(lo, hi, K)
    full_p, full_m = gather_results(preds, mv, n)
    ref_p, ref_m = _fake_results(0, n, K)
    q.put((rank, bool(torch.equal(full_p, ref_p) and torch.equal(full_m, ref_m))))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_world2_gloo():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 14, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]

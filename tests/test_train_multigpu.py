"""Data-parallel training step on 2 GPUs of one box (one process per GPU, NCCL): the flat gradient buffer after
TrainStep.all_reduce is the SUM of the ranks' gradients, Adam applies their mean, and the replicas stay bit-identical.
Skips on a single-GPU box (run it with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from rsgnet_b200 import presets, synth
    from rsgnet_b200.models import _params, pose_rsgnet
    from rsgnet_b200.train import TrainStep
    cfg = presets.preset('tiny_cp')
    net = pose_rsgnet.get_pose_net(cfg, True)
    net.load_state_dict(_params.synth_state_dict(net, seed=1))
    net = net.to(dev).train()
    b = synth.train_batch(2, cfg.MODEL.IMAGE_SIZE, cfg.MODEL.HEATMAP_SIZE, 14, 14, seed=10 + rank)       # a different shard per rank
    batch = [torch.from_numpy(b[k]).to(dev) for k in ('input', 'target', 'target_weight', 'all_ins_target',
                                                       'all_ins_target_weight', 'target_limbs')]
    ts = TrainStep(net, lr=1e-3)
    assert ts.world == world
    ts.forward_backward(*batch)
    mine = ts.store.flat_g.clone()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    ts.all_reduce()
    torch.cuda.synchronize()
    ok_sum = bool(torch.allclose(ts.store.flat_g, sum(parts), rtol=1e-6, atol=1e-9))
    differ = bool((parts[0] - parts[1]).abs().max() > 0)
    before = ts.store.flat_p.clone()
    ts.adam()
    torch.cuda.synchronize()
    moved = bool((ts.store.flat_p - before).abs().max() > 5e-4)
    digest = [torch.empty_like(ts.store.flat_p) for _ in range(world)]
    dist.all_gather(digest, ts.store.flat_p)
    same = bool(torch.equal(digest[0], digest[1]))
    # two more full steps through the public call
    for _ in range(2):
        L, _ = ts(*batch)
    digest = [torch.empty_like(ts.store.flat_p) for _ in range(world)]
    dist.all_gather(digest, ts.store.flat_p)
    same = same and bool(torch.equal(digest[0], digest[1]))
    q.put((rank, ok_sum, differ, moved, same, bool(L['loss'] == L['loss'])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_data_parallel_train_step_two_gpus():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True, True, True, True, True), (1, True, True, True, True, True)], got

"""Multi-GPU: the path shards by crop -- contiguous batch slices, one process per GPU,
weights replicated, NO data-path collective.  The only exchange is the gather of the 12 bytes per joint
of results (and, for NMS, detections of one image must end up on one rank: shard NMS by image)."""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of `n` crops for `rank`; sizes differ by at most one."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_images(img_offsets, rank, world):
    """Slice of IMAGES for `rank` (NMS groups are per image): returns (img_lo, img_hi, det_lo, det_hi)."""
    off = np.asarray(img_offsets)
    lo, hi = shard_bounds(len(off) - 1, rank, world)
    return lo, hi, int(off[lo]), int(off[hi])


def gather_results(preds, maxvals, n_total, group=None):
    """All ranks get the concatenated (preds [n_total,K,2], maxvals [n_total,K,1]) of every rank's
    shard (shards as produced by shard_bounds).  Uses torch.distributed (NCCL for CUDA tensors, gloo for
    CPU tensors); with world size 1 it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return preds, maxvals
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    K = preds.shape[1]
    packed = torch.zeros((width, K, 3), dtype=torch.float32, device=preds.device)
    lo, hi = sizes[rank]
    packed[:hi - lo, :, :2] = preds
    packed[:hi - lo, :, 2:] = maxvals
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    full = torch.cat([o[:h - l] for o, (l, h) in zip(out, sizes)])
    return full[:, :, :2].contiguous(), full[:, :, 2:].contiguous()

"""Device-resident flip-test inference: what lib/core/function.py:389-450 does
per batch -- forward, forward on the W-flipped crops, flip_back, 1-px shift, average,
get_final_preds -- without the reference loop's three device->host copies of full heat-maps.

    preds, maxvals = infer_crops(model, cfg, x, center, scale)

The crops are run as one batch [x ; flip(x)] (the stem kernel reads the second half W-reversed, no
flipped copy is materialised), and one fused kernel turns the two raw heat-map sets into 12 bytes per
joint.
"""
import numpy as np
import torch

from . import _engine, _lib
from .config import KIND_RSGNET, cfg_get
from .core.inference import decode_device
from .presets import flip_pairs_for
from .utils.transforms import flip_perm


class CropPipeline:
    """Static-buffer runner for a fixed crops-per-step count (CUDA-graph replayable)."""

    def __init__(self, model, cfg, batch, flip_pairs=None, device=None, use_graph=True, chunk=None):
        _lib.require_cuda()
        self.model = model
        self.device = torch.device(device or next(model.parameters()).device)
        self.spec = model.spec
        self.batch = batch
        self.flip = bool(cfg_get(cfg, 'TEST', 'FLIP_TEST', default=True))
        self.post = bool(cfg_get(cfg, 'TEST', 'POST_PROCESS', default=True))
        self.shift = bool(cfg_get(cfg, 'TEST', 'SHIFT_HEATMAP', default=True))
        K = self.spec.num_joints
        pairs = flip_pairs if flip_pairs is not None else flip_pairs_for(K)
        self.perm = flip_perm(K, pairs)
        self.engine = _engine.engine_for(model, self.device, chunk)
        self.use_graph = use_graph
        nf = batch * (2 if self.flip else 1)
        dev = self.device
        self.x = torch.empty((batch, 3, self.spec.image_h, self.spec.image_w), dtype=torch.float32, device=dev)
        self.heat = torch.empty((nf, K, self.spec.heat_h, self.spec.heat_w), dtype=torch.float32, device=dev)
        self.center = torch.empty((batch, 2), dtype=torch.float32, device=dev)
        self.scale = torch.empty((batch, 2), dtype=torch.float32, device=dev)
        self.n_fwd = nf

    def run_device(self):
        """Inputs already in self.x / self.center / self.scale.  Returns CUDA (preds, maxvals)."""
        self.engine.run(self.x, self.heat, self.n_fwd, self.batch, use_graph=self.use_graph)
        hm = self.heat[:self.batch]
        hf = self.heat[self.batch:] if self.flip else None
        out = decode_device(hm, self.center, self.scale, post_process=self.post, hm_flipped=hf,
                            flip_perm=self.perm, shift=self.shift)
        return out['preds'], out['maxvals']

    def launches_per_step(self):
        return self.engine.last_launches() + 1

    def __call__(self, x, center, scale):
        """Host (pinned or pageable) or device inputs -> NumPy (preds [B,K,2], maxvals [B,K,1])."""
        self.x.copy_(torch.as_tensor(x), non_blocking=True)
        self.center.copy_(torch.as_tensor(np.asarray(center, np.float32) if not isinstance(center, torch.Tensor) else center), non_blocking=True)
        self.scale.copy_(torch.as_tensor(np.asarray(scale, np.float32) if not isinstance(scale, torch.Tensor) else scale), non_blocking=True)
        preds, maxvals = self.run_device()
        return preds.cpu().numpy(), maxvals.cpu().numpy()


def infer_crops(model, cfg, x, center, scale, flip_pairs=None):
    """One-shot convenience wrapper around CropPipeline (no graph replay)."""
    pipe = CropPipeline(model, cfg, int(x.shape[0]), flip_pairs=flip_pairs, use_graph=False)
    return pipe(x, center, scale)

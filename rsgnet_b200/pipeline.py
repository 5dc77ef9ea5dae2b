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
from .utils.transforms import affine_matrices, flip_perm, warp_crops


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
        # two input slots: slot i+1 is filled from pinned host memory on a copy stream while slot i computes
        self._slots = [dict(x=torch.empty((batch, 3, self.spec.image_h, self.spec.image_w), dtype=torch.float32, device=dev),
                            center=torch.empty((batch, 2), dtype=torch.float32, device=dev),
                            scale=torch.empty((batch, 2), dtype=torch.float32, device=dev),
                            ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
        self._cur = 0
        self._copy_stream = torch.cuda.Stream(dev)
        self.heat = torch.empty((nf, K, self.spec.heat_h, self.spec.heat_w), dtype=torch.float32, device=dev)
        self.n_fwd = nf

    # the current input slot (what run_device consumes)
    @property
    def x(self):
        return self._slots[self._cur]['x']

    @property
    def center(self):
        return self._slots[self._cur]['center']

    @property
    def scale(self):
        return self._slots[self._cur]['scale']

    def run_device(self):
        """Inputs already in self.x / self.center / self.scale.  Returns CUDA (preds, maxvals)."""
        self.engine.run(self.x, self.heat, self.n_fwd, self.batch, use_graph=self.use_graph)
        hm = self.heat[:self.batch]
        hf = self.heat[self.batch:] if self.flip else None
        out = decode_device(hm, self.center, self.scale, post_process=self.post, hm_flipped=hf,
                            flip_perm=self.perm, shift=self.shift)
        return out['preds'], out['maxvals']

    def run_overlapped(self, batches, out=None):
        """Process a sequence of host batches (x, center, scale) -- pinned tensors for true overlap --
        with the H2D copy of batch i+1 running on a copy stream under the compute of batch i.
        Returns a list of (preds, maxvals) CUDA tensors, or fills `out` = list of (preds_host,
        maxvals_host) pinned tensors with non-blocking D2H copies (synchronise before reading)."""
        main = torch.cuda.current_stream(self.device)
        results = []

        def upload(i, slot):
            sl = self._slots[slot]
            xb, cb, sb = batches[i]
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(sl['free'])          # previous consumer of this slot is done
                sl['x'].copy_(xb, non_blocking=True)
                sl['center'].copy_(cb, non_blocking=True)
                sl['scale'].copy_(sb, non_blocking=True)
                sl['ready'].record(self._copy_stream)

        for sl in self._slots:
            sl['free'].record(main)
        if len(batches):
            upload(0, self._cur)
        for i in range(len(batches)):
            slot = self._cur
            if i + 1 < len(batches):
                upload(i + 1, slot ^ 1)
            main.wait_event(self._slots[slot]['ready'])
            preds, maxvals = self.run_device()
            self._slots[slot]['free'].record(main)
            if out is not None:
                out[i][0].copy_(preds, non_blocking=True)
                out[i][1].copy_(maxvals, non_blocking=True)
            else:
                results.append((preds, maxvals))
            self._cur = slot ^ 1
        return results

    def infer_images(self, images, centers, scales, image_index=None, color_rgb=True, rots=0):
        """Photos + person boxes -> key points, everything after the image upload on the device (SURVEY.md §8f-3 +
        §8f-1): the loader's crop affine + cv2.warpAffine + ToTensor/Normalize (CPJointsDataset.py:1281-1290,
        cp_test.py:107-115) as one kernel writing straight into the model's input buffer, then the flip-test run
        and the fused decode.  images: uint8 HWC BGR arrays or CUDA tensors; centers/scales f32 [B,2] as the
        dataset's db records hold them.  Returns CUDA (preds, maxvals)."""
        centers = np.asarray(centers, np.float32)
        scales = np.asarray(scales, np.float32)
        assert centers.shape == (self.batch, 2) and scales.shape == (self.batch, 2), 'one box per crop of the step'
        mats = affine_matrices(centers, scales, rots, (self.spec.image_w, self.spec.image_h))
        warp_crops(images, mats, (self.spec.image_w, self.spec.image_h), image_index=image_index, color_rgb=color_rgb,
                   out=self.x)
        self.center.copy_(torch.from_numpy(centers), non_blocking=True)
        self.scale.copy_(torch.from_numpy(scales), non_blocking=True)
        return self.run_device()

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

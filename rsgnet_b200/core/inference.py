"""Drop-in for /root/reference/lib/core/inference.py.

``get_max_preds(batch_heatmaps)`` and ``get_final_preds(config, batch_heatmaps, center, scale)``
keep the reference's signatures, argument checks and return conventions (inference.py:21-82); the
arithmetic runs in ``rsg_flip_avg_decode`` (rsgnet_b200/csrc/decode.cu).  ``batch_heatmaps`` may
also be a CUDA ``torch.Tensor`` (a compatible extension that lets a device-resident loop skip the
D2H copy of the heat-maps); NumPy inputs are copied to the device, decoded there and the
12 bytes/joint of results copied back.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _dev_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device, torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device)


def decode_device(hm, center=None, scale=None, post_process=True, hm_flipped=None, flip_perm=None,
                  shift=True, want_coords=False, want_avg=False):
    """Device-resident decode.  hm[, hm_flipped]: f32 CUDA [N,K,H,W].  Returns a dict of CUDA
    tensors: maxvals [N,K,1], preds [N,K,2] (if center/scale), coords (heat-map space), avg."""
    _lib.require_cuda()
    assert hm.is_cuda and hm.dtype == torch.float32 and hm.dim() == 4
    hm = hm.contiguous()
    dev = hm.device
    N, K, H, W = hm.shape
    out = {'maxvals': torch.empty((N, K, 1), dtype=torch.float32, device=dev)}
    preds = coords = avg = None
    if center is not None:
        center, scale = _dev_f32(center, dev), _dev_f32(scale, dev)
        assert center.shape == (N, 2) and scale.shape == (N, 2)
        preds = out['preds'] = torch.empty((N, K, 2), dtype=torch.float32, device=dev)
    if want_coords:
        coords = out['coords'] = torch.empty((N, K, 2), dtype=torch.float32, device=dev)
    perm = None
    if hm_flipped is not None:
        assert hm_flipped.shape == hm.shape and hm_flipped.is_cuda
        hm_flipped = hm_flipped.contiguous()
        perm = torch.as_tensor(np.asarray(flip_perm, np.int32), device=dev)
        assert perm.numel() == K
        if want_avg:
            avg = out['avg'] = torch.empty_like(hm)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().rsg_flip_avg_decode(
            _lib.stream_ptr(dev), p(hm), p(hm_flipped), p(perm), N, K, H, W, p(center), p(scale),
            int(bool(post_process)), int(bool(shift)), p(preds), p(out['maxvals']), p(coords), p(avg)))
    return out


def get_max_preds(batch_heatmaps):
    """inference.py:21-49: (preds f32 [N,K,2] heat-map px, maxvals f32 [N,K,1])."""
    if not isinstance(batch_heatmaps, torch.Tensor):
        assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
        assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
        hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, np.float32)).cuda()
    else:
        assert batch_heatmaps.dim() == 4, 'batch_images should be 4-ndim'
        hm = batch_heatmaps.float()
    out = decode_device(hm, post_process=False, want_coords=True)
    return out['coords'].cpu().numpy(), out['maxvals'].cpu().numpy()


def get_final_preds(config, batch_heatmaps, center, scale):
    """inference.py:52-82: (preds f32 [N,K,2] image px, maxvals f32 [N,K,1]); reads only
    config.TEST.POST_PROCESS."""
    if not isinstance(batch_heatmaps, torch.Tensor):
        assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
        assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
        hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, np.float32)).cuda()
    else:
        assert batch_heatmaps.dim() == 4, 'batch_images should be 4-ndim'
        hm = batch_heatmaps.float()
    out = decode_device(hm, center, scale, post_process=bool(config.TEST.POST_PROCESS))
    return out['preds'].cpu().numpy(), out['maxvals'].cpu().numpy()

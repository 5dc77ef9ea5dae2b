"""Drop-in for the hot-path part of /root/reference/lib/utils/transforms.py.

``flip_back(output_flipped, matched_parts)`` (transforms.py:23-37) runs on the device
(rsg_flip_back); NumPy in -> NumPy out like the reference, CUDA tensor in -> CUDA tensor out.
Unlike the reference it never aliases or mutates its argument.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def flip_perm(num_joints, matched_parts):
    perm = np.arange(num_joints, dtype=np.int32)
    for a, b in matched_parts:
        perm[a], perm[b] = b, a
    return perm


def flip_back(output_flipped, matched_parts):
    is_np = not isinstance(output_flipped, torch.Tensor)
    if is_np:
        assert output_flipped.ndim == 4, 'output_flipped should be [batch_size, num_joints, height, width]'
        src = torch.from_numpy(np.ascontiguousarray(output_flipped, np.float32)).cuda()
    else:
        assert output_flipped.dim() == 4, 'output_flipped should be [batch_size, num_joints, height, width]'
        src = output_flipped.float().contiguous()
        if not src.is_cuda:
            src = src.cuda()
    _lib.require_cuda()
    N, K, H, W = src.shape
    perm = torch.as_tensor(flip_perm(K, matched_parts), device=src.device)
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().rsg_flip_back(_lib.stream_ptr(src.device), C.c_void_p(src.data_ptr()),
                                            C.c_void_p(out.data_ptr()), C.c_void_p(perm.data_ptr()),
                                            N, K, H, W))
    return out.cpu().numpy() if is_np else out

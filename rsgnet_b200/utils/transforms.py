"""Drop-in for the hot-path part of /root/reference/lib/utils/transforms.py.

``get_affine_transform`` / ``affine_transform`` (transforms.py:65-103) are host arithmetic (a 3-point affine in
fp64 from fp32-rounded control points); ``crop`` (transforms.py:121-129) and the batched ``warp_crops`` run the
warp on the device (rsg_warp_affine), bit-exact with ``cv2.warpAffine(..., flags=cv2.INTER_LINEAR)`` on 8-bit
images, optionally fused with the loader's BGR->RGB + ToTensor + Normalize (tools/cp_test.py:107-115).

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


# ---------------------------------------------------------------------------------------------
# Input side: crop affine + warp + normalise
# ---------------------------------------------------------------------------------------------
MEAN = (0.485, 0.456, 0.406)          # tools/cp_test.py:107-109
STD = (0.229, 0.224, 0.225)


def affine_matrices(centers, scales, rots, output_size, shift=(0.0, 0.0), inv=0):
    """Batched get_affine_transform (transforms.py:65-97): f64 [N,2,3].  The three control points are rounded to fp32
    exactly where the reference rounds them; the 3-point system cv2.getAffineTransform solves by LU is solved in
    fp64 by LAPACK (agreement ~1e-13)."""
    centers = np.asarray(centers, np.float32).reshape(-1, 2)
    n = centers.shape[0]
    scales = np.asarray(scales, np.float32)
    if scales.ndim == 0:
        scales = np.full((n, 2), scales, np.float32)
    scales = scales.reshape(n, -1)
    if scales.shape[1] == 1:
        scales = np.repeat(scales, 2, axis=1)
    rots = np.broadcast_to(np.asarray(rots, np.float64), (n,))
    shift = np.asarray(shift, np.float32)
    scale_tmp = scales * np.float32(200.0)
    src_w = scale_tmp[:, 0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rots / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p1 = (src_w * np.float32(-0.5)).astype(np.float64)
    src_dir = np.stack([0 * cs - p1 * sn, 0 * sn + p1 * cs], axis=1)                  # get_dir: fp64
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    sh = scale_tmp * shift
    src = np.zeros((n, 3, 2), np.float32)
    dst = np.zeros((n, 3, 2), np.float32)
    src[:, 0] = centers + sh
    src[:, 1] = (centers.astype(np.float64) + src_dir + sh).astype(np.float32)
    dst[:, 0] = [dst_w * 0.5, dst_h * 0.5]
    dst[:, 1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir

    def third(a, b):                                                                 # get_3rd_point
        d = a - b
        return b + np.stack([-d[:, 1], d[:, 0]], axis=1).astype(np.float32)
    src[:, 2] = third(src[:, 0], src[:, 1])
    dst[:, 2] = third(dst[:, 0], dst[:, 1])
    if inv:
        src, dst = dst, src
    a = np.concatenate([src.astype(np.float64), np.ones((n, 3, 1))], axis=2)
    t = np.linalg.solve(a, dst.astype(np.float64))                                  # [N,3,2]
    return np.ascontiguousarray(t.transpose(0, 2, 1))


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """transforms.py:65-97 -> f64 [2,3]."""
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale])
    return affine_matrices(np.asarray(center)[None], np.asarray(scale, np.float32)[None], rot, output_size,
                           shift=shift, inv=inv)[0]


def affine_transform(pt, t):
    """transforms.py:100-103."""
    new_pt = np.array([pt[0], pt[1], 1.]).T
    return np.dot(t, new_pt)[:2]


def normalize_lut(mean=MEAN, std=STD):
    """f32 [3,256]: ToTensor (u/255) then Normalize ((t - mean)/std), both in fp32 like torchvision."""
    u = np.arange(256, dtype=np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32)[:, None]
    s = np.asarray(std, np.float32)[:, None]
    return ((u[None, :] - m) / s).astype(np.float32)


def warp_crops(images, mats, output_size, image_index=None, color_rgb=False, normalize=True, mean=MEAN, std=STD,
               return_u8=False, device=None, out=None):
    """Warp N crops in one launch.  images: list of uint8 HWC (3-channel) NumPy arrays or CUDA tensors; crop i reads
    images[image_index[i]] (default i) through the forward matrix mats[i] (f64 [N,2,3], as get_affine_transform
    returns it).  output_size = (W, H) like the reference's IMAGE_SIZE.
    Returns the normalised f32 [N,3,H,W] CUDA tensor (normalize=True), the uint8 [N,H,W,3] crops (return_u8=True),
    or both as a tuple (f32, u8).  color_rgb: the sources are BGR (cv2.imread) and the output is RGB.
    out: optional preallocated contiguous f32 [>=N,3,H,W] CUDA tensor to write the normalised crops into."""
    _lib.require_cuda()
    device = torch.device(device if device is not None else 'cuda')
    mats = np.ascontiguousarray(np.asarray(mats, np.float64).reshape(-1, 6))
    n = mats.shape[0]
    if image_index is None:
        image_index = np.arange(n)
    image_index = np.asarray(image_index, np.int64)
    assert image_index.shape == (n,), 'one source image index per crop'
    w_out, h_out = int(output_size[0]), int(output_size[1])
    dev_imgs = []
    for im in images:
        if isinstance(im, torch.Tensor):
            t = im if im.is_cuda else im.to(device, non_blocking=True)
        else:
            assert im.dtype == np.uint8, 'images must be uint8'
            t = torch.from_numpy(np.ascontiguousarray(im)).to(device, non_blocking=True)
        assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3 and t.stride(2) == 1 and t.stride(1) == 3, \
            'images must be uint8 HWC with 3 interleaved channels'
        dev_imgs.append(t)
    meta = np.zeros((n, 4), np.int64)                       # pointer, rows, cols, row stride
    for i, j in enumerate(image_index):
        t = dev_imgs[int(j)]
        meta[i] = (t.data_ptr(), t.shape[0], t.shape[1], t.stride(0))
    ptrs = torch.from_numpy(meta[:, 0].copy()).to(device)
    dims = torch.from_numpy(meta[:, 1:].astype(np.int32)).to(device)
    mats_d = torch.from_numpy(mats).to(device)
    if out is not None:
        assert normalize and out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and \
            out.shape[0] >= n and tuple(out.shape[1:]) == (3, h_out, w_out), 'out must be contiguous f32 [>=N,3,H,W] on the device'
        device = out.device
        out_f32 = out[:n]
    else:
        out_f32 = torch.empty((n, 3, h_out, w_out), dtype=torch.float32, device=device) if normalize else None
    out_u8 = torch.empty((n, h_out, w_out, 3), dtype=torch.uint8, device=device) if return_u8 else None
    if not normalize and not return_u8:
        raise ValueError('nothing to compute: normalize=False and return_u8=False')
    lut = torch.from_numpy(normalize_lut(mean, std)).to(device) if normalize else None
    with torch.cuda.device(device):
        _lib.check(_lib.lib().rsg_warp_affine(
            _lib.stream_ptr(device), C.c_void_p(ptrs.data_ptr()), C.c_void_p(dims.data_ptr()),
            C.c_void_p(mats_d.data_ptr()), n, h_out, w_out, int(bool(color_rgb)),
            C.c_void_p(out_u8.data_ptr() if return_u8 else None),
            C.c_void_p(out_f32.data_ptr() if normalize else None),
            C.c_void_p(lut.data_ptr() if normalize else None)))
    if normalize and return_u8:
        return out_f32, out_u8
    return out_f32 if normalize else out_u8


def crop(img, center, scale, output_size, rot=0):
    """transforms.py:121-129: uint8 HWC in -> uint8 HWC out (NumPy), the warp done on the device."""
    trans = get_affine_transform(center, scale, rot, output_size)
    out = warp_crops([img], trans[None], output_size, normalize=False, return_u8=True)
    return out[0].cpu().numpy()


# ---------------------------------------------------------------------------------------------
# Host-side helpers the reference's callers import from utils.transforms next to the hot-path functions
# (lib/core/inference.py:16, lib/core/function.py:18, lib/dataset/*JointsDataset.py:20-22).  Plain NumPy; nothing on the
# inference hot path calls them (decode does the inverse affine on the device, csrc/decode.cu).
# ---------------------------------------------------------------------------------------------
def transform_preds(coords, center, scale, output_size):
    """transforms.py:57-62: heat-map coordinates [K,>=2] -> image coordinates through the inverse crop affine
    (rot = 0).  Returns a float64 array shaped like `coords` (extra columns zero), as the reference does."""
    coords = np.asarray(coords)
    t = get_affine_transform(center, scale, 0, output_size, inv=1)
    out = np.zeros(coords.shape)
    out[:, 0:2] = coords[:, 0:2] @ t[:, :2].T + t[:, 2]
    return out


def fliplr_joints(joints, joints_vis, width, matched_parts):
    """transforms.py:40-54 (training-time augmentation): mirror the x coordinates and swap the left/right joints, in
    place like the reference; returns (joints * joints_vis, joints_vis)."""
    joints[:, 0] = width - joints[:, 0] - 1
    for a, b in matched_parts:
        joints[[a, b]] = joints[[b, a]]
        joints_vis[[a, b]] = joints_vis[[b, a]]
    return joints * joints_vis, joints_vis


# DensePose part-index symmetry (transforms.py:14-21): background 0 and the torso parts 1, 2 map to themselves, the
# remaining 22 parts swap in left/right pairs (3<->4, 5<->6, ..., 23<->24)
_DP_SYMMETRY = np.array([0, 1, 2] + [v for a in range(3, 25, 2) for v in (a + 1, a)])


def flip_dp_back(I_heatmaps):
    """transforms.py:14-21 (imported by lib/core/function.py:18; used by no RSGNet config): undo the horizontal flip of
    a 25-channel DensePose part-index map: permute the channels by the left/right symmetry and reverse W."""
    I_heatmaps = np.asarray(I_heatmaps)
    return I_heatmaps[:, _DP_SYMMETRY][:, :, :, ::-1]

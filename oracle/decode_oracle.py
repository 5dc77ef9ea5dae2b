"""CPU oracle for flip-test averaging and heatmap decode.   *** TEST INFRASTRUCTURE ONLY ***

NumPy restatement of the reference's host post-processing.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
it; the product package never does.

Pinned by ``tests/golden/decode_*.npz`` produced by the unmodified reference
(``oracle/gen_golden.py``): integer argmax coordinates, maxvals and quarter-pixel offsets are
bit-exact; image-space coordinates are within 1 fp32 ulp (the reference solves the affine with
``cv2.getAffineTransform`` (LU in fp64); this file uses the closed form with the same fp32
roundings of the control points -- SURVEY.md App. B.1).

Reference lines followed (paths under /root/reference):
  lib/utils/transforms.py:23-37    flip_back
  lib/core/function.py:417-427     flip_back -> 1px shift -> (a+b)*0.5
  lib/core/inference.py:21-49      get_max_preds
  lib/core/inference.py:52-82      get_final_preds
  lib/utils/transforms.py:57-103   transform_preds / get_affine_transform(inv=1) / affine_transform
"""
import numpy as np


def flip_perm(num_joints, flip_pairs):
    """pi(k): channel that flip_back puts at position k."""
    perm = np.arange(num_joints, dtype=np.int32)
    for a, b in flip_pairs:
        perm[a], perm[b] = b, a
    return perm


def flip_back(output_flipped, flip_pairs):
    """transforms.py:23-37 (returns a fresh array; never aliases its input)."""
    assert output_flipped.ndim == 4
    out = output_flipped[:, :, :, ::-1].copy()
    for a, b in flip_pairs:
        tmp = out[:, a].copy()
        out[:, a] = out[:, b]
        out[:, b] = tmp
    return out


def flip_average(out, out_flipped_raw, flip_pairs, shift=True):
    """function.py:417-427.  fp32 add then multiply by 0.5."""
    b = flip_back(out_flipped_raw, flip_pairs)
    if shift:
        b[:, :, :, 1:] = b.copy()[:, :, :, 0:-1]
    return ((out + b) * np.float32(0.5)).astype(np.float32)


def get_max_preds(hm):
    """inference.py:21-49."""
    assert isinstance(hm, np.ndarray) and hm.ndim == 4
    n, k, h, w = hm.shape
    flat = hm.reshape(n, k, -1)
    idx = np.argmax(flat, 2)
    maxvals = np.amax(flat, 2).reshape(n, k, 1)
    preds = np.empty((n, k, 2), np.float32)
    preds[:, :, 0] = (idx % w).astype(np.float32)
    preds[:, :, 1] = (idx // w).astype(np.float32)
    preds *= np.greater(maxvals, 0.0).astype(np.float32)
    return preds, maxvals


def quarter_offset(hm, coords):
    """inference.py:59-72, vectorised.  coords are integer-valued fp32."""
    n, k, h, w = hm.shape
    px = coords[:, :, 0].astype(np.int64)
    py = coords[:, :, 1].astype(np.int64)
    ok = (px > 1) & (px < w - 1) & (py > 1) & (py < h - 1)
    pxc = np.clip(px, 1, w - 2)
    pyc = np.clip(py, 1, h - 2)
    ni = np.arange(n)[:, None]
    ki = np.arange(k)[None, :]
    dx = hm[ni, ki, pyc, pxc + 1] - hm[ni, ki, pyc, pxc - 1]
    dy = hm[ni, ki, pyc + 1, pxc] - hm[ni, ki, pyc - 1, pxc]
    out = coords.copy()
    out[:, :, 0] += np.where(ok, np.sign(dx) * np.float32(0.25), np.float32(0)).astype(np.float32)
    out[:, :, 1] += np.where(ok, np.sign(dy) * np.float32(0.25), np.float32(0)).astype(np.float32)
    return out


def inv_affine_coeffs(center, scale, w, h):
    """Closed form of get_affine_transform(center, scale, 0, [w,h], inv=1) with the reference's
    fp32 roundings of the control points (transforms.py:65-97).  Returns fp64 (a11,a22,b1,b2).
    scale[:,1] is unused by the reference: both axes use the width ratio."""
    center = np.asarray(center, np.float32)
    scale = np.asarray(scale, np.float32)
    cx, cy = center[:, 0], center[:, 1]
    sw = (scale[:, 0] * np.float32(200.0)).astype(np.float32)
    d = (sw * np.float32(-0.5)).astype(np.float32)
    q1y = (cy.astype(np.float64) + d.astype(np.float64)).astype(np.float32)
    dd = (cy - q1y).astype(np.float32)
    q2x = (cx - dd).astype(np.float32)
    half_w = np.float64(w) * 0.5
    a11 = (cx.astype(np.float64) - q2x.astype(np.float64)) / half_w
    a22 = (cy.astype(np.float64) - q1y.astype(np.float64)) / half_w
    b1 = cx.astype(np.float64) - a11 * half_w
    b2 = cy.astype(np.float64) - a22 * (np.float64(h) * 0.5)
    return a11, a22, b1, b2


def get_final_preds(post_process, hm, center, scale):
    """inference.py:52-82.  Returns (preds f32 [N,K,2] image px, maxvals f32 [N,K,1])."""
    coords, maxvals = get_max_preds(hm)
    n, k, h, w = hm.shape
    if post_process:
        coords = quarter_offset(hm, coords)
    a11, a22, b1, b2 = inv_affine_coeffs(center, scale, w, h)
    x = coords[:, :, 0].astype(np.float64)
    y = coords[:, :, 1].astype(np.float64)
    preds = np.empty((n, k, 2), np.float32)
    preds[:, :, 0] = (a11[:, None] * x + b1[:, None]).astype(np.float32)
    preds[:, :, 1] = (a22[:, None] * y + b2[:, None]).astype(np.float32)
    return preds, maxvals


def heatmap_coords(post_process, hm):
    """Heat-map-space coordinates (before the affine) -- the bit-exact part of the contract."""
    coords, maxvals = get_max_preds(hm)
    if post_process:
        coords = quarter_offset(hm, coords)
    return coords, maxvals

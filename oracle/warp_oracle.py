"""CPU oracle for the input side of the path: person-crop affine warp + normalisation.
*** TEST INFRASTRUCTURE ONLY ***

NumPy restatement of what the reference's data loader does per crop
(``lib/dataset/CPJointsDataset.py:1281-1290`` -> ``lib/utils/transforms.py:65-97`` ``get_affine_transform``,
``cv2.warpAffine(img, trans, (W, H), flags=cv2.INTER_LINEAR)``; ``tools/cp_test.py:107-115``
``transforms.Compose([ToTensor(), Normalize(mean, std)])``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it.

Third-party arithmetic: the warp itself lives in OpenCV (unpinned by the reference's requirements.txt
``opencv-python``; 4.13.0 in the authoring container), ``ToTensor`` / ``Normalize`` in torchvision.  Restated here
from OpenCV's published algorithm for 8-bit INTER_LINEAR / BORDER_CONSTANT(0) (imgproc ``warpAffine`` ->
``WarpAffineInvoker`` -> ``remap`` / ``remapBilinear`` with ``FixedPtCast<int, uchar, 15>``):

  * the 2x3 matrix is inverted in fp64 exactly as ``warpAffine`` does (D = 1/det, no fused multiply-adds);
  * source coordinates are fixed point with 10 fractional bits:
    ``X = (round((M01*y + M02)*1024) + 16 + round(M00*x*1024)) >> 5`` (round = half-to-even, ``cvRound``), the integer
    part ``X >> 5`` is clamped to int16 and the 5-bit fraction selects one of 32 x 32 bilinear weight sets;
  * weights are 15-bit fixed point, ``(32-fy)(32-fx)*32`` etc. (exact), except the (0,0) entry: 32768 saturates to
    32767 and OpenCV's sum fix-up -- whose scan for ksize = 2 starts at the LAST tap -- puts the missing 1 on the
    diagonal tap: (32767, 0, 0, 1);
  * taps outside the image read 0; ``dst = (sum + 16384) >> 15``.

Pinned: ``tests/golden/warp_cases.npz`` holds crops produced by the UNMODIFIED reference functions + cv2 + torchvision
in the authoring container (``oracle/gen_golden.py warp``); ``tests/test_oracle_golden.py`` checks this file
against them bit for bit (and, when cv2 is importable, against ``cv2.warpAffine`` on fresh random inputs).
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)          # tools/cp_test.py:107-109
STD = (0.229, 0.224, 0.225)


def get_affine_transform(center, scale, rot, output_size, shift=(0.0, 0.0), inv=0):
    """transforms.py:65-97 with its fp32 roundings of the three control points; the 3-point system that
    ``cv2.getAffineTransform`` solves by LU in fp64 is solved in closed form in fp64 (agreement ~1e-12)."""
    scale = np.asarray(scale, np.float32).reshape(-1)
    if scale.size == 1:
        scale = np.array([scale[0], scale[0]], np.float32)
    center = np.asarray(center, np.float32)
    shift = np.asarray(shift, np.float32)
    scale_tmp = scale * np.float32(200.0)
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]            # get_dir, transforms.py:110-118
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir

    def third(a, b):                                                    # get_3rd_point, transforms.py:105-107
        d = a - b
        return b + np.array([-d[1], d[0]], np.float32)
    src[2, :] = third(src[0, :], src[1, :])
    dst[2, :] = third(dst[0, :], dst[1, :])
    if inv:
        src, dst = dst, src
    a = np.concatenate([src.astype(np.float64), np.ones((3, 1))], axis=1)       # [x y 1] @ T^T = dst
    t = np.linalg.solve(a, dst.astype(np.float64))
    return np.ascontiguousarray(t.T)


def invert_like_cv2(m):
    """warpAffine's in-place inversion of the forward 2x3 matrix (fp64, separate multiplies and adds)."""
    m = np.asarray(m, np.float64).reshape(2, 3)
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m00, m01, m10, m11 = a11, m[0, 1] * -d, m[1, 0] * -d, a22
    b1 = -m00 * m[0, 2] - m01 * m[1, 2]
    b2 = -m10 * m[0, 2] - m11 * m[1, 2]
    return np.array([[m00, m01, b1], [m10, m11, b2]], np.float64)


def bilinear_weights(fy, fx):
    """int64 [..., 4] = (w00, w01, w10, w11), 15-bit fixed point, including the (0,0) quirk."""
    fy = np.asarray(fy, np.int64)
    fx = np.asarray(fx, np.int64)
    w = np.stack([(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32], axis=-1)
    z = (fy == 0) & (fx == 0)
    w[z] = (32767, 0, 0, 1)
    return w


def warp_affine_u8(img, m, dsize):
    """``cv2.warpAffine(img, m, dsize, flags=cv2.INTER_LINEAR)`` for uint8 HxWxC images, bit for bit."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    w_out, h_out = int(dsize[0]), int(dsize[1])
    sh, sw = img.shape[:2]
    mi = invert_like_cv2(m)
    xs = np.arange(w_out, dtype=np.float64)
    ys = np.arange(h_out, dtype=np.float64)
    adelta = np.rint(mi[0, 0] * xs * 1024.0).astype(np.int64)
    bdelta = np.rint(mi[1, 0] * xs * 1024.0).astype(np.int64)
    x0 = np.rint((mi[0, 1] * ys + mi[0, 2]) * 1024.0).astype(np.int64) + 16
    y0 = np.rint((mi[1, 1] * ys + mi[1, 2]) * 1024.0).astype(np.int64) + 16
    X = (x0[:, None] + adelta[None, :]) >> 5
    Y = (y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    wts = bilinear_weights(Y & 31, X & 31)
    acc = np.zeros((h_out, w_out, img.shape[2]), np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = img[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        acc += wts[..., k][..., None] * np.where(ok[..., None], v, 0)
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def crop(img, center, scale, output_size, rot=0):
    """transforms.py:121-129."""
    return warp_affine_u8(img, get_affine_transform(center, scale, rot, output_size),
                          (int(output_size[0]), int(output_size[1])))


def normalize_lut(mean=MEAN, std=STD):
    """f32 [3,256]: ToTensor (u/255 in fp32) followed by Normalize ((t - mean)/std in fp32)."""
    u = np.arange(256, dtype=np.float32) / np.float32(255.0)
    m = np.asarray(mean, np.float32)[:, None]
    s = np.asarray(std, np.float32)[:, None]
    return ((u[None, :] - m) / s).astype(np.float32)


def to_tensor_normalize(img_u8, mean=MEAN, std=STD):
    """uint8 HWC -> f32 CHW, ``Compose([ToTensor(), Normalize(mean, std)])``."""
    lut = normalize_lut(mean, std)
    return np.stack([lut[c][img_u8[:, :, c]] for c in range(3)]).astype(np.float32)


def crop_input(img_bgr, center, scale, image_size, rot=0, color_rgb=True, mean=MEAN, std=STD):
    """One loader sample (CPJointsDataset.py:1228-1232, 1281-1290): optional BGR->RGB, warp, normalise."""
    img = img_bgr[:, :, ::-1] if color_rgb else img_bgr
    return to_tensor_normalize(crop(np.ascontiguousarray(img), center, scale, image_size, rot), mean, std)

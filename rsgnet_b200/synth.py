"""Seeded synthetic workloads: crops, heat-maps, centres/scales, detections.

NumPy's legacy ``RandomState`` streams are stable across platforms and versions, so the same
seed gives the same bytes here, on the GPU box and in the committed golden fixtures.
"""
import numpy as np


def crops(n, image_wh, seed=0):
    """fp32 NCHW ~N(0,1) (ImageNet-normalised images look like this)."""
    rs = np.random.RandomState(seed)
    return rs.standard_normal((n, 3, image_wh[1], image_wh[0])).astype(np.float32)


def centers_scales(n, seed=0):
    rs = np.random.RandomState(seed + 1000)
    c = rs.uniform(20, 600, (n, 2)).astype(np.float32)
    s = rs.uniform(0.3, 4.0, (n, 2)).astype(np.float32)
    return c, s


def heatmaps(n, k, h, w, seed=2, dead_frac=0.15, noise=0.01):
    """Gaussian bump (sigma 2, centre anywhere incl. borders, amplitude U(0.1,1)) + N(0,noise);
    `dead_frac` of the maps get 2.0 subtracted so that their max <= 0."""
    rs = np.random.RandomState(seed)
    cx = rs.uniform(-1, w, (n, k, 1, 1)).astype(np.float32)
    cy = rs.uniform(-1, h, (n, k, 1, 1)).astype(np.float32)
    amp = rs.uniform(0.1, 1.0, (n, k, 1, 1)).astype(np.float32)
    xs = np.arange(w, dtype=np.float32)[None, None, None, :]
    ys = np.arange(h, dtype=np.float32)[None, None, :, None]
    hm = amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / np.float32(8.0))
    hm = hm + rs.normal(0, noise, hm.shape).astype(np.float32)
    dead = rs.uniform(size=(n, k, 1, 1)) < dead_frac
    hm = np.where(dead, hm - np.float32(2.0), hm)
    return np.ascontiguousarray(hm.astype(np.float32))


def crafted_heatmaps(k, h, w):
    """Edge cases of App. B.1 as one [C,k,h,w] batch: constant map (tie -> index 0), all-negative,
    all-zero, maxima at the four corners and at px in {1,2,W-3,W-2}, equal neighbours."""
    cases = []

    def blank(v=0.0):
        return np.full((k, h, w), v, np.float32)
    cases.append(blank(0.5))
    cases.append(blank(-1.0))
    cases.append(blank(0.0))
    for (y, x) in ((0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1)):
        m = blank(0.01)
        m[:, y, x] = 1.0
        cases.append(m)
    for x in (1, 2, w - 3, w - 2):
        for y in (1, 2, h - 3, h - 2):
            m = blank(0.0)
            m[:, y, x] = 1.0
            m[:, y, min(x + 1, w - 1)] = 0.5
            m[:, min(y + 1, h - 1), x] = 0.25
            cases.append(m)
    m = blank(0.0)                       # symmetric neighbours -> sign(0) = 0
    m[:, h // 2, w // 2] = 1.0
    m[:, h // 2, w // 2 - 1] = m[:, h // 2, w // 2 + 1] = 0.3
    m[:, h // 2 - 1, w // 2] = m[:, h // 2 + 1, w // 2] = 0.3
    cases.append(m)
    m = blank(0.0)                       # duplicated maximum -> first occurrence wins
    m[:, 3, 5] = 0.7
    m[:, 2, w - 4] = 0.7
    cases.append(m)
    return np.stack(cases).astype(np.float32)


def detections(n_imgs, per_img, k, seed=3, ragged=False):
    """Pose detections grouped per image: 3 pose clusters per image with jitter
    sigma in {1,5,30} px so that NMS both keeps and suppresses; distinct scores.
    Returns kpts f32 [N,k,3], scores f64 [N], areas f64 [N], offsets i32 [n_imgs+1]."""
    rs = np.random.RandomState(seed)
    counts = (rs.randint(0, 2 * per_img + 1, n_imgs) if ragged
              else np.full(n_imgs, per_img)).astype(np.int64)
    offsets = np.zeros(n_imgs + 1, np.int32)
    offsets[1:] = np.cumsum(counts)
    n = int(offsets[-1])
    kpts = np.zeros((n, k, 3), np.float32)
    for i in range(n_imgs):
        c = int(counts[i])
        if c == 0:
            continue
        base = rs.uniform(50, 500, (3, k, 2))
        which = rs.randint(0, 3, c)
        sig = np.array([1.0, 5.0, 30.0])[rs.randint(0, 3, c)]
        pts = base[which] + rs.normal(0, 1, (c, k, 2)) * sig[:, None, None]
        kpts[offsets[i]:offsets[i + 1], :, :2] = pts.astype(np.float32)
    kpts[:, :, 2] = rs.uniform(0, 1, (n, k)).astype(np.float32)
    scores = rs.permutation(n).astype(np.float64) / max(n, 1) + rs.uniform(0, 0.5 / max(n, 1), n)
    areas = rs.uniform(2e3, 4e4, n).astype(np.float32).astype(np.float64)
    return kpts, scores, areas, offsets


def evaluate_inputs(n_imgs, per_img, k, seed=13, ragged=True, zero_frac=0.03):
    """What rsgnet_validate hands to dataset.evaluate() (function.py:376-380, 452-458, 479): all_preds f32 [N,k,3]
    (image-space x, y, maxval), all_boxes f64 [N,6] (centre, scale, area = prod(scale*200) computed in fp32, box score)
    and one image id per detection -- with the detections of an image NOT contiguous (the loader order is arbitrary), a
    few detections whose maxvals all sit below IN_VIS_THRE (rescored to exactly 0 -- at most ONE per image: the order
    NumPy's argsort gives to exactly equal scores is implementation-defined (introsort / SIMD sorting networks), so
    a fixture with score ties would pin an artefact of one NumPy build) and maxvals sitting exactly on the threshold.
    all_boxes[:, 0] carries the detection's own index (the centre is unused by rescoring and NMS), so that the
    reference's output dicts can be mapped back."""
    kpts, _, _, off = detections(n_imgs, per_img, k, seed=seed, ragged=ragged)
    rs = np.random.RandomState(seed + 500)
    n = kpts.shape[0]
    img_of = np.repeat(np.arange(n_imgs), np.diff(off))
    counts = np.diff(off)
    mv = rs.uniform(0.0, 1.0, (n, k)).astype(np.float32)
    mv[rs.uniform(size=(n, k)) < 0.1] = np.float32(0.2)               # exactly at the threshold: not visible (strict >)
    zero = rs.uniform(size=n) < zero_frac
    zero &= np.concatenate([[True], img_of[1:] != img_of[:-1]])        # only an image's first detection: no score ties
    mv[zero] = rs.uniform(0.0, 0.19, (int(zero.sum()), k)).astype(np.float32)
    kpts[:, :, 2] = mv
    perm = rs.permutation(n)                                          # interleave the images
    kpts, img_of = kpts[perm], img_of[perm]
    sc = rs.uniform(0.3, 4.0, (n, 2)).astype(np.float32)
    boxes = np.zeros((n, 6))
    boxes[:, 0] = np.arange(n)
    boxes[:, 1] = rs.uniform(20, 600, n).astype(np.float32)
    boxes[:, 2:4] = sc
    boxes[:, 4] = np.prod(sc * 200, 1)
    boxes[:, 5] = rs.uniform(0.05, 1.0, n)
    image_ids = (100000 + 7 * img_of).astype(np.int64)
    return np.ascontiguousarray(kpts), boxes, image_ids


def images(n, seed=7, hw_range=((120, 480), (160, 640))):
    """n uint8 HWC "photos" of different sizes: smooth colour waves + blobs + noise (so that bilinear taps differ and
    the fixtures compress), plus the person boxes (centre, scale) the loader would crop: some far inside, some
    hanging over every border, tiny and huge scales."""
    rs = np.random.RandomState(seed)
    imgs, centers, scales = [], [], []
    for _ in range(n):
        h = int(rs.randint(*hw_range[0]))
        w = int(rs.randint(*hw_range[1]))
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.zeros((h, w, 3), np.float32)
        for c in range(3):
            fx, fy, ph = rs.uniform(0.01, 0.15), rs.uniform(0.01, 0.15), rs.uniform(0, 6.28)
            img[:, :, c] = 127 + 90 * np.sin(fx * xx + fy * yy + ph) + rs.normal(0, 12, (h, w))
        imgs.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
        centers.append([rs.uniform(-0.1, 1.1) * w, rs.uniform(-0.1, 1.1) * h])
        s = rs.uniform(0.15, 3.5)
        scales.append([s, s * 1.25])
    return imgs, np.asarray(centers, np.float32), np.asarray(scales, np.float32)


def train_batch(n, image_wh, heat_wh, num_joints, num_limbs, seed=0):
    """One synthetic training batch in the layout of the reference's loader (lib/core/function.py:258: input, target,
    target_weight, all_ins_target, all_ins_target_weight, target_limbs): Gaussian joint maps (sigma 2) of a target person,
    the same plus a second person for the 'all instances' maps, 0/1 joint weights, smooth limb maps in (0, 1)."""
    rs = np.random.RandomState(seed + 77)
    w, h = heat_wh
    xs = np.arange(w, dtype=np.float32)[None, None, None, :]
    ys = np.arange(h, dtype=np.float32)[None, None, :, None]

    def bumps(k):
        cx = rs.uniform(2, w - 2, (n, k, 1, 1)).astype(np.float32)
        cy = rs.uniform(2, h - 2, (n, k, 1, 1)).astype(np.float32)
        return np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / np.float32(8.0)).astype(np.float32)

    target = bumps(num_joints)
    tw = (rs.uniform(size=(n, num_joints, 1)) > 0.2).astype(np.float32)
    target = target * tw[..., None]
    other = bumps(num_joints)
    all_target = np.maximum(target, other).astype(np.float32)
    all_tw = np.maximum(tw, (rs.uniform(size=(n, num_joints, 1)) > 0.3).astype(np.float32))
    limbs = np.clip(0.5 * bumps(num_limbs) + 0.5 * bumps(num_limbs), 0.0, 1.0).astype(np.float32)
    x = rs.standard_normal((n, 3, image_wh[1], image_wh[0])).astype(np.float32)
    c = np.ascontiguousarray
    return dict(input=c(x), target=c(target), target_weight=c(tw), all_ins_target=c(all_target),
                all_ins_target_weight=c(all_tw), target_limbs=c(limbs))

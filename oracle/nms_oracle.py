"""CPU oracle for OKS-NMS and the evaluate()-side rescoring.   *** TEST INFRASTRUCTURE ONLY ***

NumPy restatement; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import it.  Pinned by ``tests/golden/nms_*.npz`` produced by the
unmodified reference's ``oks_nms`` (``oracle/gen_golden.py``).

Reference lines followed (paths under /root/reference):
  lib/nms/nms.py:75-94            oks_iou (dx,dy,squares in fp32; the rest fp64; np.sum order)
  lib/nms/nms.py:97-124           oks_nms greedy sweep (keep oks <= thresh)
  lib/nms/nms.py:127-180          rescore ('gaussian') + soft_oks_nms (at most 20 detections)
  lib/dataset/crowdpose.py:1294-1306   rescoring: box_score * mean(maxval > in_vis_thre)
"""
import numpy as np

COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07,
                        .87, .87, .89, .89]) / 10.0
CROWDPOSE_SIGMAS = np.array([.79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89,
                             .35, .35]) / 10.0


def oks_iou(g, d, a_g, a_d, sigmas=None, in_vis_thre=None):
    """nms.py:75-94.  g: f32[3K]; d: f32[M,3K]; areas fp64."""
    if not isinstance(sigmas, np.ndarray):
        sigmas = COCO_SIGMAS
    var = (sigmas * 2) ** 2
    xg, yg, vg = g[0::3], g[1::3], g[2::3]
    ious = np.zeros(d.shape[0])
    for i in range(d.shape[0]):
        dx = d[i, 0::3] - xg
        dy = d[i, 1::3] - yg
        e = (dx ** 2 + dy ** 2) / var / ((a_g + a_d[i]) / 2 + np.spacing(1)) / 2
        if in_vis_thre is not None:
            # the reference evaluates `list(a) and list(b)`, i.e. the SECOND list only
            e = e[list(d[i, 2::3] > in_vis_thre)]
        ious[i] = np.sum(np.exp(-e)) / e.shape[0] if e.shape[0] != 0 else 0.0
    return ious


def oks_order(scores):
    """Descending order used by the sweep: reverse of NumPy's ascending argsort (nms.py:112)."""
    return np.asarray(scores).argsort()[::-1]


def oks_nms(kpts_db, thresh, sigmas=None, in_vis_thre=None):
    """nms.py:97-124.  Returns indices in greedy selection order."""
    if len(kpts_db) == 0:
        return []
    scores = np.array([k['score'] for k in kpts_db])
    kpts = np.array([k['keypoints'].flatten() for k in kpts_db])
    areas = np.array([k['area'] for k in kpts_db])
    order = oks_order(scores)
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, in_vis_thre)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


def oks_nms_arrays(kpts, scores, areas, thresh, sigmas=None):
    """Same sweep on plain arrays (kpts f32 [n,K,3]); also returns the minimum |oks-thresh|
    margin seen, so tests can tell whether a 1-ulp exp difference could flip a decision."""
    n = kpts.shape[0]
    if n == 0:
        return [], np.inf
    flat = kpts.reshape(n, -1)
    order = oks_order(scores)
    keep, margin = [], np.inf
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        ovr = oks_iou(flat[i], flat[order[1:]], areas[i], areas[order[1:]], sigmas)
        if ovr.size:
            margin = min(margin, float(np.min(np.abs(ovr - thresh))))
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep, margin


def rescore(box_score, maxvals, in_vis_thre):
    """crowdpose.py:1294-1306 / coco.py:1249-1261 for one detection.
    maxvals: fp32 [K]; accumulation runs in fp32 (np.float32 scalars), product with the fp64
    box score in fp64."""
    kpt_score = 0
    valid = 0
    for t in maxvals:
        if t > in_vis_thre:
            kpt_score = kpt_score + t
            valid += 1
    if valid != 0:
        kpt_score = kpt_score / valid
    return kpt_score * box_score


def soft_oks_nms(kpts_db, thresh, sigmas=None, in_vis_thre=None):
    """nms.py:138-180 (with rescore(), nms.py:127-135, type 'gaussian')."""
    if len(kpts_db) == 0:
        return []
    scores = np.array([k['score'] for k in kpts_db])
    kpts = np.array([k['keypoints'].flatten() for k in kpts_db])
    areas = np.array([k['area'] for k in kpts_db])
    order = scores.argsort()[::-1]
    scores = scores[order]
    max_dets = 20
    keep = np.zeros(max_dets, dtype=np.intp)
    keep_cnt = 0
    while order.size > 0 and keep_cnt < max_dets:
        i = order[0]
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, in_vis_thre)
        order = order[1:]
        scores = scores[1:] * np.exp(-ovr ** 2 / thresh)
        tmp = scores.argsort()[::-1]
        order = order[tmp]
        scores = scores[tmp]
        keep[keep_cnt] = i
        keep_cnt += 1
    return keep[:keep_cnt]


def evaluate(all_preds, all_boxes, image_ids, sigmas, in_vis_thre, oks_thre, soft_nms=False):
    """dataset.evaluate() up to the kept lists (crowdpose.py:1272-1324 / coco.py:1227-1277): group the detections by
    image in first-appearance order, rescore each one (box score * mean of the maxvals above in_vis_thre), run
    (soft-)OKS-NMS per image WITHOUT in_vis_thre (no caller passes it), keep everything if NMS keeps nothing.
    all_preds f32 [N,K,3]; all_boxes f64 [N,6] (.., area at 4, box score at 5); image_ids int [N].
    Returns (images i64 [n_imgs], counts i32 [n_imgs], keep i32 [sum counts] GLOBAL detection indices in the
    per-image selection order, scores f64 [sum counts] rescored)."""
    groups = {}
    for idx, img in enumerate(image_ids):
        groups.setdefault(int(img), []).append(idx)
    images, counts, keep_all, scores_all = [], [], [], []
    for img, members in groups.items():
        db = []
        for idx in members:
            db.append(dict(keypoints=all_preds[idx], area=all_boxes[idx][4],
                           score=rescore(all_boxes[idx][5], all_preds[idx][:, 2], in_vis_thre)))
        keep = (soft_oks_nms if soft_nms else oks_nms)(db, oks_thre, sigmas)
        if len(keep) == 0:
            keep = list(range(len(db)))
        images.append(img)
        counts.append(len(keep))
        keep_all.extend(members[int(j)] for j in keep)
        scores_all.extend(db[int(j)]['score'] for j in keep)
    return (np.asarray(images, np.int64), np.asarray(counts, np.int32), np.asarray(keep_all, np.int32),
            np.asarray(scores_all, np.float64))

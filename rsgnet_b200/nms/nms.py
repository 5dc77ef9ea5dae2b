"""Drop-in for the pose part of /root/reference/lib/nms/nms.py.

``oks_nms(kpts_db, thresh, sigmas=None, in_vis_thre=None)`` keeps the reference's signature and
return convention (nms.py:97-124: Python list of indices in greedy selection order; [] for an empty
input).  ``oks_nms_batched`` is the form the device kernel is built for: all images of an
``evaluate()`` call in one segmented launch (crowdpose.py:1283-1324).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07,
                        .87, .87, .89, .89]) / 10.0


def _p(t):
    return C.c_void_p(t.data_ptr())


def _vis(in_vis_thre):
    """(use_in_vis_thre, in_vis_thre) of the C ABI for the reference's `in_vis_thre=None | float`."""
    return (0, 0.0) if in_vis_thre is None else (1, float(in_vis_thre))


def oks_nms_batched(kpts, scores, areas, img_offsets, thresh, sigmas=None, device=None, in_vis_thre=None):
    """kpts f32 [n,K,3]; scores, areas f64 [n]; img_offsets i32 [n_imgs+1] (NumPy or CUDA tensors).
    Returns (keep i32 [n] -- per image, at keep[off[i]:off[i]+count[i]], indices relative to the
    image in selection order --, counts i32 [n_imgs]) as NumPy arrays."""
    _lib.require_cuda()
    device = torch.device(device or 'cuda')
    off_np = np.ascontiguousarray(img_offsets.cpu().numpy() if isinstance(img_offsets, torch.Tensor)
                                  else img_offsets, dtype=np.int32)
    n_imgs = len(off_np) - 1
    n = int(off_np[-1]) if n_imgs >= 0 else 0
    if n_imgs <= 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    max_per = int(np.max(np.diff(off_np))) if n_imgs else 0
    to = lambda a, dt: (a.to(device, dt).contiguous() if isinstance(a, torch.Tensor)
                        else torch.from_numpy(np.ascontiguousarray(a, dtype=dt_np[dt])).to(device))
    dt_np = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32}
    k = to(kpts, torch.float32)
    K = int(k.shape[1]) if k.dim() == 3 else int(k.shape[-1]) // 3
    s = to(scores, torch.float64)
    a = to(areas, torch.float64)
    off = torch.from_numpy(off_np).to(device)
    if sigmas is None or not isinstance(sigmas, np.ndarray):
        sigmas = COCO_SIGMAS
    assert len(sigmas) == K, f'{len(sigmas)} sigmas for K={K} key points'
    sg = torch.from_numpy(np.ascontiguousarray(sigmas, np.float64)).to(device)
    keep = torch.full((max(n, 1),), -1, dtype=torch.int32, device=device)
    counts = torch.zeros(n_imgs, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().rsg_oks_nms(_lib.stream_ptr(device), _p(k), _p(s), _p(a), _p(off),
                                          n_imgs, max_per, _p(sg), K, float(thresh), _p(keep),
                                          _p(counts), *_vis(in_vis_thre)))
    return keep[:n].cpu().numpy(), counts.cpu().numpy()


def oks_iou(g, d, a_g, a_d, sigmas=None, in_vis_thre=None):
    """nms.py:75-94: OKS of detection g (f32 [3K]) with each row of d (f32 [M,3K]); returns f64 [M]."""
    _lib.require_cuda()
    if not isinstance(sigmas, np.ndarray):
        sigmas = COCO_SIGMAS
    dev = torch.device('cuda')
    gt = torch.from_numpy(np.ascontiguousarray(g, np.float32).reshape(-1)).to(dev)
    dt = torch.from_numpy(np.ascontiguousarray(d, np.float32).reshape(len(d), -1)).to(dev)
    at = torch.from_numpy(np.ascontiguousarray(a_d, np.float64).reshape(-1)).to(dev)
    sg = torch.from_numpy(np.ascontiguousarray(sigmas, np.float64)).to(dev)
    K, M = gt.numel() // 3, dt.shape[0]
    assert len(sigmas) == K
    out = torch.zeros(M, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().rsg_oks_iou(_lib.stream_ptr(dev), _p(gt), _p(dt), float(a_g), _p(at), _p(sg), K, M, _p(out),
                                      *_vis(in_vis_thre)))
    return out.cpu().numpy()


def oks_nms(kpts_db, thresh, sigmas=None, in_vis_thre=None):
    """nms.py:97-124.  `in_vis_thre` (never passed by the reference's own callers, crowdpose.py:1315, coco.py:1269)
    follows the reference's oks_iou: only the compared detection's key points with score > in_vis_thre count."""
    if len(kpts_db) == 0:
        return []
    scores = np.array([kpts_db[i]['score'] for i in range(len(kpts_db))], np.float64)
    kpts = np.array([np.asarray(kpts_db[i]['keypoints'], np.float32).reshape(-1, 3)
                     for i in range(len(kpts_db))], np.float32)
    areas = np.array([kpts_db[i]['area'] for i in range(len(kpts_db))], np.float64)
    keep, counts = oks_nms_batched(kpts, scores, areas, np.array([0, len(kpts_db)], np.int32),
                                   thresh, sigmas, in_vis_thre=in_vis_thre)
    return [int(v) for v in keep[:int(counts[0])]]


def rescore(maxvals, box_scores, in_vis_thre, device=None):
    """crowdpose.py:1294-1306 for all detections at once: box_score * mean(maxvals > in_vis_thre).
    maxvals f32 [n,K(,1)], box_scores f64 [n] -> f64 [n] (NumPy)."""
    _lib.require_cuda()
    device = torch.device(device or 'cuda')
    mv = (maxvals if isinstance(maxvals, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(maxvals, np.float32)))
    mv = mv.to(device, torch.float32).reshape(mv.shape[0], -1).contiguous()
    bs = (box_scores if isinstance(box_scores, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(box_scores, np.float64)))
    bs = bs.to(device, torch.float64).contiguous()
    out = torch.empty(mv.shape[0], dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().rsg_rescore(_lib.stream_ptr(device), _p(mv), _p(bs), mv.shape[0],
                                          mv.shape[1], float(in_vis_thre), _p(out)))
    return out.cpu().numpy()


def soft_oks_nms_batched(kpts, scores, areas, img_offsets, thresh, sigmas=None, max_dets=20, device=None, in_vis_thre=None):
    """Segmented soft OKS-NMS (nms.py:138-180 for all images of an evaluate() call): returns (keep i32 [n_imgs, max_dets]
    -- indices relative to the image, in selection order --, counts i32 [n_imgs]) as NumPy arrays."""
    _lib.require_cuda()
    device = torch.device(device or 'cuda')
    off_np = np.ascontiguousarray(img_offsets, dtype=np.int32)
    n_imgs = len(off_np) - 1
    if n_imgs <= 0:
        return np.zeros((0, max_dets), np.int32), np.zeros(0, np.int32)
    max_per = int(np.max(np.diff(off_np)))
    k = torch.from_numpy(np.ascontiguousarray(kpts, np.float32)).to(device)
    K = int(k.shape[1]) if k.dim() == 3 else int(k.shape[-1]) // 3
    s = torch.from_numpy(np.ascontiguousarray(scores, np.float64)).to(device)
    a = torch.from_numpy(np.ascontiguousarray(areas, np.float64)).to(device)
    off = torch.from_numpy(off_np).to(device)
    if sigmas is None or not isinstance(sigmas, np.ndarray):
        sigmas = COCO_SIGMAS
    assert len(sigmas) == K, f'{len(sigmas)} sigmas for K={K} key points'
    sg = torch.from_numpy(np.ascontiguousarray(sigmas, np.float64)).to(device)
    keep = torch.full((n_imgs, max_dets), -1, dtype=torch.int32, device=device)
    counts = torch.zeros(n_imgs, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().rsg_soft_oks_nms(_lib.stream_ptr(device), _p(k), _p(s), _p(a), _p(off), n_imgs, max_per,
                                               _p(sg), K, float(thresh), int(max_dets), _p(keep), _p(counts),
                                               *_vis(in_vis_thre)))
    return keep.cpu().numpy(), counts.cpu().numpy()


def soft_oks_nms(kpts_db, thresh, sigmas=None, in_vis_thre=None):
    """nms.py:138-180: at most 20 detections, chosen by repeatedly taking the best score and decaying the others by
    exp(-oks^2 / thresh).  Returns an integer array of indices like the reference (``keep[:keep_cnt]``)."""
    if len(kpts_db) == 0:
        return []
    scores = np.array([kpts_db[i]['score'] for i in range(len(kpts_db))], np.float64)
    kpts = np.array([np.asarray(kpts_db[i]['keypoints'], np.float32).reshape(-1, 3)
                     for i in range(len(kpts_db))], np.float32)
    areas = np.array([kpts_db[i]['area'] for i in range(len(kpts_db))], np.float64)
    keep, counts = soft_oks_nms_batched(kpts, scores, areas, np.array([0, len(kpts_db)], np.int32), thresh, sigmas,
                                        in_vis_thre=in_vis_thre)
    return keep[0, :int(counts[0])].astype(np.intp)

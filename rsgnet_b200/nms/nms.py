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
    counts_np = counts.cpu().numpy()
    if (counts_np < 0).any():
        raise _lib.RsgError('oks_nms: an image holds more detections than max_per_img')
    return keep[:n].cpu().numpy(), counts_np


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
    counts_np = counts.cpu().numpy()
    if (counts_np < 0).any():
        raise _lib.RsgError('soft_oks_nms: an image holds more detections than max_per_img')
    return keep.cpu().numpy(), counts_np


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


class EvalResult:
    """Result of evaluate_device: everything lives in ONE device buffer; `.host()` is the single D2H copy."""

    def __init__(self, buf, n, layout):
        self.buf, self.n, self._layout = buf, n, layout

    def _view(self, host, name):
        off, count, dt = self._layout[name]
        return host[off:off + count * np.dtype(dt).itemsize].view(dt)

    def device(self, name):
        """CUDA tensor view of one output ('n_imgs', 'images', 'img_offsets', 'scores', 'keep', 'keep_counts')."""
        off, count, dt = self._layout[name]
        tdt = {np.int32: torch.int32, np.int64: torch.int64, np.float64: torch.float64}[dt]
        return self.buf[off:off + count * np.dtype(dt).itemsize].view(tdt)

    def host(self):
        """One D2H copy -> dict(images i64 [n_imgs], counts i32 [n_imgs], keep i32 [sum counts] (global detection indices,
        per image in selection order, images in first-appearance order), keep_offsets i32 [n_imgs+1] into `keep`,
        scores f64 [n] (rescored, every detection))."""
        h = self.buf.cpu().numpy()
        n_imgs = int(self._view(h, 'n_imgs')[0])
        if n_imgs < 0:
            raise _lib.RsgError('evaluate_device: an image id equals the reserved value 0x8080808080808080')
        off = self._view(h, 'img_offsets')[:n_imgs + 1]
        counts = self._view(h, 'keep_counts')[:n_imgs].copy()
        keep_all = self._view(h, 'keep')
        keep = np.concatenate([keep_all[off[r]:off[r] + counts[r]] for r in range(n_imgs)]) if n_imgs else np.zeros(0, np.int32)
        ko = np.zeros(n_imgs + 1, np.int32)
        ko[1:] = np.cumsum(counts)
        return dict(images=self._view(h, 'images')[:n_imgs].copy(), counts=counts, keep=keep.astype(np.int32), keep_offsets=ko,
                    scores=self._view(h, 'scores')[:self.n].copy())


def evaluate_device(all_preds, all_boxes, image_ids, oks_thre, in_vis_thre, sigmas=None, soft_nms=False, max_dets=20,
                    device=None):
    """dataset.evaluate() up to the kept lists, on the device in one call (crowdpose.py:1272-1324, coco.py:1227-1277):
    group by image (first-appearance order) -> rescore -> (soft-)OKS-NMS per image -> compacted keep lists.
    all_preds f32 [N,K,3] (x, y, maxval -- what function.py:452-453 accumulates; a CUDA tensor stays on the device),
    all_boxes f64 [N,6] (area at 4, box score at 5), image_ids int64 [N] in any order.  Returns an EvalResult."""
    _lib.require_cuda()
    device = torch.device(device or (all_preds.device if isinstance(all_preds, torch.Tensor) and all_preds.is_cuda else 'cuda'))

    def to(a, tdt, ndt):
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=ndt))
        return t.to(device, tdt).contiguous()
    p = to(all_preds, torch.float32, np.float32)
    b = to(all_boxes, torch.float64, np.float64)
    ids = to(image_ids, torch.int64, np.int64)
    n, K = int(p.shape[0]), int(p.shape[1])
    assert p.dim() == 3 and p.shape[2] == 3 and tuple(b.shape) == (n, 6) and tuple(ids.shape) == (n,)
    if sigmas is None or not isinstance(sigmas, np.ndarray):
        sigmas = COCO_SIGMAS
    assert len(sigmas) == K, f'{len(sigmas)} sigmas for K={K} key points'
    sg = torch.from_numpy(np.ascontiguousarray(sigmas, np.float64)).to(device)
    N = max(n, 1)
    layout, off = {}, 0
    for name, count, dt in (('n_imgs', 2, np.int32), ('images', N, np.int64), ('scores', N, np.float64),
                            ('img_offsets', N + 1, np.int32), ('keep', N, np.int32), ('keep_counts', N, np.int32)):
        layout[name] = (off, count, dt)
        off += (count * np.dtype(dt).itemsize + 7) // 8 * 8
    buf = torch.zeros(off, dtype=torch.uint8, device=device)
    ws_bytes = C.c_size_t()
    _lib.check(_lib.lib().rsg_evaluate_workspace_bytes(n, C.byref(ws_bytes)))
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=device)
    ptr = lambda name: C.c_void_p(buf.data_ptr() + layout[name][0])
    with torch.cuda.device(device):
        _lib.check(_lib.lib().rsg_evaluate(_lib.stream_ptr(device), _p(p), _p(b), _p(ids), n, K, _p(sg), float(in_vis_thre),
                                           float(oks_thre), int(bool(soft_nms)), int(max_dets), _p(ws), ws_bytes.value,
                                           ptr('n_imgs'), ptr('images'), ptr('img_offsets'), ptr('scores'), ptr('keep'),
                                           ptr('keep_counts')))
    res = EvalResult(buf, n, layout)
    res._keepalive = (p, b, ids, sg, ws)
    return res

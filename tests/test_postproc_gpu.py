"""GPU parity of the post-processing kernels through the C ABI: decode / flip-average / flip_back
are bit-exact against the CPU oracle and the reference-executed golden fixtures; OKS-NMS keep lists
equal the reference's."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, nms_oracle
from rsgnet_b200 import presets, synth
from rsgnet_b200.core.inference import decode_device, get_final_preds, get_max_preds
from rsgnet_b200.nms.nms import oks_iou, oks_nms, oks_nms_batched, rescore, soft_oks_nms, soft_oks_nms_batched
from rsgnet_b200.utils.transforms import flip_back, flip_perm

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, np.int64(-2 ** 31) - a, a)
    b = np.where(b < 0, np.int64(-2 ** 31) - b, b)
    return np.abs(a - b)


@pytest.mark.parametrize('tag', ['small', 'hrnet', 'rsgnet'])
def test_decode_vs_golden_and_oracle(golden_dir, tag):
    g = _load(golden_dir, f'decode_{tag}.npz')
    k, h, w, n = int(g['k']), int(g['h']), int(g['w']), int(g['n'])
    hm = np.concatenate([synth.heatmaps(n, k, h, w, seed=int(g['seed'])), synth.crafted_heatmaps(k, h, w)])
    c, s = synth.centers_scales(hm.shape[0], seed=int(g['cs_seed']))
    coords, mv = get_max_preds(hm)
    assert np.array_equal(coords, g['coords_raw'])
    for pp in (0, 1):
        cfg = presets.make_cfg(post_process=bool(pp))
        preds, maxvals = get_final_preds(cfg, hm, c, s)
        assert preds.dtype == np.float32 and preds.shape == (hm.shape[0], k, 2)
        assert maxvals.shape == (hm.shape[0], k, 1)
        assert np.array_equal(maxvals, g[f'maxvals_pp{pp}'])          # reference, bit-exact
        d = ulp_diff(preds, g[f'preds_pp{pp}'])
        assert d.max() <= 1                                            # reference solves the affine by LU
        o_preds, o_mv = decode_oracle.get_final_preds(bool(pp), hm, c, s)
        assert np.array_equal(preds, o_preds) and np.array_equal(maxvals, o_mv)   # oracle, bit-exact
        out = decode_device(torch.from_numpy(hm).cuda(), post_process=bool(pp), want_coords=True)
        o_coords, _ = decode_oracle.heatmap_coords(bool(pp), hm)
        assert np.array_equal(out['coords'].cpu().numpy(), o_coords)


@pytest.mark.parametrize('tag', ['small', 'hrnet', 'rsgnet'])
def test_flip_average_fused(golden_dir, tag):
    g = _load(golden_dir, f'decode_{tag}.npz')
    k, h, w = int(g['k']), int(g['h']), int(g['w'])
    rs = np.random.RandomState(int(g['flip_seed']))
    nb = int(g['flip_n'])
    a = rs.standard_normal((nb, k, h, w)).astype(np.float32)
    b = rs.standard_normal((nb, k, h, w)).astype(np.float32)
    pairs = presets.flip_pairs_for(k)
    fb = flip_back(b, pairs)
    assert np.array_equal(fb, decode_oracle.flip_back(b, pairs))
    if 'flip_back' in g:
        assert np.array_equal(fb, g['flip_back'])
    c, s = synth.centers_scales(nb, seed=7)
    for shift in (True, False):
        out = decode_device(torch.from_numpy(a).cuda(), c, s, post_process=True,
                            hm_flipped=torch.from_numpy(b).cuda(), flip_perm=flip_perm(k, pairs),
                            shift=shift, want_coords=True, want_avg=True)
        avg = decode_oracle.flip_average(a, b, pairs, shift=shift)
        if shift:
            assert np.array_equal(avg, g['flip_avg'])
        assert np.array_equal(out['avg'].cpu().numpy(), avg)
        o_preds, o_mv = decode_oracle.get_final_preds(True, avg, c, s)
        assert np.array_equal(out['preds'].cpu().numpy(), o_preds)
        assert np.array_equal(out['maxvals'].cpu().numpy(), o_mv)


def test_decode_ragged_and_edge_shapes():
    # odd widths (scalar path), single-row maps, N=0
    for (n, k, h, w) in [(3, 5, 7, 9), (2, 1, 3, 33), (1, 17, 130, 6), (0, 17, 64, 48)]:
        hm = synth.heatmaps(n, k, h, w, seed=11) if n else np.zeros((0, k, h, w), np.float32)
        c, s = synth.centers_scales(n, seed=3)
        preds, mv = get_final_preds(presets.make_cfg(), hm, c, s)
        o_preds, o_mv = decode_oracle.get_final_preds(True, hm, c, s) if n else (np.zeros((0, k, 2), np.float32), np.zeros((0, k, 1), np.float32))
        assert np.array_equal(preds, o_preds) and np.array_equal(mv, o_mv)
    with pytest.raises(AssertionError):
        get_final_preds(presets.make_cfg(), np.zeros((4, 17, 64), np.float32), None, None)


def test_decode_full_size_properties():
    """BASELINE.json config 4 scale (a 20k-crop slice): idempotence + argmax property."""
    n, k, h, w = 20000, 17, 64, 48
    g = torch.Generator(device='cuda').manual_seed(5)
    hm = torch.rand((n, k, h, w), device='cuda', generator=g) - 0.1
    out = decode_device(hm, post_process=False, want_coords=True)
    flat = hm.view(n, k, -1)
    mv, idx = flat.max(dim=2)
    assert torch.equal(out['maxvals'].view(n, k), mv)
    xy = out['coords']
    got_idx = (xy[..., 1] * w + xy[..., 0]).long()
    pos = mv > 0
    assert torch.equal(flat.gather(2, got_idx[..., None]).squeeze(2)[pos], mv[pos])
    # first-occurrence rule
    first = (flat == mv[..., None]).float().argmax(dim=2)
    assert torch.equal(got_idx[pos], first[pos])


@pytest.mark.parametrize('tag', ['coco', 'crowdpose'])
def test_oks_nms_vs_reference(golden_dir, tag):
    g = _load(golden_dir, f'nms_{tag}.npz')
    k = int(g['k'])
    sig = None if tag == 'coco' else nms_oracle.CROWDPOSE_SIGMAS
    kpts, scores, areas, off = synth.detections(int(g['n_imgs']), int(g['per_img']), k, seed=int(g['seed']), ragged=True)
    keep, counts = oks_nms_batched(kpts, scores, areas, off, float(g['thresh']), sig)
    assert list(counts) == list(g['counts'])
    flat = []
    for i in range(len(off) - 1):
        flat.extend(int(v) for v in keep[off[i]:off[i] + counts[i]])
    assert flat == list(g['keep'])
    kb, sb, ab, _ = synth.detections(1, int(g['big_n']), k, seed=int(g['big_seed']))
    db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
    for th in (0.5, 0.9, 0.99):
        assert oks_nms(db, th, sig) == list(g[f'big_keep_{th}'])
    assert oks_nms([], 0.9) == []
    one = oks_nms(db[:1], 0.9, sig)
    assert one == [0]


def test_oks_nms_large_property():
    """100k detections / 5000 images (BASELINE.json config 4): per-image keep sets equal the oracle
    on a sample, and every image keeps its top-scoring detection first."""
    kpts, scores, areas, off = synth.detections(5000, 20, 17, seed=21)
    keep, counts = oks_nms_batched(kpts, scores, areas, off, 0.9)
    assert counts.min() >= 1
    for i in range(0, 5000, 97):
        sl = slice(off[i], off[i + 1])
        ref, _ = nms_oracle.oks_nms_arrays(kpts[sl], scores[sl], areas[sl], 0.9)
        assert list(keep[off[i]:off[i] + counts[i]]) == ref
        assert keep[off[i]] == int(np.argmax(scores[sl]))


def test_rescore_vs_oracle():
    rs = np.random.RandomState(4)
    mv = rs.uniform(0, 1, (500, 14)).astype(np.float32)
    mv[:5] = 0.1                       # nothing visible -> score 0
    box = rs.uniform(0, 1, 500)
    got = rescore(mv, box, 0.2)
    ref = np.array([nms_oracle.rescore(box[i], mv[i], 0.2) for i in range(500)], np.float64)
    assert np.array_equal(got, ref)


def test_oks_iou_vs_oracle():
    kpts, scores, areas, off = synth.detections(1, 40, 17, seed=8)
    flat = kpts.reshape(40, -1)
    got = oks_iou(flat[0], flat[1:], areas[0], areas[1:])
    ref = nms_oracle.oks_iou(flat[0], flat[1:], areas[0], areas[1:])
    assert np.abs(got - ref).max() <= 4e-16          # fp64 exp may differ from NumPy's by an ulp


@pytest.mark.parametrize('tag', ['coco', 'crowdpose'])
def test_soft_oks_nms_vs_reference(golden_dir, tag):
    """soft_oks_nms (nms.py:138-180) on the device: the kept indices, in selection order, equal the reference's on the
    fixtures (per image through the reference's dict interface, and all images in one segmented launch)."""
    g = _load(golden_dir, f'nms_{tag}.npz')
    k = int(g['k'])
    sig = None if tag == 'coco' else nms_oracle.CROWDPOSE_SIGMAS
    kpts, scores, areas, off = synth.detections(int(g['n_imgs']), int(g['per_img']), k, seed=int(g['seed']), ragged=True)
    keep, counts = soft_oks_nms_batched(kpts[:off[40]], scores[:off[40]], areas[:off[40]], off[:41], 0.9, sig)
    assert list(counts) == list(g['soft_counts'])
    pos = 0
    for i in range(40):
        c = int(counts[i])
        assert list(keep[i, :c]) == list(g['soft_keep'][pos:pos + c]), i
        pos += c
    kb, sb, ab, _ = synth.detections(1, int(g['big_n']), k, seed=int(g['big_seed']))
    db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
    got = soft_oks_nms(db, 0.5, sig)
    assert got.dtype == np.intp and list(got) == list(g['soft_big_keep']) and len(got) == 20
    assert list(got) == list(nms_oracle.soft_oks_nms(db, 0.5, sig))
    assert soft_oks_nms([], 0.9) == []


@pytest.mark.parametrize('tag', ['coco', 'crowdpose'])
def test_in_vis_thre_vs_reference(golden_dir, tag):
    """in_vis_thre (nms.py:85-90, the reference's `list(a) and list(b)` = the compared detection's visible key points):
    oks_iou within 1 ulp of the reference (device exp), keep lists of oks_nms / soft_oks_nms identical."""
    g = _load(golden_dir, f'nms_{tag}.npz')
    k = int(g['k'])
    sig = None if tag == 'coco' else nms_oracle.CROWDPOSE_SIGMAS
    kb, sb, ab, _ = synth.detections(1, int(g['big_n']), k, seed=int(g['big_seed']))
    flat = kb.reshape(len(sb), -1)
    got = oks_iou(flat[0], flat[1:], ab[0], ab[1:], sig, 0.4)
    assert np.allclose(got, g['vis_oks'], rtol=4e-16, atol=1e-300)
    assert not np.array_equal(g['vis_oks'], nms_oracle.oks_iou(flat[0], flat[1:], ab[0], ab[1:], sig))   # the mask matters
    kpts, scores, areas, off = synth.detections(int(g['n_imgs']), int(g['per_img']), k, seed=int(g['seed']), ragged=True)
    keep, counts = oks_nms_batched(kpts[:off[40]], scores[:off[40]], areas[:off[40]], off[:41], 0.9, sig, in_vis_thre=0.4)
    assert list(counts) == list(g['vis_counts'])
    pos = 0
    for i in range(40):
        c = int(counts[i])
        assert list(keep[off[i]:off[i] + c]) == list(g['vis_keep'][pos:pos + c]), i
        pos += c
    db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
    assert list(soft_oks_nms(db, 0.5, sig, 0.4)) == list(g['vis_soft_big_keep'])
    assert oks_nms(db, 0.9, sig, in_vis_thre=0.4) == [int(v) for v in nms_oracle.oks_nms(db, 0.9, sig, 0.4)]


@pytest.mark.parametrize('tag', ['crowdpose', 'coco'])
def test_evaluate_device_vs_reference(golden_dir, tag):
    """evaluate_device (grouping + rescoring + (soft-)OKS-NMS in one device call) against the UNMODIFIED reference's
    dataset.evaluate() (crowdpose.py:1255-1324 / coco.py:1210-1277) on interleaved detections: image order, kept
    detections in selection order and the rescored values, all bit-exact."""
    from rsgnet_b200.nms.nms import evaluate_device
    g = _load(golden_dir, f'evaluate_{tag}.npz')
    k = int(g['k'])
    sig = nms_oracle.CROWDPOSE_SIGMAS if tag == 'crowdpose' else None
    preds, boxes, ids = synth.evaluate_inputs(int(g['n_imgs']), int(g['per_img']), k, seed=int(g['seed']))
    for soft, sfx in ((False, ''), (True, '_soft')):
        out = evaluate_device(preds, boxes, ids, float(g['oks_thre']), float(g['in_vis_thre']), sig, soft_nms=soft).host()
        assert np.array_equal(out['images'], g['images' + sfx])
        assert np.array_equal(out['counts'], g['counts' + sfx])
        assert np.array_equal(out['keep'], g['keep' + sfx])
        assert np.array_equal(out['scores'][out['keep']], g['scores' + sfx])
    # every detection's rescored value equals the stand-alone rescoring entry point and the oracle
    every = rescore(preds[:, :, 2], boxes[:, 5], float(g['in_vis_thre']))
    assert np.array_equal(out['scores'], every)
    assert np.array_equal(every[:64], np.array([nms_oracle.rescore(boxes[i, 5], preds[i, :, 2], float(g['in_vis_thre']))
                                                for i in range(64)]))


def test_evaluate_device_large_and_edge_cases():
    """100 k detections / 5000 images (BASELINE.json config 4), shuffled: equals the oracle's evaluate() on the same
    inputs; plus an empty call, a single detection, one image holding everything and NaN scores."""
    from rsgnet_b200.nms.nms import evaluate_device
    preds, boxes, ids = synth.evaluate_inputs(5000, 20, 17, seed=31, ragged=False)
    out = evaluate_device(preds, boxes, ids, 0.9, 0.2).host()
    images, counts, keep, scores = nms_oracle.evaluate(preds, boxes, ids, None, 0.2, 0.9)
    assert np.array_equal(out['images'], images) and np.array_equal(out['counts'], counts)
    assert np.array_equal(out['keep'], keep) and np.array_equal(out['scores'][keep], scores)
    # device-resident inputs, one big image (600 detections: far more than any shared-memory table would hold)
    p1, b1, _ = synth.evaluate_inputs(1, 600, 14, seed=5, ragged=False)
    one = evaluate_device(torch.from_numpy(p1).cuda(), torch.from_numpy(b1).cuda(), torch.zeros(600, dtype=torch.int64).cuda(),
                          0.9, 0.2, nms_oracle.CROWDPOSE_SIGMAS).host()
    im, ct, kp, sc = nms_oracle.evaluate(p1, b1, np.zeros(600, np.int64), nms_oracle.CROWDPOSE_SIGMAS, 0.2, 0.9)
    assert np.array_equal(one['keep'], kp) and np.array_equal(one['counts'], ct)
    empty = evaluate_device(np.zeros((0, 17, 3), np.float32), np.zeros((0, 6)), np.zeros(0, np.int64), 0.9, 0.2).host()
    assert len(empty['images']) == 0 and len(empty['keep']) == 0
    single = evaluate_device(preds[:1], boxes[:1], ids[:1], 0.9, 0.2).host()
    assert list(single['keep']) == [0] and list(single['counts']) == [1]
    # NaN box scores: ranked like NumPy's argsort()[::-1] ranks them (first), every detection gets exactly one rank
    bn = boxes[:40].copy()
    bn[[3, 17], 5] = np.nan
    nan = evaluate_device(preds[:40], bn, np.zeros(40, np.int64), 2.0, 0.2).host()       # thresh 2: nothing is suppressed
    assert sorted(nan['keep']) == list(range(40)) and list(nan['keep'][:2]) == [17, 3]


def test_oks_nms_refuses_oversized_image_and_orders_nan_like_numpy():
    import ctypes as C
    from rsgnet_b200 import _lib
    kpts, scores, areas, off = synth.detections(1, 30, 17, seed=2)
    scores = scores.copy()
    scores[5] = np.nan
    keep, counts = oks_nms_batched(kpts, scores, areas, off, 2.0)
    assert sorted(keep[:30]) == list(range(30)) and keep[0] == 5
    assert list(keep[:30]) == [int(v) for v in scores.argsort()[::-1]]
    # the device-side guard: max_per_img smaller than the image -> keep_counts = -1, nothing written out of bounds
    dev = torch.device('cuda')
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
    k_, s_, a_, o_ = t(kpts, np.float32), t(np.nan_to_num(scores), np.float64), t(areas, np.float64), t(off, np.int32)
    sg = t(nms_oracle.COCO_SIGMAS, np.float64)
    kp = torch.full((30,), -7, dtype=torch.int32, device=dev)
    ct = torch.zeros(1, dtype=torch.int32, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    _lib.check(_lib.lib().rsg_oks_nms(_lib.stream_ptr(dev), p(k_), p(s_), p(a_), p(o_), 1, 8, p(sg), 17, 0.9, p(kp), p(ct), 0, 0.0))
    assert int(ct.item()) == -1 and bool((kp == -7).all())

"""The CPU oracle (oracle/*.py) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  This is what pins the oracle; everything GPU-side is then compared
with the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, model_oracle, nms_oracle
from rsgnet_b200 import presets, synth
from rsgnet_b200.models import _params, pose_hrnet, pose_rsgnet

MODEL_CASES = ['tiny', 'tiny_cp_sub', 'tiny_hrnet', 'w32_coco', 'w32_crowdpose', 'hrnet_w32_coco', 'w48_coco_384']


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, np.int64(-2 ** 31) - a, a)
    b = np.where(b < 0, np.int64(-2 ** 31) - b, b)
    return np.abs(a - b)


@pytest.mark.parametrize('tag', ['small', 'hrnet', 'rsgnet'])
def test_decode_oracle_vs_reference(golden_dir, tag):
    g = _load(golden_dir, f'decode_{tag}.npz')
    k, h, w, n = int(g['k']), int(g['h']), int(g['w']), int(g['n'])
    hm = np.concatenate([synth.heatmaps(n, k, h, w, seed=int(g['seed'])),
                         synth.crafted_heatmaps(k, h, w)])
    c, s = synth.centers_scales(hm.shape[0], seed=int(g['cs_seed']))
    coords, mv = decode_oracle.get_max_preds(hm)
    assert np.array_equal(coords, g['coords_raw'])
    for pp in (0, 1):
        preds, maxvals = decode_oracle.get_final_preds(bool(pp), hm, c, s)
        assert np.array_equal(maxvals, g[f'maxvals_pp{pp}'])
        d = ulp_diff(preds, g[f'preds_pp{pp}'])
        assert d.max() <= 1, d.max()
        assert (d == 0).mean() > 0.95
    # flip + shift + average is exact fp32 arithmetic
    rs = np.random.RandomState(int(g['flip_seed']))
    nb = int(g['flip_n'])
    a = rs.standard_normal((nb, k, h, w)).astype(np.float32)
    b = rs.standard_normal((nb, k, h, w)).astype(np.float32)
    avg = decode_oracle.flip_average(a, b, presets.flip_pairs_for(k), shift=True)
    assert np.array_equal(avg, g['flip_avg'])
    if 'flip_back' in g:
        assert np.array_equal(decode_oracle.flip_back(b, presets.flip_pairs_for(k)),
                              g['flip_back'])


@pytest.mark.parametrize('tag', ['coco', 'crowdpose'])
def test_nms_oracle_vs_reference(golden_dir, tag):
    g = _load(golden_dir, f'nms_{tag}.npz')
    k = int(g['k'])
    sig = None if tag == 'coco' else nms_oracle.CROWDPOSE_SIGMAS
    kpts, scores, areas, off = synth.detections(int(g['n_imgs']), int(g['per_img']), k,
                                                seed=int(g['seed']), ragged=True)
    keeps, counts = [], []
    for i in range(len(off) - 1):
        keep, _ = nms_oracle.oks_nms_arrays(kpts[off[i]:off[i + 1]], scores[off[i]:off[i + 1]],
                                            areas[off[i]:off[i + 1]], float(g['thresh']), sig)
        keeps.extend(keep)
        counts.append(len(keep))
    assert counts == list(g['counts'])
    assert keeps == list(g['keep'])
    kb, sb, ab, _ = synth.detections(1, int(g['big_n']), k, seed=int(g['big_seed']))
    for th in (0.5, 0.9, 0.99):
        keep, _ = nms_oracle.oks_nms_arrays(kb, sb, ab, th, sig)
        assert keep == list(g[f'big_keep_{th}'])
    # dict interface
    db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
    assert [int(v) for v in nms_oracle.oks_nms(db, 0.9, sig)] == list(g['big_keep_0.9'])
    assert nms_oracle.oks_nms([], 0.9) == []
    # soft_oks_nms (nms.py:138-180): first 40 images + the 150-detection image (more than max_dets = 20 candidates)
    pos = 0
    for i in range(40):
        db = [dict(keypoints=kpts[j], score=scores[j], area=areas[j]) for j in range(off[i], off[i + 1])]
        c = int(g['soft_counts'][i])
        assert list(nms_oracle.soft_oks_nms(db, 0.9, sig)) == list(g['soft_keep'][pos:pos + c])
        pos += c
    db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
    assert list(nms_oracle.soft_oks_nms(db, 0.5, sig)) == list(g['soft_big_keep'])
    # in_vis_thre = 0.4 (nms.py:85-90)
    flat = kb.reshape(len(sb), -1)
    assert np.array_equal(nms_oracle.oks_iou(flat[0], flat[1:], ab[0], ab[1:], sig, 0.4), g['vis_oks'])
    pos = 0
    for i in range(40):
        db = [dict(keypoints=kpts[j], score=scores[j], area=areas[j]) for j in range(off[i], off[i + 1])]
        c = int(g['vis_counts'][i])
        assert [int(v) for v in nms_oracle.oks_nms(db, 0.9, sig, 0.4)] == list(g['vis_keep'][pos:pos + c])
        pos += c
    db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
    assert list(nms_oracle.soft_oks_nms(db, 0.5, sig, 0.4)) == list(g['vis_soft_big_keep'])


@pytest.mark.parametrize('tag', ['crowdpose', 'coco'])
def test_evaluate_oracle_vs_reference(golden_dir, tag):
    """oracle/nms_oracle.evaluate (+ rescore) against the UNMODIFIED reference's dataset.evaluate() executed on the
    same inputs (crowdpose.py:1255-1324 / coco.py:1210-1277): image order, per-image kept detections in selection
    order and the rescored values, all bit-exact; greedy and soft NMS."""
    g = _load(golden_dir, f'evaluate_{tag}.npz')
    k = int(g['k'])
    sig = nms_oracle.CROWDPOSE_SIGMAS if tag == 'crowdpose' else None
    preds, boxes, ids = synth.evaluate_inputs(int(g['n_imgs']), int(g['per_img']), k, seed=int(g['seed']))
    # rescoring alone, detection by detection, against the scores the reference attached to what it kept
    direct = np.array([nms_oracle.rescore(boxes[i, 5], preds[i, :, 2], float(g['in_vis_thre'])) for i in g['keep']])
    assert np.array_equal(direct, g['scores'])
    assert (g['scores'] == 0.0).sum() > 0                # the nothing-visible path is exercised
    for soft, sfx in ((False, ''), (True, '_soft')):
        images, counts, keep, scores = nms_oracle.evaluate(preds, boxes, ids, sig, float(g['in_vis_thre']),
                                                           float(g['oks_thre']), soft_nms=soft)
        assert np.array_equal(images, g['images' + sfx])
        assert np.array_equal(counts, g['counts' + sfx])
        assert np.array_equal(keep, g['keep' + sfx])
        assert np.array_equal(scores, g['scores' + sfx])


@pytest.mark.parametrize('key', MODEL_CASES)
def test_model_oracle_vs_reference(golden_dir, key):
    g = _load(golden_dir, f'model_{key}.npz')
    cfg = presets.preset(key)
    mod = pose_rsgnet if cfg.MODEL.NAME == 'pose_rsgnet' else pose_hrnet
    net = mod.get_pose_net(cfg, False)
    sd = _params.synth_state_dict(net, seed=int(g['seed']))
    x = torch.from_numpy(synth.crops(int(g['batch']), cfg.MODEL.IMAGE_SIZE,
                                     seed=int(g['seed']) + 11))
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    stages = {}
    out = model_oracle.forward(sd, cfg, x, stages=stages)
    if cfg.MODEL.NAME == 'pose_hrnet':
        outs = dict(heatmaps=out)
    else:
        outs = dict(zip(('multi_kpt_scores', 'kpt_scores', 'limbs_scores', 'relation_scores'),
                        out))
    sub = int(g['sub'])
    for name, t in outs.items():
        ref_absmax = float(g['absmax.' + name])
        if 'out.' + name in g:
            ref = g['out.' + name]
            got = t.numpy()
        else:
            ref = g['sub.' + name]
            flat = t.reshape(-1).numpy()
            got = flat[::sub] if flat.size > 65536 else flat
        assert got.shape == ref.shape
        err = np.abs(got - ref).max()
        assert err <= 1e-4 * max(ref_absmax, 1e-3), (name, err, ref_absmax)      # fp32 accumulation-order noise (thread count, oneDNN blocking)


# ---------------------------------------------------------------------------------------------
# Input side: crop affine + cv2.warpAffine + ToTensor/Normalize
# ---------------------------------------------------------------------------------------------
def test_warp_oracle_vs_reference(golden_dir):
    """oracle/warp_oracle.py against crops made by the reference's get_affine_transform / crop (cv2.getAffineTransform +
    cv2.warpAffine) and torchvision's ToTensor + Normalize: matrices to 1e-9, every byte and every float bit-exact."""
    from oracle import warp_oracle
    from rsgnet_b200.utils import transforms
    g = _load(golden_dir, 'warp_cases.npz')
    imgs, c, s = synth.images(int(g['n']), seed=int(g['seed']))
    assert np.array_equal(warp_oracle.normalize_lut(), g['lut'])
    assert np.array_equal(transforms.normalize_lut(), g['lut'])
    for i, img in enumerate(imgs):
        rot, size = float(g['rots'][i]), g['sizes'][i]
        for inv, key in ((0, f'trans{i}'), (1, f'trans_inv{i}')):
            t = warp_oracle.get_affine_transform(c[i], s[i], rot, size, inv=inv)
            assert np.abs(t - g[key]).max() <= 1e-9
            # the product's host-side matrix (rsgnet_b200/utils/transforms.py) agrees with both
            assert np.abs(transforms.get_affine_transform(c[i], s[i], rot, size, inv=inv) - g[key]).max() <= 1e-9
        assert np.array_equal(warp_oracle.crop(img, c[i], s[i], size, rot), g[f'crop{i}'])
        assert np.array_equal(warp_oracle.warp_affine_u8(img, g[f'trans{i}'], size), g[f'crop{i}'])
        if f'input{i}' in g.files:
            assert np.array_equal(warp_oracle.crop_input(img, c[i], s[i], size, rot, color_rgb=True), g[f'input{i}'])
    m = transforms.affine_matrices(c, s, g['rots'], (48, 64))
    for i in range(len(imgs)):
        assert np.array_equal(m[i], transforms.get_affine_transform(c[i], s[i], float(g['rots'][i]), (48, 64)))
    assert np.allclose(transforms.affine_transform([3.0, 4.0], m[0]), m[0] @ np.array([3.0, 4.0, 1.0]))


def test_warp_oracle_vs_cv2_random():
    """Where OpenCV is importable (the authoring container; any box with the image's cv2), the restatement is compared
    with cv2.warpAffine itself on fresh random images and general affine matrices."""
    cv2 = pytest.importorskip('cv2')
    from oracle import warp_oracle
    rs = np.random.RandomState(11)
    for trial in range(12):
        sh, sw = int(rs.randint(1, 200)), int(rs.randint(1, 300))
        img = rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
        ang, sc = rs.uniform(0, 6.28), rs.uniform(0.2, 4.0)
        m = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rs.uniform(-40, 80)],
                      [sc * np.sin(ang), sc * np.cos(ang), rs.uniform(-40, 80)]])
        if trial == 0:
            m = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])          # integer-aligned: the (32767, 0, 0, 1) weights
        if trial == 1:
            m = np.array([[2.0, 0.0, 3.0], [0.0, 2.0, -5.0]])
        w, h = int(rs.randint(1, 100)), int(rs.randint(1, 100))
        ref = cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_LINEAR)
        assert np.array_equal(warp_oracle.warp_affine_u8(img, m, (w, h)), ref), trial

"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (this container only).

    python oracle/gen_golden.py            # writes tests/golden/

*** TEST INFRASTRUCTURE ONLY ***  Every fixture is (seeded input description, reference output).
Weights are ``rsgnet_b200.models._params.synth_state_dict`` of this package's parameter
containers loaded into the reference model with ``strict=True`` -- which also proves the
state_dict names/shapes are drop-in compatible.  Inputs come from ``rsgnet_b200.synth`` (NumPy
RandomState: identical bytes on any machine), so fixtures store seeds, not inputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from rsgnet_b200 import presets, synth  # noqa: E402
from rsgnet_b200.models import _params, pose_hrnet, pose_rsgnet  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
SUB = 4099          # prime stride for sub-sampled tensors


def subsample(t):
    flat = t.detach().reshape(-1).numpy()
    return flat[::SUB].copy() if flat.size > 65536 else flat.copy()


def model_case(key, batch, seed, full):
    cfg = presets.preset(key)
    name = cfg.MODEL.NAME
    ours = (pose_rsgnet if name == 'pose_rsgnet' else pose_hrnet).get_pose_net(cfg, False)
    sd = _params.synth_state_dict(ours, seed=seed)
    ref = ref_import.ref_model(cfg, name)
    missing = ref.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.from_numpy(synth.crops(batch, cfg.MODEL.IMAGE_SIZE, seed=seed + 11))
    taps = {}
    hooks = []
    top = dict(ref.named_children())
    for nm in ('layer1', 'stage2', 'stage3', 'stage4', 'vis_conv', 'type_conv',
               'predict_contact_net', 'kpt_net', 'predict_net'):
        if nm in top:
            def hook(_m, _i, o, nm=nm):
                if isinstance(o, (list, tuple)):
                    for b, t in enumerate(o):
                        taps[f'{nm}.{b}'] = t.detach().clone()
                else:
                    taps[nm] = o.detach().clone()
            hooks.append(top[nm].register_forward_hook(hook))
    with torch.no_grad():
        out = ref(x)
    for h in hooks:
        h.remove()
    rec = dict(preset=key, batch=batch, seed=seed, sub=SUB)
    if name == 'pose_hrnet':
        outs = dict(heatmaps=out)
    else:
        outs = dict(multi_kpt_scores=out[0], kpt_scores=out[1], limbs_scores=out[2],
                    relation_scores=out[3])
    for k, v in outs.items():
        if full or k == 'kpt_scores' or k == 'heatmaps':
            rec['out.' + k] = v.numpy().astype(np.float32)
        else:
            rec['sub.' + k] = subsample(v)
        rec['absmax.' + k] = np.float32(v.abs().max())
    for k, v in taps.items():
        rec['tap.' + k] = v.numpy().astype(np.float32) if full else subsample(v)
        rec['tapabsmax.' + k] = np.float32(v.abs().max())
    fn = os.path.join(OUT, f'model_{key}.npz')
    np.savez_compressed(fn, **rec)
    print('wrote', fn, os.path.getsize(fn) // 1024, 'KiB')


def decode_cases():
    f = ref_import.ref_functions()
    for tag, (k, h, w, n) in dict(small=(17, 16, 12, 24), hrnet=(17, 64, 48, 6),
                                  rsgnet=(14, 128, 96, 3)).items():
        hm = np.concatenate([synth.heatmaps(n, k, h, w, seed=2),
                             synth.crafted_heatmaps(k, h, w)])
        n_all = hm.shape[0]
        c, s = synth.centers_scales(n_all, seed=5)
        rec = dict(k=k, h=h, w=w, n=n, seed=2, cs_seed=5)
        for pp in (0, 1):
            cfg = presets.make_cfg(post_process=bool(pp))
            preds, maxvals = f['get_final_preds'](cfg, hm.copy(), c, s)
            rec[f'preds_pp{pp}'] = preds.astype(np.float32)
            rec[f'maxvals_pp{pp}'] = maxvals.astype(np.float32)
            # heat-map-space coordinates: run the reference with an identity-like affine is not
            # possible, so recover them from get_max_preds (+ our own restated offsets, checked
            # in tests against preds through the affine)
        coords, mv = f['get_max_preds'](hm.copy())
        rec['coords_raw'] = coords.astype(np.float32)
        # flip path: reference flip_back on a fresh copy + shift + average (function.py:417-427)
        rs = np.random.RandomState(9)
        nb = 2 if h * w <= 3072 else 1
        a = rs.standard_normal((nb, k, h, w)).astype(np.float32)
        b = rs.standard_normal((nb, k, h, w)).astype(np.float32)
        pairs = presets.flip_pairs_for(k)
        fb = f['flip_back'](b.copy(), pairs)
        fb = np.ascontiguousarray(fb)
        t = torch.from_numpy(fb.copy())
        t[:, :, :, 1:] = t.clone()[:, :, :, 0:-1]
        avg = ((torch.from_numpy(a) + t) * 0.5).numpy()
        rec['flip_seed'] = 9
        rec['flip_avg'] = avg.astype(np.float32)
        rec['flip_n'] = nb
        if h * w <= 192:
            rec['flip_back'] = fb.astype(np.float32)
        fn = os.path.join(OUT, f'decode_{tag}.npz')
        np.savez_compressed(fn, **rec)
        print('wrote', fn, os.path.getsize(fn) // 1024, 'KiB')


def nms_cases():
    f = ref_import.ref_functions()
    from oracle.nms_oracle import CROWDPOSE_SIGMAS
    for tag, (k, sig) in dict(coco=(17, None), crowdpose=(14, CROWDPOSE_SIGMAS)).items():
        kpts, scores, areas, off = synth.detections(300, 20, k, seed=3, ragged=True)
        keeps, counts = [], []
        for i in range(len(off) - 1):
            db = [dict(keypoints=kpts[j], score=scores[j], area=areas[j])
                  for j in range(off[i], off[i + 1])]
            keep = f['oks_nms'](db, 0.9, sig)
            keeps.extend(int(v) for v in keep)
            counts.append(len(keep))
        # a few thresholds on one bigger image
        kb, sb, ab, _ = synth.detections(1, 150, k, seed=4)
        big = {}
        for th in (0.5, 0.9, 0.99):
            db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
            big[f'big_keep_{th}'] = np.asarray(f['oks_nms'](db, th, sig), np.int32)
        # soft_oks_nms (nms.py:138-180) on the first 40 images and on the bigger one (more than 20 detections)
        soft_keep, soft_counts = [], []
        for i in range(40):
            db = [dict(keypoints=kpts[j], score=scores[j], area=areas[j]) for j in range(off[i], off[i + 1])]
            kp = f['soft_oks_nms'](db, 0.9, sig)
            soft_keep.extend(int(v) for v in kp)
            soft_counts.append(len(kp))
        db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
        big['soft_big_keep'] = np.asarray(f['soft_oks_nms'](db, 0.5, sig), np.int32)
        big['soft_keep'] = np.asarray(soft_keep, np.int32)
        big['soft_counts'] = np.asarray(soft_counts, np.int32)
        # in_vis_thre = 0.4 (nms.py:85-90; no caller of the reference passes it): oks_iou values, oks_nms and soft_oks_nms
        flat = kb.reshape(len(sb), -1)
        big['vis_oks'] = np.asarray(f['oks_iou'](flat[0], flat[1:], ab[0], ab[1:], sig, 0.4), np.float64)
        vis_keep, vis_counts = [], []
        for i in range(40):
            db = [dict(keypoints=kpts[j], score=scores[j], area=areas[j]) for j in range(off[i], off[i + 1])]
            kp = f['oks_nms'](db, 0.9, sig, 0.4)
            vis_keep.extend(int(v) for v in kp)
            vis_counts.append(len(kp))
        big['vis_keep'] = np.asarray(vis_keep, np.int32)
        big['vis_counts'] = np.asarray(vis_counts, np.int32)
        db = [dict(keypoints=kb[j], score=sb[j], area=ab[j]) for j in range(len(sb))]
        big['vis_soft_big_keep'] = np.asarray(f['soft_oks_nms'](db, 0.5, sig, 0.4), np.int32)
        fn = os.path.join(OUT, f'nms_{tag}.npz')
        np.savez_compressed(fn, k=k, seed=3, n_imgs=300, per_img=20, thresh=0.9,
                            keep=np.asarray(keeps, np.int32), counts=np.asarray(counts, np.int32),
                            big_seed=4, big_n=150, **big)
        print('wrote', fn, os.path.getsize(fn) // 1024, 'KiB')


def evaluate_cases():
    """dataset.evaluate() of the UNMODIFIED reference (crowdpose.py:1255-1324 with K=14 and its own sigmas;
    coco.py:1210-1277 with K=17 and the default sigmas): per-image grouping in first-appearance order, rescoring,
    OKS-NMS.  Pins rsgnet_b200.nms.evaluate_device and oracle/nms_oracle.{rescore,evaluate}."""
    from oracle.nms_oracle import CROWDPOSE_SIGMAS
    for tag, (k, sig) in dict(crowdpose=(14, CROWDPOSE_SIGMAS), coco=(17, None)).items():
        rec = dict(k=k, n_imgs=400, per_img=12, seed=13, in_vis_thre=0.2, oks_thre=0.9)
        preds, boxes, ids = synth.evaluate_inputs(400, 12, k, seed=13)
        if tag == 'crowdpose':
            paths = [f'data/crowdpose/images/{int(i)}.jpg' for i in ids]
        else:
            paths = [f'data/coco/images/val2017/{int(i):012d}.jpg' for i in ids]
        run = ref_import.ref_dataset_evaluate(tag)
        for soft in (False, True):
            out = run(preds.copy(), boxes.copy(), paths, k, sig, 0.2, 0.9, soft_nms=soft)
            sfx = '_soft' if soft else ''
            rec['images' + sfx] = np.asarray([img[0]['image'] for img in out], np.int64)
            rec['counts' + sfx] = np.asarray([len(img) for img in out], np.int32)
            rec['keep' + sfx] = np.asarray([int(d['center'][0]) for img in out for d in img], np.int32)
            rec['scores' + sfx] = np.asarray([d['score'] for img in out for d in img], np.float64)
        fn = os.path.join(OUT, f'evaluate_{tag}.npz')
        np.savez_compressed(fn, **rec)
        print('wrote', fn, os.path.getsize(fn) // 1024, 'KiB', 'kept', int(rec['counts'].sum()), 'of', len(ids))


def warp_cases():
    """Loader input path (CPJointsDataset.py:1281-1290, cp_test.py:107-115) executed by the reference's own
    get_affine_transform / crop (-> cv2.getAffineTransform, cv2.warpAffine) and torchvision's ToTensor + Normalize."""
    import cv2
    import torchvision.transforms as T
    f = ref_import.ref_functions()
    imgs, centers, scales = synth.images(10, seed=7)
    rots = [0, 0, 0, 0, 0, 0, 30.0, -47.5, 0, 80.0]                 # test-time crops use rot = 0; training rotates
    sizes = [(48, 64)] * 6 + [(192, 256), (48, 64), (288, 384), (48, 64)]
    rec = dict(seed=7, n=10, rots=np.asarray(rots, np.float64), sizes=np.asarray(sizes, np.int32),
               cv2_version=np.bytes_(cv2.__version__))
    tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    for i, img in enumerate(imgs):
        trans = f['get_affine_transform'](centers[i], scales[i], rots[i], sizes[i])
        rec[f'trans{i}'] = np.asarray(trans, np.float64)
        rec[f'trans_inv{i}'] = np.asarray(f['get_affine_transform'](centers[i], scales[i], rots[i], sizes[i], inv=1), np.float64)
        out = f['crop'](img, centers[i], scales[i], sizes[i], rots[i])
        rec[f'crop{i}'] = out
        if i in (0, 6):
            rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
            rec[f'input{i}'] = tf(f['crop'](rgb, centers[i], scales[i], sizes[i], rots[i])).numpy()
    ramp = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, axis=2)       # [256,1,3]
    rec['lut'] = tf(ramp).numpy()[:, :, 0]                                           # [3,256]
    fn = os.path.join(OUT, 'warp_cases.npz')
    np.savez_compressed(fn, **rec)
    print('wrote', fn, os.path.getsize(fn) // 1024, 'KiB')


TSUB = 997          # prime stride of the sub-sampled flat gradient / parameter vectors of the training fixtures


def train_case(key, batch, seed):
    """One iteration of the UNMODIFIED lib/core/function.py:rsgnet_train (or :train for the vanilla HRNet) on the CPU:
    reference model in train mode, reference criterion, torch.optim.Adam(lr=1e-3) as lib/utils/utils.py:70-74 builds it.
    Tensor.cuda() / Module.cuda() are identity here (no GPU in this container) -- the only patch, outside the reference.
    The loop runs twice: in fp32 (what the reference does) and in fp64 (the same code on .double() weights and inputs).
    The fp64 run is the yard-stick: the GroupNorm behind the TRP amplifies fp32 rounding (its input has a mean ~1000x its
    spread with these weights), so two CORRECT fp32 implementations differ by ~1 % in the gradients upstream of it;
    tests bound an implementation's distance to the fp64 run by a multiple of the reference's own fp32 distance."""
    import contextlib
    import io
    import re
    import tempfile
    import types
    f = ref_import.ref_train_functions()
    cfg = presets.preset(key)
    name = cfg.MODEL.NAME
    ours = (pose_rsgnet if name == 'pose_rsgnet' else pose_hrnet).get_pose_net(cfg, False)
    sd = _params.synth_state_dict(ours, seed=seed)
    hw = cfg.MODEL.HEATMAP_SIZE
    b = synth.train_batch(batch, cfg.MODEL.IMAGE_SIZE, hw, cfg.MODEL.NUM_JOINTS, max(int(cfg.MODEL.NUM_LIMBS), 1), seed=seed)
    config = presets._wrap(dict(MODEL=dict(UDP_POSE_ON=False), LOSS=dict(USE_TARGET_WEIGHT=True), PRINT_FREQ=100,
                                DEBUG=dict(DEBUG=False, SAVE_BATCH_IMAGES_GT=False, SAVE_BATCH_IMAGES_PRED=False,
                                           SAVE_HEATMAPS_GT=False, SAVE_HEATMAPS_PRED=False)))

    def run(dtype):
        ref = ref_import.ref_model(cfg, name)
        ref.load_state_dict(sd, strict=True)
        ref = ref.to(dtype)
        tb = {k: torch.from_numpy(v).to(dtype) for k, v in b.items()}
        meta = {}
        if name == 'pose_rsgnet':
            loader = [(tb['input'], tb['target'], tb['target_weight'], tb['all_ins_target'], tb['all_ins_target_weight'],
                       tb['target_limbs'], meta)]
            loop = f['rsgnet_train']
        else:
            loader = [(tb['input'], tb['target'], tb['target_weight'], meta)]
            loop = f['train']
        criterion = f['JointsMSELoss'](use_target_weight=True)
        optimizer = torch.optim.Adam(ref.parameters(), lr=1e-3)
        writer = types.SimpleNamespace(add_scalar=lambda *a, **k: None)
        writer_dict = dict(writer=writer, train_global_steps=0)
        old_t, old_m = torch.Tensor.cuda, torch.nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        out = io.StringIO()
        logging_disable = __import__('logging').disable
        try:
            logging_disable(50)
            with contextlib.redirect_stdout(out), tempfile.TemporaryDirectory() as tmp:
                loop(config, loader, ref, criterion, optimizer, 0, tmp, tmp, writer_dict)
        finally:
            logging_disable(0)
            torch.Tensor.cuda, torch.nn.Module.cuda = old_t, old_m
        losses = None
        if name == 'pose_rsgnet':
            m = re.search(r'multi_kpt_loss: (\S+) kpt_loss: (\S+)\s+limbs_loss: (\S+) relation loss: (\S+)', out.getvalue())
            assert m, out.getvalue()
            losses = np.array([float(v) for v in m.groups()], np.float64)     # multi, target, skeleton, relation
        return ref, losses

    ref, losses = run(torch.float32)
    ref64, losses64 = run(torch.float64)
    rec = dict(preset=key, batch=batch, seed=seed, sub=TSUB)
    if losses is not None:
        rec['losses'], rec['losses64'] = losses, losses64
    new_sd, new_sd64 = ref.state_dict(), ref64.state_dict()
    names = [k for k, p in ref.named_parameters() if p.requires_grad]
    grads, grads64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    rec['names'] = np.array(names)
    rec['grad_norm'] = np.array([float(grads[k].grad.double().norm()) for k in names])
    rec['grad_norm64'] = np.array([float(grads64[k].grad.norm()) for k in names])
    rec['grad_err32'] = np.array([float((grads[k].grad.double() - grads64[k].grad).norm()) for k in names])
    flat_g = torch.cat([grads[k].grad.reshape(-1) for k in names])
    flat_g64 = torch.cat([grads64[k].grad.reshape(-1) for k in names])
    flat_d = torch.cat([(new_sd[k] - sd[k]).reshape(-1) for k in names])
    rec['grad_sub'] = flat_g[::TSUB].numpy().copy()
    rec['grad_sub64'] = flat_g64[::TSUB].numpy().copy()
    rec['delta_sub'] = flat_d[::TSUB].numpy().copy()
    for k in ('final_layer.weight', 'kt_machine.matrix_limb', 'relation_head.g.weight', 'conv1.weight', 'type_features',
              'stage2.0.fuse_layers.0.1.0.weight'):
        if k in grads and grads[k].grad is not None:
            rec['grad.' + k] = grads[k].grad.numpy().copy()
            rec['grad64.' + k] = grads64[k].grad.numpy().copy()
    bufs = [k for k in new_sd if k.endswith('running_mean') or k.endswith('running_var')]
    rec['buf_names'] = np.array(bufs)
    rec['buf_sub'] = torch.cat([new_sd[k].reshape(-1) for k in bufs])[::7].numpy().copy()
    rec['buf_sub64'] = torch.cat([new_sd64[k].reshape(-1) for k in bufs])[::7].numpy().copy()
    rec['num_batches_tracked'] = np.int64(new_sd['bn1.num_batches_tracked'])
    fn = os.path.join(OUT, f'train_{key}.npz')
    np.savez_compressed(fn, **rec)
    rel = rec['grad_err32'] / np.maximum(rec['grad_norm64'], 1e-30)
    print('wrote', fn, os.path.getsize(fn) // 1024, 'KiB', rec.get('losses'), 'fp32-vs-fp64 grad error per tensor: median %.2e max %.2e'
          % (np.median(rel), rel.max()))


def main():
    assert ref_import.available(), 'needs /root/reference'
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ['decode', 'nms', 'warp', 'evaluate', 'model']
    if 'decode' in which:
        decode_cases()
    if 'nms' in which:
        nms_cases()
    if 'warp' in which:
        warp_cases()
    if 'evaluate' in which:
        evaluate_cases()
    if 'model' in which:
        model_case('tiny', 2, 0, full=True)
        model_case('tiny_cp_sub', 2, 1, full=True)
        model_case('tiny_hrnet', 2, 2, full=True)
        model_case('w32_coco', 1, 3, full=False)
        model_case('w32_crowdpose', 1, 4, full=False)
        model_case('hrnet_w32_coco', 1, 5, full=False)
        model_case('w48_coco_384', 1, 6, full=False)
    if 'train' in which:
        train_case('tiny', 2, 0)
        train_case('tiny_cp', 3, 1)
        train_case('tiny_hrnet', 2, 2)
        train_case('w32_coco', 2, 3)          # BASELINE.json configs[4]'s model at full depth and resolution


if __name__ == '__main__':
    main()

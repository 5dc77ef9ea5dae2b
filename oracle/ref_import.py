"""Import the UNMODIFIED reference hot path from /root/reference (this container only).

*** TEST INFRASTRUCTURE ONLY ***  Used by ``oracle/gen_golden.py`` (fixture generation) and by
CPU tests that are skipped when /root/reference is absent (it does not exist on the GPU box).
The three shims are the ones verified in SURVEY.md App. C: an empty pre-registered ``models``
package (its __init__ imports modules that are not in the tree), dummy ``nms.cpu_nms`` /
``nms.gpu_nms`` extension modules (never called by oks_nms), and an attribute-dict config in
place of yacs.  No reference source is copied or modified.
"""
import contextlib
import importlib
import io
import os
import shutil
import sys
import tempfile
import types

import yaml

REF = os.environ.get('RSG_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'lib', 'models'))


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


def load_yaml_cfg(relpath, **model_overrides):
    with open(os.path.join(REF, relpath)) as f:
        cfg = yaml.safe_load(f)
    m = cfg['MODEL']
    m.setdefault('UDP_POSE_ON', False)
    m.setdefault('RELATION_SUB_SAMPLE', False)
    m.setdefault('UP_SCALE', 1)
    if isinstance(m.get('NUM_TYPE_VECTOR'), list):
        m['NUM_TYPE_VECTOR'] = m['NUM_TYPE_VECTOR'][0]
    m.update(model_overrides)
    return to_attr(cfg)


_installed = False


def _install_shims():
    global _installed
    if _installed:
        return
    pkg = types.ModuleType('models')
    pkg.__path__ = [os.path.join(REF, 'lib', 'models')]
    sys.modules['models'] = pkg
    for n in ('nms.cpu_nms', 'nms.gpu_nms'):
        m = types.ModuleType(n)
        setattr(m, n.split('.')[1], None)
        sys.modules[n] = m
    sys.path.insert(0, os.path.join(REF, 'lib'))
    _installed = True


def ref_model(cfg, name):
    """name: 'pose_rsgnet' | 'pose_hrnet'.  Handles the CWD-relative kpt_word_embs.pkl."""
    _install_shims()
    mod = importlib.import_module('models.' + name)
    cwd = os.getcwd()
    tmp = None
    try:
        if name == 'pose_rsgnet':
            k = int(cfg.MODEL.NUM_JOINTS)
            src = os.path.join(REF, 'kpt_word_embs.pkl' if k == 17 else 'cp_kpt_word_embs.pkl')
            tmp = tempfile.mkdtemp()
            shutil.copy(src, os.path.join(tmp, 'kpt_word_embs.pkl'))
            os.chdir(tmp)
        with contextlib.redirect_stdout(io.StringIO()):
            model = mod.get_pose_net(cfg, is_train=False)
    finally:
        os.chdir(cwd)
        if tmp:
            shutil.rmtree(tmp, ignore_errors=True)
    return model.eval()


def ref_functions():
    _install_shims()
    from core.inference import get_final_preds, get_max_preds
    from nms.nms import oks_iou, oks_nms, soft_oks_nms
    from utils.transforms import crop, flip_back, get_affine_transform
    return dict(get_final_preds=get_final_preds, get_max_preds=get_max_preds,
                oks_nms=oks_nms, oks_iou=oks_iou, soft_oks_nms=soft_oks_nms, flip_back=flip_back,
                get_affine_transform=get_affine_transform, crop=crop)


def ref_dataset_evaluate(kind):
    """The reference's own ``evaluate()`` (rescoring + per-image grouping + OKS-NMS), callable without the dataset
    files: ``kind`` = 'crowdpose' -> CrowdPoseSkeletonDataset.evaluate (lib/dataset/crowdpose.py:1255-1324, what
    cp_test.py / rsgnet_validate drive), 'coco' -> COCOCPDataset.evaluate (lib/dataset/coco.py:1210-1277).
    The third-party evaluation packages the module imports at the top (json_tricks, crowdposetools, pycocotools; absent
    here, never reached by this path) are stubbed in sys.modules, ``dataset`` is pre-registered as an empty package (its
    __init__ imports every dataset), and the method is called UNBOUND on a stand-in ``self`` whose
    ``_write_coco_keypoint_results`` captures the per-image kept lists.  Returns f(preds, all_boxes, img_path, k, sigmas,
    in_vis_thre, oks_thre, soft_nms) -> list (per image, first-appearance order) of lists of kept detection dicts."""
    _install_shims()
    for n in ('json_tricks', 'crowdposetools', 'crowdposetools.coco', 'crowdposetools.cocoeval', 'pycocotools',
              'pycocotools.coco', 'pycocotools.cocoeval', 'pycocotools._mask', 'pycocotools.mask'):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.COCO = m.COCOeval = None
            m.__path__ = []
            sys.modules[n] = m
    if 'dataset' not in sys.modules:
        pkg = types.ModuleType('dataset')
        pkg.__path__ = [os.path.join(REF, 'lib', 'dataset')]
        sys.modules['dataset'] = pkg
    if kind == 'crowdpose':
        cls = importlib.import_module('dataset.crowdpose').CrowdPoseSkeletonDataset
    else:
        cls = importlib.import_module('dataset.coco').COCOCPDataset

    def run(preds, all_boxes, img_path, k, sigmas, in_vis_thre, oks_thre, soft_nms=False):
        captured = {}

        def write(kpts, res_file):
            captured['kpts'] = kpts
        fake = types.SimpleNamespace(num_joints=k, in_vis_thre=in_vis_thre, oks_thre=oks_thre, soft_nms=soft_nms,
                                     nms_sigmas=sigmas, image_set='val', _write_coco_keypoint_results=write,
                                     _do_python_keypoint_eval=lambda *a: [('AP', 0.0)])
        cfg = AttrDict(RANK=0)
        out_dir = tempfile.mkdtemp()
        try:
            cls.evaluate(fake, cfg, preds, out_dir, all_boxes, img_path)
        finally:
            shutil.rmtree(out_dir, ignore_errors=True)
        return captured['kpts']
    return run


def ref_train_functions():
    """The reference's own training loops and criterion, importable here: lib/core/function.py (rsgnet_train, train) and
    lib/core/loss.py (JointsMSELoss), UNMODIFIED.  ``yacs`` (lib/config builds its defaults with CfgNode at import) and
    ``tensorboardX`` are absent offline and never reached by one iteration of the loop: both are stubbed in sys.modules."""
    _install_shims()
    sys.modules.setdefault('tensorboardX', types.ModuleType('tensorboardX'))
    if 'yacs' not in sys.modules:
        class CN(dict):
            def __init__(self, init_dict=None, new_allowed=False):
                super().__init__(init_dict or {})
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__
        y, yc = types.ModuleType('yacs'), types.ModuleType('yacs.config')
        yc.CfgNode = CN
        y.config = yc
        sys.modules['yacs'], sys.modules['yacs.config'] = y, yc
    fn = importlib.import_module('core.function')
    loss = importlib.import_module('core.loss')
    return dict(rsgnet_train=fn.rsgnet_train, train=fn.train, JointsMSELoss=loss.JointsMSELoss)

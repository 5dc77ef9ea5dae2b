"""The zero-edit route: with rsgnet_b200/shim in front of the reference's lib/, the reference's own
``core.function`` (the eval loop, lib/core/function.py) imports and binds to the sm_100a drop-ins.  Needs the reference
checkout (this container only; skipped on the GPU box) and runs in a subprocess so that sys.modules stays clean."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('RSG_REFERENCE_ROOT', '/root/reference')

CODE = r'''
import sys, types
sys.path.insert(0, %(root)r)
# third-party packages the reference's utils/vis.py and loops import that are absent offline
for n in ('tensorboardX',):
    sys.modules.setdefault(n, types.ModuleType(n))
if 'yacs' not in sys.modules:                  # lib/config/default.py builds its defaults with yacs' CfgNode at import
    class CN(dict):
        def __init__(self, init_dict=None, new_allowed=False):
            super().__init__(init_dict or {})
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    y, yc = types.ModuleType('yacs'), types.ModuleType('yacs.config')
    yc.CfgNode = CN
    y.config = yc
    sys.modules['yacs'], sys.modules['yacs.config'] = y, yc
import rsgnet_b200.shim
rsgnet_b200.shim.install(lib_dir=%(lib)r)
import core.function as F                      # the reference's loop module, UNMODIFIED
import models, utils.transforms, nms.nms, core.inference
import rsgnet_b200.core.inference as ours_inf, rsgnet_b200.utils.transforms as ours_tr, rsgnet_b200.nms.nms as ours_nms
assert F.__file__.startswith(%(lib)r), F.__file__
assert F.get_final_preds is ours_inf.get_final_preds
assert F.flip_back is ours_tr.flip_back and F.flip_dp_back is ours_tr.flip_dp_back
assert core.inference.get_max_preds is ours_inf.get_max_preds
assert nms.nms.oks_nms is ours_nms.oks_nms and nms.nms.soft_oks_nms is ours_nms.soft_oks_nms
import rsgnet_b200.models.pose_rsgnet as pr, rsgnet_b200.models.pose_hrnet as ph
assert models.pose_rsgnet is pr and models.pose_hrnet is ph
assert eval('models.' + 'pose_rsgnet' + '.get_pose_net') is pr.get_pose_net          # tools/cp_test.py:85
import models.pose_resnet as res                # not shadowed: the reference's own file
assert res.__file__.startswith(%(lib)r)
import utils.utils, core.loss, core.evaluate    # the rest of lib/ keeps working
assert core.evaluate.get_max_preds is ours_inf.get_max_preds
for name in ('transform_preds', 'get_affine_transform', 'affine_transform', 'fliplr_joints', 'crop'):
    assert hasattr(utils.transforms, name), name
print('SHIM-OK')
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'lib', 'core')), reason='needs the reference checkout')
def test_reference_loop_module_imports_through_the_shim():
    code = CODE % dict(root=ROOT, lib=os.path.join(REF, 'lib'))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'SHIM-OK' in r.stdout, r.stdout + r.stderr

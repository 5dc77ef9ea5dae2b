"""CPU-side checks: the C-ABI library is built, loads, exports every symbol include/rsg_b200.h
declares, rejects bad arguments before any launch, and the host-side plan builder produces the
expected op graph.  No compute call is made here (no GPU in this container)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from rsgnet_b200 import _engine, _lib, presets
from rsgnet_b200.models import _params, pose_hrnet, pose_rsgnet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'rsg_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rsg_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_lib.LIB_PATH), 'run `python __graft_entry__.py` to build librsg_b200.so'
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), n
    assert set(_lib.EXPORTS) <= set(names)
    assert _lib.lib().rsg_abi_version() == 1


def test_argument_errors_are_reported_without_a_device():
    L = _lib.lib()
    rc = L.rsg_flip_avg_decode(None, None, None, None, 4, 17, 0, 48, None, None, 1, 1, None, None, None, None)
    assert rc != 0 and b'bad shape' in L.rsg_last_error()
    rc = L.rsg_oks_nms(None, None, None, None, None, 3, 5, None, 99, 0.9, None, None, 0, 0.0)
    assert rc != 0 and b'K=99' in L.rsg_last_error()
    # empty inputs are a no-op, like the reference's `if len(kpts_db) == 0: return []`
    assert L.rsg_oks_nms(None, None, None, None, None, 0, 0, None, 17, 0.9, None, None, 0, 0.0) == 0
    assert L.rsg_flip_avg_decode(None, None, None, None, 0, 17, 64, 48, None, None, 1, 1, None, None, None, None) == 0


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    cfg = presets.preset('tiny_hrnet')
    net = pose_hrnet.get_pose_net(cfg, False).eval()
    with pytest.raises(_lib.RsgError):
        net(torch.zeros(1, 3, 96, 64))
    from rsgnet_b200.core.inference import get_final_preds
    from rsgnet_b200.nms.nms import oks_nms
    with pytest.raises((_lib.RsgError, RuntimeError, AssertionError)):
        get_final_preds(cfg, np.zeros((1, 17, 24, 16), np.float32), np.zeros((1, 2), np.float32),
                        np.ones((1, 2), np.float32))
    assert oks_nms([], 0.9) == []


@pytest.mark.parametrize('key,n_conv', [('w32_crowdpose', 305), ('hrnet_w32_coco', 292)])
def test_plan_builder_graph(key, n_conv):
    cfg = presets.preset(key)
    mod = pose_rsgnet if cfg.MODEL.NAME == 'pose_rsgnet' else pose_hrnet
    net = mod.get_pose_net(cfg, False)
    pb = _engine.PlanBuilder(8)
    info = _engine.build_network(pb, _params.synth_state_dict(net, 0), net.spec)
    kinds = {}
    for k, _, _ in pb.ops:
        kinds[k] = kinds.get(k, 0) + 1
    # a fused BasicBlock op (C0 = 32 branch) stands for two convs, a fused Bottleneck op (layer1) for three
    assert kinds['stem'] == 1 and kinds['fuse'] == 8
    # ... and the fp32 TRP tail (1x1 W conv + GroupNorm) for one conv
    assert kinds['conv'] + 2 * kinds.get('bblock', 0) + 3 * kinds.get('bneck', 0) + kinds.get('trptail', 0) == n_conv
    assert kinds.get('bneck', 0) == 4
    if key == 'w32_crowdpose':
        assert kinds['attention'] == 1 and kinds['trptail'] == 1 and info['S'] == 3072
        # executed FLOPs: the reference graph (18.881 GFLOP/fwd) minus the folded type
        # branch (~1.1 GFLOP), plus nothing else
        assert 17.0e9 < pb.flops_per_fwd < 18.9e9
    # every buffer that is read was written by an earlier op or is a pack-time constant
    for b in pb.bufs:
        assert b.first is not None, b.name


def test_bf16_rounding_matches_torch():
    rs = np.random.RandomState(0)
    a = (rs.standard_normal(4096) * 10 ** rs.uniform(-6, 6, 4096)).astype(np.float32)
    ours = _engine._bf16_bits(a)
    ref = torch.from_numpy(a).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(ours, ref)


def test_engine_cache_invalidation_hooks():
    cfg = presets.preset('tiny_hrnet')
    net = pose_hrnet.get_pose_net(cfg, False)
    net.__dict__['_rsg_engines'] = {'cuda:0': ('v', object())}
    net.load_state_dict(net.state_dict())
    assert net.__dict__['_rsg_engines'] == {}
    net.__dict__['_rsg_engines']['cuda:0'] = ('v', object())
    net.eval()
    assert net.__dict__['_rsg_engines'] == {}

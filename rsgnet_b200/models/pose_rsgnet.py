"""Drop-in for /root/reference/lib/models/pose_rsgnet.py (what tools/cp_test.py loads).

``get_pose_net(cfg, is_train)`` -> nn.Module with the reference's parameter names
(pose_rsgnet.py:603-777) whose eval ``forward(x, relation_target=None)`` returns the reference's
4-tuple (pose_rsgnet.py:955-1021):
    multi_kpt_scores [B,K,2h,2w], kpt_scores [B,K,2h,2w], limbs_scores [B,L,2h,2w] in (0,1),
    relation_scores [B,S,S]  (or [B] when a relation_target is given)
All four are produced by the sm_100a library.  Callers that only use ``kpt_scores`` (the
reference's eval loop, lib/core/function.py:389-398) can set ``model.lazy_aux = True`` to have the
three unused outputs computed only when first touched.
"""
import torch.nn as nn

from ..config import KIND_RSGNET, ModelSpec, cfg_get
from . import _params
from .pose_hrnet import EngineOwner, _init_weights


class RSGNet(EngineOwner):
    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.spec = ModelSpec.from_cfg(cfg, KIND_RSGNET)
        chans = _params.add_backbone(self, self.spec)
        _params.add_rsgnet_heads(self, self.spec, chans[0])
        self.pretrained_layers = cfg_get(cfg, 'MODEL', 'EXTRA', 'PRETRAINED_LAYERS', default=['*'])
        self.lazy_aux = False

    def forward(self, x, relation_target=None):
        from .. import _engine
        return _engine.module_forward(self, x, relation_target=relation_target)

    def init_weights(self, pretrained=''):
        _init_weights(self, pretrained)


def get_pose_net(cfg, is_train, **kwargs):
    model = RSGNet(cfg, **kwargs)
    if is_train and cfg_get(cfg, 'MODEL', 'INIT_WEIGHTS', default=False):
        model.init_weights(cfg_get(cfg, 'MODEL', 'PRETRAINED', default=''))
    return model

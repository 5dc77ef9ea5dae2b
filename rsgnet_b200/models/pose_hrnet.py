"""Drop-in for /root/reference/lib/models/pose_hrnet.py (what tools/test.py loads).

``get_pose_net(cfg, is_train)`` returns an nn.Module whose parameters carry the reference's names
(pose_hrnet.py:277-426) and whose eval forward returns one ``[B, K, H/4, W/4]`` fp32 tensor
(pose_hrnet.py:428-463), computed by the sm_100a library.
"""
import logging
import os

import torch
import torch.nn as nn

from ..config import KIND_HRNET, ModelSpec, cfg_get
from . import _params

logger = logging.getLogger(__name__)


class EngineOwner(nn.Module):
    """nn.Module whose packed-weight engines (rsgnet_b200._engine) are dropped whenever the
    parameters can have changed: load_state_dict, .to()/.cuda()/.half() (``_apply``), train().

    ``torch.nn.DataParallel`` (tools/cp_test.py:99, tools/test.py:97) replicates the module on every forward:
    a replica has an EMPTY ``_parameters`` dict and per-forward broadcast copies of the weights, so it cannot own
    packed weights.  A replica therefore only remembers the module it came from (``_rsg_source``) and the device of
    its parameter copies; its forward takes the per-device engine from the SOURCE module's cache (built once per
    device from the source parameters, never from the broadcast copies)."""

    chunk = 512       # forwards per pass of the plan (one pass per 256-crop flip-test step)

    def _drop_engines(self):
        cache = self.__dict__.get('_rsg_engines')
        if cache:
            cache.clear()
        self.__dict__.pop('_rsg_tensors', None)

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica.__dict__['_rsg_source'] = self.__dict__.get('_rsg_source') or self
        return replica

    def _apply(self, fn, *a, **kw):
        self._drop_engines()
        self.__dict__.pop('_rsg_train_store', None)      # .cuda() / .to() re-allocate the parameters: the flat views are stale
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._drop_engines()
        return super().load_state_dict(*a, **kw)

    def train(self, mode=True):
        self._drop_engines()
        return super().train(mode)


class PoseHighResolutionNet(EngineOwner):
    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.spec = ModelSpec.from_cfg(cfg, KIND_HRNET)
        chans = _params.add_backbone(self, self.spec)
        fk = self.spec.final_conv_kernel
        self.final_layer = nn.Conv2d(chans[0], self.spec.num_joints, fk, 1, 1 if fk == 3 else 0)
        self.pretrained_layers = cfg_get(cfg, 'MODEL', 'EXTRA', 'PRETRAINED_LAYERS', default=['*'])

    def forward(self, x):
        from .. import _engine
        return _engine.module_forward(self, x)

    def init_weights(self, pretrained=''):
        _init_weights(self, pretrained)


def _init_weights(net, pretrained):
    """Same policy as pose_hrnet.py:465-495: N(0, 1e-3) conv/linear weights, zero biases, unit BN,
    then an optional partial load filtered by the first path component."""
    for m in net.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            nn.init.normal_(m.weight, std=0.001)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
    if pretrained and os.path.isfile(pretrained):
        state = torch.load(pretrained, map_location='cpu')
        layers = list(net.pretrained_layers)
        picked = {k: v for k, v in state.items()
                  if layers[0] == '*' or k.split('.')[0] in layers}
        logger.info('=> loading pretrained model %s', pretrained)
        net.load_state_dict(picked, strict=False)
    elif pretrained:
        raise ValueError('{} is not exist!'.format(pretrained))


def get_pose_net(cfg, is_train, **kwargs):
    model = PoseHighResolutionNet(cfg, **kwargs)
    if is_train and cfg_get(cfg, 'MODEL', 'INIT_WEIGHTS', default=False):
        model.init_weights(cfg_get(cfg, 'MODEL', 'PRETRAINED', default=''))
    return model

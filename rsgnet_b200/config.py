"""Model description extracted from a reference-style cfg.

The reference reads its hyper-parameters from a yacs ``CfgNode`` both by attribute and by item
access (/root/reference/lib/models/pose_rsgnet.py:607 vs :621).  Anything that answers to either
style works here: yacs nodes, plain nested dicts (``yaml.safe_load`` output) or attribute dicts.
"""
from dataclasses import dataclass, field
from typing import List

KIND_HRNET = 0
KIND_RSGNET = 1


def cfg_get(cfg, *path, default=None):
    cur = cfg
    for p in path:
        try:
            cur = cur[p]
        except (KeyError, TypeError, IndexError):
            try:
                cur = getattr(cur, p)
            except AttributeError:
                return default
    return cur


@dataclass
class StageSpec:
    num_modules: int
    num_branches: int
    num_blocks: List[int]
    num_channels: List[int]


@dataclass
class ModelSpec:
    kind: int
    num_joints: int
    num_limbs: int
    image_w: int
    image_h: int
    up_scale: int = 1
    relation_sub_sample: bool = False
    type_dim: int = 600
    deconv_kernel: int = 4
    final_conv_kernel: int = 1
    head_channels: int = 32
    stages: List[StageSpec] = field(default_factory=list)

    @property
    def feat_h(self):
        return self.image_h // 4

    @property
    def feat_w(self):
        return self.image_w // 4

    @property
    def heat_h(self):
        return self.feat_h * (self.up_scale if self.kind == KIND_RSGNET else 1)

    @property
    def heat_w(self):
        return self.feat_w * (self.up_scale if self.kind == KIND_RSGNET else 1)

    @classmethod
    def from_cfg(cls, cfg, kind):
        extra = cfg_get(cfg, 'MODEL', 'EXTRA')
        stages = []
        for s in (2, 3, 4):
            sc = cfg_get(extra, f'STAGE{s}')
            block = cfg_get(sc, 'BLOCK', default='BASIC')
            if block != 'BASIC':
                raise ValueError(f'STAGE{s}.BLOCK={block!r}: only BASIC stages are supported '
                                 '(every RSGNet / HRNet pose config uses BASIC)')
            fuse = cfg_get(sc, 'FUSE_METHOD', default='SUM')
            if fuse != 'SUM':
                raise ValueError(f'STAGE{s}.FUSE_METHOD={fuse!r}: only SUM is supported')
            stages.append(StageSpec(int(cfg_get(sc, 'NUM_MODULES')),
                                    int(cfg_get(sc, 'NUM_BRANCHES')),
                                    [int(v) for v in cfg_get(sc, 'NUM_BLOCKS')],
                                    [int(v) for v in cfg_get(sc, 'NUM_CHANNELS')]))
        size = cfg_get(cfg, 'MODEL', 'IMAGE_SIZE')
        tdim = cfg_get(cfg, 'MODEL', 'NUM_TYPE_VECTOR', default=600)
        if isinstance(tdim, (list, tuple)):       # three COCO YAMLs write [600]
            tdim = tdim[0]
        if kind == KIND_RSGNET and cfg_get(cfg, 'MODEL', 'UDP_POSE_ON', default=False):
            raise ValueError('MODEL.UDP_POSE_ON=True is not supported (the reference path itself '
                             'calls an undefined flip_back_offset, lib/core/function.py:406)')
        head = cfg_get(extra, 'OUTPUT_CONVS', default=None)
        head_c = int(head[0]) if head else stages[-1].num_channels[0]
        return cls(kind=kind,
                   num_joints=int(cfg_get(cfg, 'MODEL', 'NUM_JOINTS')),
                   num_limbs=int(cfg_get(cfg, 'MODEL', 'NUM_LIMBS', default=0) or 0),
                   image_w=int(size[0]), image_h=int(size[1]),
                   up_scale=int(cfg_get(cfg, 'MODEL', 'UP_SCALE', default=1) or 1),
                   relation_sub_sample=bool(cfg_get(cfg, 'MODEL', 'RELATION_SUB_SAMPLE',
                                                    default=False)),
                   type_dim=int(tdim),
                   deconv_kernel=int(cfg_get(cfg, 'MODEL', 'FINAL_DECONV_KERNEL_SIZE',
                                             default=4) or 4),
                   final_conv_kernel=int(cfg_get(extra, 'FINAL_CONV_KERNEL', default=1)),
                   head_channels=head_c,
                   stages=stages)

"""cfg objects for the configurations BASELINE.json names, built in code (no YAML needed).

They carry exactly the keys the reference's model constructors and ``get_final_preds`` read
(SURVEY.md §5 'Config / flags'), with both attribute and item access, so the same object can be
handed to the reference (for golden generation) and to this package.
"""


class Cfg(dict):
    """dict with attribute access (stand-in for a yacs CfgNode)."""
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(d):
    return Cfg({k: _wrap(v) for k, v in d.items()}) if isinstance(d, dict) else d


def make_cfg(name='pose_rsgnet', width=32, image_wh=(192, 256), num_joints=17, num_limbs=18,
             up_scale=2, sub_sample=False, channels=None, modules=(1, 4, 3), blocks=4,
             post_process=True, shift_heatmap=True, flip_test=True):
    ch = list(channels) if channels else [width, width * 2, width * 4, width * 8]
    stages = {}
    for i, s in enumerate((2, 3, 4)):
        nb = i + 2
        stages[f'STAGE{s}'] = dict(NUM_MODULES=modules[i], NUM_BRANCHES=nb, BLOCK='BASIC',
                                   NUM_BLOCKS=[blocks] * nb, NUM_CHANNELS=ch[:nb],
                                   FUSE_METHOD='SUM')
    is_rsg = (name == 'pose_rsgnet')
    hm = (image_wh[0] // 4 * (up_scale if is_rsg else 1),
          image_wh[1] // 4 * (up_scale if is_rsg else 1))
    model = dict(NAME=name, INIT_WEIGHTS=False, PRETRAINED='', NUM_JOINTS=num_joints,
                 NUM_LIMBS=num_limbs, IMAGE_SIZE=list(image_wh), HEATMAP_SIZE=list(hm),
                 UP_SCALE=up_scale if is_rsg else 1, RELATION_SUB_SAMPLE=sub_sample,
                 UDP_POSE_ON=False, NUM_TYPE_VECTOR=600, FINAL_DECONV_KERNEL_SIZE=4,
                 TARGET_TYPE='gaussian', SIGMA=2,
                 EXTRA=dict(PRETRAINED_LAYERS=['*'], FINAL_CONV_KERNEL=1, OUTPUT_CONVS=ch,
                            **stages))
    test = dict(FLIP_TEST=flip_test, POST_PROCESS=post_process, SHIFT_HEATMAP=shift_heatmap,
                OKS_THRE=0.9, IN_VIS_THRE=0.2, SOFT_NMS=False, BATCH_SIZE_PER_GPU=32)
    return _wrap(dict(MODEL=model, TEST=test, GPUS=(0,)))


COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
CROWDPOSE_FLIP_PAIRS = [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9], [10, 11]]


def flip_pairs_for(num_joints):
    return COCO_FLIP_PAIRS if num_joints == 17 else CROWDPOSE_FLIP_PAIRS


PRESETS = {
    # BASELINE.json configs[0]: W32 256x192 COCO K=17 (the CPU-reference config)
    'w32_coco': dict(width=32, image_wh=(192, 256), num_joints=17, num_limbs=18),
    # configs[1]: W32 256x192 CrowdPose K=14 -- the config the metric is quoted on
    'w32_crowdpose': dict(width=32, image_wh=(192, 256), num_joints=14, num_limbs=14),
    # configs[2]: W48 384x288 COCO
    'w48_coco_384': dict(width=48, image_wh=(288, 384), num_joints=17, num_limbs=18),
    # vanilla HRNet (tools/test.py)
    'hrnet_w32_coco': dict(name='pose_hrnet', width=32, image_wh=(192, 256), num_joints=17),
    # small shapes for fast tests (same topology: 3 stages, 2/3/4 branches)
    'tiny': dict(channels=[16, 32, 64, 128], image_wh=(64, 96), num_joints=17, num_limbs=18,
                 modules=(1, 2, 2), blocks=2),
    'tiny_cp_sub': dict(channels=[16, 32, 64, 128], image_wh=(64, 96), num_joints=14,
                        num_limbs=14, modules=(1, 1, 2), blocks=2, sub_sample=True),
    # CrowdPose joints without TRP sub-sampling (the reference's training loop builds its relation target at the
    # feature-map size, lib/core/function.py:256-269, so RELATION_SUB_SAMPLE configs cannot be trained by the reference)
    'tiny_cp': dict(channels=[16, 32, 64, 128], image_wh=(64, 96), num_joints=14, num_limbs=14,
                    modules=(1, 1, 2), blocks=2),
    'tiny_hrnet': dict(name='pose_hrnet', channels=[16, 32, 64, 128], image_wh=(64, 96),
                       num_joints=17, modules=(1, 2, 2), blocks=2),
}


def preset(key, **over):
    kw = dict(PRESETS[key])
    kw.update(over)
    return make_cfg(**kw)

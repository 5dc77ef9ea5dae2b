"""Shadows lib/nms/nms.py (whose import of the stale cpu_nms / gpu_nms extension modules fails on Python 3.12,
lib/nms/nms.py:13-14): the pose entry points lib/dataset/{coco,crowdpose,hie}.py import."""
from rsgnet_b200.nms.nms import (evaluate_device, oks_iou, oks_nms, oks_nms_batched, rescore, soft_oks_nms,  # noqa: F401
                                 soft_oks_nms_batched)

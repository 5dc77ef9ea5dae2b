"""rsgnet_b200 -- B200-native (sm_100a) implementation of RSGNet's per-crop inference hot path.

Sub-packages mirror the reference's ``lib/`` layout so its drivers bind unchanged:
``models.pose_rsgnet.get_pose_net``, ``models.pose_hrnet.get_pose_net``,
``core.inference.get_final_preds``, ``nms.nms.oks_nms``, ``utils.transforms.flip_back``.
All compute goes through ``librsg_b200.so`` (C ABI, include/rsg_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"

"""Shadows lib/utils/transforms.py: every name the reference imports from it (lib/core/function.py:18,
lib/core/inference.py:16, lib/dataset/*JointsDataset.py:20-22, lib/dataset/CrowdPoseDataset.py:14-16)."""
from rsgnet_b200.utils.transforms import (affine_matrices, affine_transform, crop, flip_back, flip_dp_back,  # noqa: F401
                                          fliplr_joints, get_affine_transform, transform_preds, warp_crops)

from . import pose_hrnet, pose_rsgnet  # noqa: F401

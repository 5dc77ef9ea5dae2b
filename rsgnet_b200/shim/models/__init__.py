"""Shadows lib/models/__init__.py (which imports modules missing from the reference tree, lib/models/__init__.py:17-19):
``models.pose_rsgnet`` / ``models.pose_hrnet`` are the sm_100a drop-ins, anything else (``models.pose_resnet``) still comes
from the reference's lib/models through the extended ``__path__``."""
import sys
from pkgutil import extend_path

from rsgnet_b200.models import pose_hrnet, pose_rsgnet  # noqa: F401

__path__ = extend_path(__path__, __name__)
sys.modules[__name__ + '.pose_rsgnet'] = pose_rsgnet
sys.modules[__name__ + '.pose_hrnet'] = pose_hrnet

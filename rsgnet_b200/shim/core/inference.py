"""Shadows lib/core/inference.py (``core`` is a namespace package in the reference: no __init__.py, so
``core.function`` / ``core.loss`` / ``core.evaluate`` keep resolving to lib/core)."""
from rsgnet_b200.core.inference import decode_device, get_final_preds, get_max_preds  # noqa: F401

"""Training step of the drop-in models on the sm_100a library (reference lib/core/function.py:240-363)."""
from .step import FusedAdam, TrainStep, module_forward_train  # noqa: F401

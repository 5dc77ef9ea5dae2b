"""Zero-edit route into the reference's drivers.

The reference's tools (``tools/cp_test.py``, ``tools/test.py``) put ``lib/`` on ``sys.path`` (``tools/_init_paths.py:21-24``)
and import ``models``, ``core.function``, ``utils.transforms``, ``nms.nms`` by those names.  This directory holds modules
with the SAME names that re-export the sm_100a implementations; put it in FRONT of ``lib/`` and the reference's loop
(``lib/core/function.py:366-518``) and datasets (``lib/dataset/crowdpose.py:1315``) run unchanged on librsg_b200:

    import rsgnet_b200.shim; rsgnet_b200.shim.install()      # before `import _init_paths` / `import models`

Only the four hot-path modules are shadowed.  Every package here extends its ``__path__`` over the same-named package
of ``lib/`` so that everything else (``core.function``, ``core.loss``, ``utils.utils``, ``utils.vis``,
``models.pose_resnet`` ...) still resolves to the reference's own files.
"""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def install(lib_dir=None):
    """Put the shim directory at the front of sys.path (and `lib_dir`, the reference's lib/, right behind it when
    given).  Idempotent."""
    for p in (lib_dir, SHIM_DIR):
        if p:
            p = os.path.abspath(p)
            if p in sys.path:
                sys.path.remove(p)
            sys.path.insert(0, p)
    return SHIM_DIR

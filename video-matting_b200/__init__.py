"""video_matting_b200: B200-native (sm_100a) replacement for the per-frame data path of
tangih/video-matting - .flo flow fields, bilinear flow warping, forward/backward occlusion
mask, thin-plate-spline deformation and alpha compositing.

The directory name carries a hyphen (``video-matting_b200``); load it with
``__graft_entry__.load_package()`` (registers it as ``video_matting_b200``) or put
``video-matting_b200/dropin`` on ``sys.path`` to shadow the reference's bare module names
(``import flow, tps, augmentation, reader``).
"""
from . import _native, pipeline          # noqa: F401
from . import flow, reader, tps, augmentation, loader, data   # noqa: F401

__version__ = "0.1.0"


def install_dropin():
    """Register the drop-in modules under the reference's bare names."""
    import sys
    for name, mod in (("flow", flow), ("reader", reader), ("tps", tps), ("augmentation", augmentation),
                      ("loader", loader), ("data", data)):
        sys.modules[name] = mod

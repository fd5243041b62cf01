"""Bare-name shim: put this directory on sys.path and `import augmentation` resolves to the B200
implementation (video_matting_b200.augmentation) instead of the reference module."""
import importlib.util as _u
import os as _os
import sys as _sys

if "video_matting_b200" not in _sys.modules:
    _pkg = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    _spec = _u.spec_from_file_location("video_matting_b200", _os.path.join(_pkg, "__init__.py"),
                                       submodule_search_locations=[_pkg])
    _mod = _u.module_from_spec(_spec)
    _sys.modules["video_matting_b200"] = _mod
    try:
        _spec.loader.exec_module(_mod)
    except BaseException:                       # no half-initialised package behind: the next import shows the real error again
        _sys.modules.pop("video_matting_b200", None)
        _sys.modules.pop(__name__, None)
        raise
_sys.modules[__name__] = _sys.modules["video_matting_b200"].augmentation

"""Import the UNMODIFIED reference modules (SURVEY Appendix A.0 shim) from, in this order, $VM_REFERENCE_DIR,
baseline/_ref (what travels to the GPU box, see install_ref.py) or /root/reference (build container).

Test / benchmark infrastructure only: the product never imports this.  The shim restores the numpy aliases the
reference uses (np.float, np.int: removed in numpy 1.24), stubs the absent `progressbar` package and the GUI
calls of headless cv2 (flow.py:52).  Modules are loaded under private names (`_vmref_flow`, ...) with the
reference's bare names visible only while they are being imported, so they never shadow - and are never
shadowed by - the drop-in modules.
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_dir():
    for cand in (os.environ.get("VM_REFERENCE_DIR"), os.path.join(HERE, "_ref"), "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "flow.py")):
            return cand
    return None


def load(names=("reader", "flow", "tps", "augmentation"), overrides=None):
    """dict name -> module of the reference.  ``overrides``: bare names that must resolve to other module
    objects while the reference modules import each other (e.g. {"flow": dropin_flow} to run the reference's
    loader.py on top of the product's flow / reader)."""
    d = reference_dir()
    if d is None:
        raise ImportError("reference modules not found (VM_REFERENCE_DIR, baseline/_ref, /root/reference)")
    import numpy as np
    import cv2
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if "progressbar" not in sys.modules:
        pb = types.ModuleType("progressbar")
        pb.progressbar = lambda it, *a, **k: it
        sys.modules["progressbar"] = pb
    if not hasattr(cv2, "imshow") or getattr(cv2.imshow, "__name__", "") != "_vm_noop":
        def _vm_noop(*a, **k):
            return 27
        try:
            cv2.imshow("", None)
        except Exception:
            cv2.imshow = _vm_noop
            cv2.waitKey = _vm_noop
    overrides = dict(overrides or {})
    order = [n for n in ("params", "reader", "flow", "tps", "augmentation", "data", "loader")
             if n in names or n in ("params", "reader", "flow", "tps")]
    saved = {n: sys.modules.get(n) for n in set(order) | set(overrides)}
    mods = {}
    try:
        for n, m in overrides.items():
            sys.modules[n] = m
        for n in order:
            if n in overrides:
                continue
            path = os.path.join(d, n + ".py")
            if not os.path.exists(path):
                continue
            spec = importlib.util.spec_from_file_location("_vmref_" + n, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[n] = mod            # bare name: what the reference's own `import reader` finds
            spec.loader.exec_module(mod)
            mods[n] = mod
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return {n: mods[n] for n in names if n in mods}

"""Drop-in replacement for the reference ``flow`` module (reference flow.py), running on B200.

Same names, argument meaning and error behaviour as the reference functions.  Inputs may be
NumPy arrays (results come back as NumPy arrays, exactly like the reference), CPU tensors or
CUDA tensors (results stay on the device, no synchronisation except where the reference
semantics need one).  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _native as N
from . import pipeline as P

#: the reference prints ``forward.shape`` from correct_alpha (flow.py:40); kept, but switchable
PRINT_SHAPE = True


def _flow_to_device(flow):
    t, _ = N.to_device(flow)
    if t.dtype != torch.float32:
        # (identity + flow).astype(np.float32) (flow.py:17): a float64 flow is added in float64
        # and rounded once; only float32 flows (what reader.read_flow returns) are supported.
        raise TypeError("flow must be float32 (as returned by reader.read_flow)")
    return t


def warp_img(img, flow):
    """warp img following optical flow (image must be 1 channel) - reference flow.py:9-18."""
    assert len(img.shape) == 2
    src, kind = N.to_device(img)
    return N.from_device(P.flow_warp(src, _flow_to_device(flow)), kind)


def warp_bgr(img, flow):
    """warp img following optical flow - reference flow.py:21-33 (first three channels)."""
    src, kind = N.to_device(img)
    if src.dim() != 3 or src.shape[2] < 3:
        raise IndexError("too many indices for array: warp_bgr needs an (H, W, >=3) image")
    if src.shape[2] != 3:
        src = src[:, :, :3].contiguous()
    return N.from_device(P.flow_warp(src, _flow_to_device(flow)), kind)


def correct_alpha(backward, forward, alpha):
    """Zero alpha where forward/backward flows disagree by more than 15 px - reference
    flow.py:36-65.  Mutates ``alpha`` in place and returns the same object.  The reference's
    ``cv2.imshow`` of the error map is not reproduced (headless)."""
    if PRINT_SHAPE:
        print(tuple(forward.shape))                            # flow.py:40
    b, _ = N.to_device(backward)
    f, _ = N.to_device(forward)
    if b.dtype != torch.float32 or f.dtype != torch.float32:
        b, f = b.to(torch.float32), f.to(torch.float32)
    mask, status = P.occlusion_mask(b, f)
    st = status.cpu()
    if int(st[N.STATUS_NAN_ERR]):
        raise ValueError("cannot convert float NaN to integer")
    if int(st[N.STATUS_INDEX_ERR]):
        raise IndexError("index out of bounds for forward flow (flow.py:46)")
    if isinstance(alpha, torch.Tensor) and alpha.is_cuda:
        if not alpha.is_contiguous():
            raise ValueError("alpha must be contiguous to be corrected in place")
        P.apply_mask(alpha, mask)
        return alpha
    m = mask.cpu().numpy().astype(bool)
    if isinstance(alpha, torch.Tensor):
        alpha[torch.from_numpy(m)] = 0
    else:
        alpha[m] = 0.
    return alpha

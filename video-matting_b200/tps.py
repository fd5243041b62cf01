"""Drop-in replacement for the reference ``tps`` module (reference tps.py: thin-plate-spline
inverse warps after Bookstein, "Principal Warps").  The control-point solve stays on the host
(numpy's truncated pseudo-inverse, exactly the reference's call); the radial-basis evaluation,
the bilinear up-sampling of the transform and the resampling run on the GPU in float64.
"""
import numpy as np
import torch

from . import _native as N
from . import pipeline as P


def _coarse_for(from_points, to_points, plan):
    # reference tps.py:51: the spline is built on the *to* points and maps back to *from*
    coef = P.tps_solve(to_points, from_points)
    ctrl = np.asarray(to_points, dtype=np.float64)
    dev = plan.device
    return P.tps_coarse(torch.from_numpy(ctrl[None]).to(dev), torch.from_numpy(coef[None]).to(dev), plan)[0]


def _make_inverse_warp(from_points, to_points, output_region, approximate_grid):
    """[row coords, col coords] of the inverse transform as CUDA float64 tensors - reference
    tps.py:41-75 (shape (h+1, w+1) when approximate_grid != 1)."""
    N.require_cuda()
    plan = P.get_plan(output_region, approximate_grid)
    t = P.tps_transform(_coarse_for(from_points, to_points, plan), plan)
    return [t[0], t[1]]


def warp_images(from_points, to_points, images, output_region, interpolation_order=1, approximate_grid=2):
    """Warp ``images`` by the thin-plate spline taking from_points to to_points - reference
    tps.py:14-34.  Returns a list with one warped image per input (same dtype, and the
    (x_max-x_min+1, y_max-y_min+1) shape the reference produces)."""
    if interpolation_order not in (0, 1):
        # scipy's spline orders 2..5 need its prefilter; the reference only ever passes 1 (and documents 0)
        raise NotImplementedError("interpolation_order must be 0 (nearest) or 1 (bilinear)")
    N.require_cuda()
    plan = P.get_plan(output_region, approximate_grid)
    coarse = _coarse_for(from_points, to_points, plan)
    out = []
    for image in images:
        src, kind = N.to_device(image if isinstance(image, torch.Tensor) else np.asarray(image))
        if src.dim() != 2:
            raise RuntimeError("invalid shape for coordinate array")      # what scipy raises
        if src.dtype not in (torch.uint8, torch.float64, torch.float32):
            raise TypeError(f"unsupported image dtype {src.dtype}")
        out.append(N.from_device(P.tps_warp(src, coarse, plan, order=interpolation_order), kind))
    return out


def deform(img):
    """Random 5x5 TPS deformation of a 3-channel image - reference tps.py:126-153 (consumes
    the global np.random stream in the reference's order)."""
    from .augmentation import deform_grid
    h, w = img.shape[:2]
    grid, new_grid = deform_grid(h, w, 5)
    res = warp_images(grid, new_grid, [img[:, :, 0], img[:, :, 1], img[:, :, 2]], (0, 0, h, w),
                      interpolation_order=1, approximate_grid=2)
    if isinstance(res[0], torch.Tensor):
        return torch.stack(res, dim=2).contiguous()
    return np.transpose(res, axes=(1, 2, 0)).copy()

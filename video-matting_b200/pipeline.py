"""Device-resident API of the B200 video-matting data path.

Everything here takes and returns CUDA tensors in the canonical layouts and only
enqueues work on the current stream - no host synchronisation, no allocation beyond the
outputs.  The drop-in modules (flow.py, tps.py, augmentation.py, reader.py) are thin
NumPy-compatible wrappers over these functions.

Canonical layouts
    fg     (n, H, W, 4) uint8  BGRA, alpha = A/255          (reference reader.py:16-17)
    flow   (n, H, W, 2) float32 (dx, dy)                    (reference reader.py:29)
    bg     (n_bg, H, W, 3) uint8 BGR
    out    (n, H, W, 4) float32 {B, G, R composite, alpha'} / (n,H,W,3) uint8 + (n,H,W) float32
"""
import ctypes
import functools

import numpy as np
import torch

from . import _native as N

AXIS_DTYPE = np.dtype([("frac", "<f8"), ("i0", "<i4"), ("i1", "<i4")])


# ----------------------------------------------------------------------------------------
# generic single-image ops (any size; uint8 / float32 / float64)
# ----------------------------------------------------------------------------------------

def flow_warp(src, flow):
    """cv2.remap(src, grid + flow, INTER_LINEAR) - reference flow.py:9-33.
    src (sh, sw) or (sh, sw, C) with C in {1,3,4}; flow (h, w, 2) float32."""
    lib = N.load()
    assert flow.dtype == torch.float32 and flow.dim() == 3 and flow.shape[2] == 2
    src = src.contiguous()
    flow = flow.contiguous()
    ch = 1 if src.dim() == 2 else src.shape[2]
    h, w = flow.shape[:2]
    shape = (h, w) if src.dim() == 2 else (h, w, ch)
    dst = torch.empty(shape, dtype=src.dtype, device=src.device)
    N.check(lib.vm_flow_warp(N.ptr(src), N.dtype_code(src), ch, src.shape[0], src.shape[1],
                             N.ptr(flow), h, w, N.ptr(dst), N.stream_ptr()))
    return dst


def occlusion_mask(backward, forward, status=None):
    """uint8 mask of the pixels reference flow.py:41-50 zeroes; status counts the pixels
    where the reference would raise (IndexError / ValueError)."""
    lib = N.load()
    backward = backward.contiguous()
    forward = forward.contiguous()
    h, w = backward.shape[:2]
    if status is None:
        status = N.new_status(backward.device)
    mask = torch.empty((h, w), dtype=torch.uint8, device=backward.device)
    # the reference indexes `forward` with coordinates clamped to backward's size
    if forward.shape[1] != w or forward.shape[0] < h:
        raise IndexError("forward flow must have backward's width and at least its height")
    N.check(lib.vm_occlusion_mask(N.ptr(backward), N.ptr(forward), h, w, N.ptr(mask), N.ptr(status),
                                  N.stream_ptr()))
    return mask, status


def apply_mask(alpha, mask):
    """alpha[mask] = 0 in place (reference flow.py:49-50)."""
    lib = N.load()
    assert alpha.is_contiguous() and mask.is_contiguous() and alpha.numel() == mask.numel()
    N.check(lib.vm_apply_mask(N.ptr(alpha), N.dtype_code(alpha), N.ptr(mask), alpha.numel(), N.stream_ptr()))
    return alpha


def composite(fg, bg, alpha):
    """alpha*fg + (1-alpha)*bg in float64 - reference reader.py:72-79."""
    lib = N.load()
    fg, bg = fg.contiguous(), bg.contiguous()
    alpha = alpha.to(torch.float64).contiguous()
    h, w = fg.shape[:2]
    assert fg.shape == (h, w, 3) and bg.shape == (h, w, 3) and alpha.shape == (h, w)
    out = torch.empty((h, w, 3), dtype=torch.float64, device=fg.device)
    N.check(lib.vm_composite(N.ptr(fg), N.dtype_code(fg), N.ptr(bg), N.dtype_code(bg), N.ptr(alpha), h, w,
                             N.ptr(out), N.stream_ptr()))
    return out


def warp_affine(src, M, dsize):
    """cv2.warpAffine(src, M, (dw, dh)) default flags - reference augmentation.py:59-62."""
    lib = N.load()
    src = src.contiguous()
    dw, dh = int(dsize[0]), int(dsize[1])
    ch = 1 if src.dim() == 2 else src.shape[2]
    Mh = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(6))
    dst = torch.empty((dh, dw) if src.dim() == 2 else (dh, dw, ch), dtype=src.dtype, device=src.device)
    N.check(lib.vm_warp_affine(N.ptr(src), N.dtype_code(src), ch, src.shape[0], src.shape[1],
                               Mh.ctypes.data_as(ctypes.c_void_p), dh, dw, N.ptr(dst), N.stream_ptr()))
    return dst


def illumination(bgr, lut, hsv_vec=None):
    """BGR2HSV -> S,V through a 256-entry table -> HSV2BGR (reference augmentation.py:88-99) of (..., H, W, 3)
    uint8 images, bit-exact for this host's cv2 (``hsv_vec``: pixels per SIMD step of its HSV2BGR, probed by
    default - the results of a row's SIMD body are truncated, those of its tail rounded)."""
    lib = N.load()
    bgr = bgr.contiguous()
    assert bgr.dtype == torch.uint8 and bgr.dim() >= 2 and bgr.shape[-1] == 3
    lut = np.ascontiguousarray(lut, dtype=np.uint8)
    assert lut.shape == (256,)
    out = torch.empty_like(bgr)
    w = bgr.shape[-2]
    rows = bgr.numel() // 3 // max(w, 1)
    N.check(lib.vm_illumination_lut_rows(N.ptr(bgr), rows, w, lut.ctypes.data_as(ctypes.c_void_p),
                                         int(hsv_vec or N.hsv_vec()), N.ptr(out), N.stream_ptr()))
    return out


def resize_u8(img, dsize):
    """cv2.resize(img, dsize=(width, height), interpolation=cv2.INTER_LINEAR) for uint8 images (H, W), (H, W, C) or
    a batch (n, H, W, C), C in {1, 3, 4} - bit-exact for OpenCV 4.13 (reference reader.py:41,53, augmentation.py:160)."""
    lib = N.load()
    img = img.contiguous()
    assert img.dtype == torch.uint8 and img.dim() in (2, 3, 4)
    dw, dh = int(dsize[0]), int(dsize[1])
    if img.dim() == 2:
        n, sh, sw, ch, shape = 1, img.shape[0], img.shape[1], 1, (dh, dw)
    elif img.dim() == 3:
        n, sh, sw, ch, shape = 1, img.shape[0], img.shape[1], img.shape[2], (dh, dw, img.shape[2])
    else:
        n, sh, sw, ch, shape = img.shape[0], img.shape[1], img.shape[2], img.shape[3], (img.shape[0], dh, dw, img.shape[3])
    out = torch.empty(shape, dtype=torch.uint8, device=img.device)
    N.check(lib.vm_resize_u8(N.ptr(img), n, sh, sw, ch, N.ptr(out), dh, dw, N.stream_ptr()))
    return out


def alpha_stats(alpha):
    """device uint64[3] = {count(alpha != 0), sum(rows), sum(cols)} (augmentation.py:10-21)."""
    lib = N.load()
    alpha = alpha.contiguous()
    h, w = alpha.shape
    out = torch.zeros(3, dtype=torch.int64, device=alpha.device)
    N.check(lib.vm_alpha_stats(N.ptr(alpha), N.dtype_code(alpha), h, w, N.ptr(out), N.stream_ptr()))
    return out


def map_coordinates(src, t0, t1, status=None, order=1):
    """scipy.ndimage.map_coordinates(src, [t0, t1], order=order) - reference tps.py:34; order 1 (bilinear) or
    0 (nearest: the sample at floor(t + 1/2), 0 outside [0, n-1])."""
    lib = N.load()
    src = src.contiguous()
    t0 = t0.to(torch.float64).contiguous()
    t1 = t1.to(torch.float64).contiguous()
    ch = 1 if src.dim() == 2 else src.shape[2]
    oh, ow = t0.shape
    dst = torch.empty((oh, ow) if src.dim() == 2 else (oh, ow, ch), dtype=src.dtype, device=src.device)
    N.check(lib.vm_map_coordinates_order(N.ptr(src), N.dtype_code(src), ch, src.shape[0], src.shape[1], N.ptr(t0),
                                         N.ptr(t1), oh, ow, N.ptr(dst), N.ptr(status), int(order), N.stream_ptr()))
    return dst


# ----------------------------------------------------------------------------------------
# thin-plate spline: host solve + device evaluation
# ----------------------------------------------------------------------------------------

from .hostpool import vm_tps_host as _host

_tps_kernel_matrix = _host.tps_kernel_matrix
tps_solve = _host.tps_solve
SolverPool = _host.SolverPool


def axis_table(lo, hi, steps):
    """Up-sampling indices/fractions for one axis (reference tps.py:55-63)."""
    new = np.arange(lo, hi + 1)
    frac, idx = np.modf((steps - 1) * (new - lo) / float(hi - lo))
    i0 = idx.astype(int)
    i1 = (i0 + 1).clip(0, steps - 1).astype(int)
    tab = np.zeros(len(new), dtype=AXIS_DTYPE)
    tab["frac"], tab["i0"], tab["i1"] = frac, i0, i1
    return tab


class TpsPlan:
    """Everything about the TPS evaluation that depends on the output region only."""

    def __init__(self, region, approximate_grid=2, device=None):
        x_min, y_min, x_max, y_max = region
        if approximate_grid is None:
            approximate_grid = 1
        self.region = (x_min, y_min, x_max, y_max)
        self.approximate_grid = approximate_grid
        self.x_steps = (x_max - x_min) / approximate_grid
        self.y_steps = (y_max - y_min) / approximate_grid
        self.nx, self.ny = int(self.x_steps), int(self.y_steps)
        if self.nx < 2 or self.ny < 2:
            raise ValueError("output region too small for the TPS grid")
        # np.mgrid[a:b:n*1j] -> arange(n) * ((b - a) / (n - 1)) + a
        self.step_x = (x_max - x_min) / float(self.nx - 1)
        self.step_y = (y_max - y_min) / float(self.ny - 1)
        self.device = torch.device(device if device is not None else "cuda")
        if approximate_grid != 1:
            self.out_h, self.out_w = x_max - x_min + 1, y_max - y_min + 1
            rows = axis_table(x_min, x_max, self.x_steps)
            cols = axis_table(y_min, y_max, self.y_steps)
            self.rows = torch.from_numpy(rows.view(np.uint8).copy()).to(self.device)
            self.cols = torch.from_numpy(cols.view(np.uint8).copy()).to(self.device)
        else:
            self.out_h, self.out_w = self.nx, self.ny
            self.rows = self.cols = None


@functools.lru_cache(maxsize=32)
def _cached_plan(region, approximate_grid, device_index):
    return TpsPlan(region, approximate_grid, torch.device("cuda", device_index))


def get_plan(region, approximate_grid=2, device=None):
    dev = torch.device(device if device is not None else "cuda")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return _cached_plan(tuple(int(v) for v in region), approximate_grid, idx)


def tps_coarse(ctrl, coef, plan, out=None):
    """Evaluate the spline on the coarse grid of ``plan``.  ctrl (n, N, 2), coef (n, N+3, 2)
    float64 CUDA tensors -> coarse (n, 2, nx, ny) float64."""
    lib = N.load()
    ctrl = ctrl.to(torch.float64).contiguous()
    coef = coef.to(torch.float64).contiguous()
    n, Np = ctrl.shape[0], ctrl.shape[1]
    assert coef.shape == (n, Np + 3, 2)
    if out is None:
        out = torch.empty((n, 2, plan.nx, plan.ny), dtype=torch.float64, device=ctrl.device)
    N.check(lib.vm_tps_coarse(N.ptr(ctrl), N.ptr(coef), n, Np, plan.nx, plan.ny, plan.step_x, plan.step_y,
                              float(plan.region[0]), float(plan.region[1]), N.ptr(out), N.stream_ptr()))
    return out


def tps_transform(coarse, plan):
    """(2, H+1, W+1) float64 inverse transform (reference tps._make_inverse_warp)."""
    lib = N.load()
    assert coarse.shape == (2, plan.nx, plan.ny)
    if plan.approximate_grid == 1:
        return coarse
    h, w = plan.out_h - 1, plan.out_w - 1
    out = torch.empty((2, h + 1, w + 1), dtype=torch.float64, device=coarse.device)
    N.check(lib.vm_tps_upsample(N.ptr(coarse), plan.nx, plan.ny, N.ptr(plan.rows), N.ptr(plan.cols), h, w,
                                N.ptr(out), N.stream_ptr()))
    return out


def tps_warp(src, coarse, plan, out_hw=None, status=None, order=1):
    """Fused up-sample + map_coordinates of one image (reference tps.warp_images, one entry)."""
    lib = N.load()
    src = src.contiguous()
    if plan.approximate_grid == 1:
        return map_coordinates(src, coarse[0], coarse[1], status, order)
    ch = 1 if src.dim() == 2 else src.shape[2]
    oh, ow = out_hw if out_hw is not None else (plan.out_h, plan.out_w)
    assert oh <= plan.out_h and ow <= plan.out_w
    dst = torch.empty((oh, ow) if src.dim() == 2 else (oh, ow, ch), dtype=src.dtype, device=src.device)
    N.check(lib.vm_tps_warp_order(N.ptr(src), N.dtype_code(src), ch, src.shape[0], src.shape[1], N.ptr(coarse),
                                  plan.nx, plan.ny, N.ptr(plan.rows), N.ptr(plan.cols), oh, ow, N.ptr(dst),
                                  N.ptr(status), int(order), N.stream_ptr()))
    return dst


_tps_kernel_matrices = _host.tps_kernel_matrices
_stacked_kernel_ok = _host._stacked_kernel_ok      # same list object: tests flip it
_solve_chunk = _host.solve_chunk


def solve_grids_host(grids, pool=None):
    """solve_grids without the upload: NumPy (ctrl (n, N, 2), coef (n, N + 3, 2)) float64."""
    return pool.solve(grids, per_task=8) if pool is not None else _host.solve_many(grids)


def solve_grids(grids, device=None, pool=None):
    """Host solve for a batch of (regular grid, deformed grid) pairs as used by
    augmentation.warp_image(..., thin=grids): the system is built from the DEFORMED grid and maps
    back onto the regular one (reference tps.py:51).  ``pool``: a SolverPool to spread the solves over
    host cores.  Returns CUDA (ctrl, coef)."""
    ctrl, coef = pool.solve(grids, per_task=8) if pool is not None else _host.solve_many(grids)
    dev = torch.device(device if device is not None else "cuda")
    return N.upload(ctrl, dev), N.upload(coef, dev)


# ----------------------------------------------------------------------------------------
# fused clip-level kernels
# ----------------------------------------------------------------------------------------

def _check_clip(fg, h, w):
    assert fg.dtype == torch.uint8 and fg.dim() == 4 and fg.shape[3] == 4, \
        "fg must be an (n, H, W, 4) uint8 BGRA tensor"


def _flow_arg(flow, n, h, w, name):
    if flow is None:
        return None
    assert flow.dtype == torch.float32 and flow.shape == (n, h, w, 2), f"{name} must be (n, H, W, 2) float32"
    return flow.contiguous()


def flow_warp_mask(fg, backward, forward=None, out_bgr=None, out_alpha=None, status=None):
    """warp_bgr + warp_img + correct_alpha for a clip (reference flow.py:9-65), one pass.
    Returns (bgr uint8 (n,H,W,3), alpha float32 (n,H,W), status)."""
    lib = N.load()
    n, h, w = fg.shape[:3]
    _check_clip(fg, h, w)
    fg = fg.contiguous()
    backward, forward = _flow_arg(backward, n, h, w, "backward"), _flow_arg(forward, n, h, w, "forward")
    if out_bgr is None:
        out_bgr = torch.empty((n, h, w, 3), dtype=torch.uint8, device=fg.device)
    if out_alpha is None:
        out_alpha = torch.empty((n, h, w), dtype=torch.float32, device=fg.device)
    if status is None:
        status = N.new_status(fg.device)
    N.check(lib.vm_flow_warp_mask_bgra(N.ptr(fg), N.ptr(backward), N.ptr(forward), n, h, w, N.ptr(out_bgr),
                                       N.ptr(out_alpha), N.ptr(status), N.stream_ptr()))
    return out_bgr, out_alpha, status


def _tps_args(ctrl, coef, n):
    ctrl = ctrl.to(torch.float64).contiguous()
    coef = coef.to(torch.float64).contiguous()
    Np = ctrl.shape[1]
    assert ctrl.shape == (n, Np, 2) and coef.shape == (n, Np + 3, 2), "ctrl (n,N,2) / coef (n,N+3,2) expected"
    return ctrl, coef, Np


_scratch_cache = {}


def _fused_scratch(lib, n, plan, scratch, device):
    """Workspace of the fused entry points: the packed flow-warped intermediate of the split
    pipeline (a few frames, L2 resident) or the coarse transform of the gather variant.  One
    buffer per (device, stream) is cached, so calls on different streams never share a workspace."""
    h, w = plan.region[2], plan.region[3]
    need = max(int(lib.vm_fused_scratch_bytes(n, h, w)), 256)
    if scratch is None or scratch.numel() * scratch.element_size() < need:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, torch.cuda.current_stream(idx).cuda_stream)      # one workspace per (device, stream): no races
        cached = _scratch_cache.get(key)
        if cached is None or cached.numel() < need:
            cached = torch.empty(need, dtype=torch.uint8, device=device)
            _scratch_cache[key] = cached
        scratch = cached
    return scratch


def tps_composite(fg, bg, ctrl, coef, plan=None, out=None, scratch=None, status=None):
    """TPS warp of fg/alpha + composite onto bg (C3), spline evaluation fused in.
    Returns (out float32 (n,H,W,4), status)."""
    lib = N.load()
    n, h, w = fg.shape[:3]
    _check_clip(fg, h, w)
    assert bg.dtype == torch.uint8 and bg.dim() == 4 and bg.shape[1:] == (h, w, 3)
    fg, bg = fg.contiguous(), bg.contiguous()
    plan = plan or get_plan((0, 0, h, w), 2, fg.device)
    ctrl, coef, Np = _tps_args(ctrl, coef, n)
    if out is None:
        out = torch.empty((n, h, w, 4), dtype=torch.float32, device=fg.device)
    if status is None:
        status = N.new_status(fg.device)
    scratch = _fused_scratch(lib, n, plan, scratch, fg.device)
    N.check(lib.vm_tps_composite_bgra(N.ptr(fg), N.ptr(bg), bg.shape[0], N.ptr(ctrl), N.ptr(coef), Np, plan.nx,
                                      plan.ny, plan.step_x, plan.step_y, N.ptr(plan.rows), N.ptr(plan.cols),
                                      n, h, w, N.ptr(out), N.ptr(scratch), N.ptr(status), N.stream_ptr()))
    return out, status


def flow_tps_composite(fg, backward, forward, bg, ctrl, coef, plan=None, out=None, scratch=None, status=None):
    """flow warp + consistency mask + TPS + composite (SURVEY 8d C4 pipeline), 39 B/px, one
    kernel.  Returns (out float32 (n,H,W,4), status)."""
    lib = N.load()
    n, h, w = fg.shape[:3]
    _check_clip(fg, h, w)
    backward, forward = _flow_arg(backward, n, h, w, "backward"), _flow_arg(forward, n, h, w, "forward")
    assert bg.dtype == torch.uint8 and bg.dim() == 4 and bg.shape[1:] == (h, w, 3)
    fg, bg = fg.contiguous(), bg.contiguous()
    plan = plan or get_plan((0, 0, h, w), 2, fg.device)
    ctrl, coef, Np = _tps_args(ctrl, coef, n)
    if out is None:
        out = torch.empty((n, h, w, 4), dtype=torch.float32, device=fg.device)
    if status is None:
        status = N.new_status(fg.device)
    scratch = _fused_scratch(lib, n, plan, scratch, fg.device)
    N.check(lib.vm_flow_tps_composite_bgra(N.ptr(fg), N.ptr(backward), N.ptr(forward), N.ptr(bg), bg.shape[0],
                                           N.ptr(ctrl), N.ptr(coef), Np, plan.nx, plan.ny, plan.step_x,
                                           plan.step_y, N.ptr(plan.rows), N.ptr(plan.cols), n, h, w, N.ptr(out),
                                           N.ptr(scratch), N.ptr(status), N.stream_ptr()))
    return out, status


_variant = [4]
DEFAULT_VARIANT = 4


def set_fused_variant(v):
    """4 = lean split pipeline (default: flow stage, float64 spline stage, TMA-tiled resampling stage -
    csrc/vm_lean.cu), 5 = single-pass warp-specialised kernel for C4 (csrc/vm_fuse.cu: no intermediate in HBM,
    1.10 x the algorithmic traffic, but slower - DESIGN.md 5f; C3 takes the lean pipeline), 1 = per-pixel gather
    kernels (the generic fallback, any control-point count).  All three give the same bits for 4 and 5 and the
    same values within tolerance for 1 (differential tests)."""
    N.set_option("fused_variant", int(v))
    _variant[0] = int(v)


# ----------------------------------------------------------------------------------------
# clip sharding: independent clips, no collective on the data path
# ----------------------------------------------------------------------------------------

def shard_range(n_units, rank, world):
    """Contiguous slice [lo, hi) of n_units owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ----------------------------------------------------------------------------------------
# host-buffer entry: the call a NumPy user makes for a whole clip (PCIe inside)
# ----------------------------------------------------------------------------------------

class HostClipRunner:
    """flow warp + mask + TPS + composite for clips that live in HOST memory.

    Frames are streamed through two device slots in chunks: H2D copies, the fused kernels and
    the D2H copy of the result run on three streams so that the PCIe transfers of neighbouring
    chunks overlap the kernels.  Inputs should be pinned (torch ``pin_memory()``) for the
    copies to be asynchronous; pageable NumPy arrays work but serialise.
    """

    def __init__(self, h, w, chunk=4, device=None):
        N.require_cuda()
        self.h, self.w, self.chunk = h, w, chunk
        self.device = torch.device(device if device is not None else "cuda")
        d = self.device
        self.plan = get_plan((0, 0, h, w), 2, d)
        self.slots = []
        for _ in range(2):
            self.slots.append(dict(
                fg=torch.empty((chunk, h, w, 4), dtype=torch.uint8, device=d),
                fb=torch.empty((chunk, h, w, 2), dtype=torch.float32, device=d),
                ff=torch.empty((chunk, h, w, 2), dtype=torch.float32, device=d),
                bg=torch.empty((chunk, h, w, 3), dtype=torch.uint8, device=d),
                out=torch.empty((chunk, h, w, 4), dtype=torch.float32, device=d),
                loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event()))
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(d) for _ in range(3))
        self.status = N.new_status(d)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def _as_tensor(x):
        return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))

    def run(self, fg, backward, forward, bg, grids, out, pool=None, kernels=True):
        """fg (n,H,W,4) u8, flows (n,H,W,2) f32, bg (n,H,W,3) u8 host arrays/tensors, ``grids``
        a list of n (grid, deformed grid) pairs, ``out`` a host (n,H,W,4) float32 buffer.
        ``pool``: a SolverPool - the whole clip's TPS systems are then solved by its worker processes while
        the first chunks are on the wire; otherwise each chunk is solved in this thread (stacked pinv,
        bit-identical to the per-frame call).  ``kernels=False`` skips solve and kernels: the copy-only
        ceiling of the same byte volume (bench.py reports it beside the end-to-end rate)."""
        import time
        fg, backward, forward, bg, out = map(self._as_tensor, (fg, backward, forward, bg, out))
        n = fg.shape[0]
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.solve_s = 0.0
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        handle = pool.submit(grids) if (pool is not None and kernels) else None
        ctrl_all = coef_all = None
        for ci, lo in enumerate(range(0, n, self.chunk)):
            hi = min(lo + self.chunk, n)
            m = hi - lo
            slot = self.slots[ci & 1]
            with torch.cuda.stream(self.s_in):
                if ci >= 2:
                    self.s_in.wait_event(slot["computed"])        # inputs of chunk ci-2 consumed
                for key, src in (("fg", fg), ("fb", backward), ("ff", forward), ("bg", bg)):
                    slot[key][:m].copy_(src[lo:hi], non_blocking=True)
                    self.h2d_bytes += src[lo:hi].numel() * src.element_size()
                slot["loaded"].record(self.s_in)
            if kernels:
                # host TPS solve (reference tps.py:113-119) while the frames are on the wire
                t0 = time.perf_counter()
                if handle is not None:
                    if ctrl_all is None:
                        ctrl_all, coef_all = pool.collect(handle)
                    ctrl_h, coef_h = ctrl_all[lo:hi], coef_all[lo:hi]
                else:
                    ctrl_h, coef_h = _host.solve_many(grids[lo:hi])
                self.solve_s += time.perf_counter() - t0
                self.h2d_bytes += ctrl_h.nbytes + coef_h.nbytes
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(slot["loaded"])
                if ci >= 2:
                    self.s_run.wait_event(slot["drained"])        # output of chunk ci-2 copied out
                if kernels:
                    ctrl_d = torch.from_numpy(np.ascontiguousarray(ctrl_h)).to(self.device, non_blocking=True)
                    coef_d = torch.from_numpy(np.ascontiguousarray(coef_h)).to(self.device, non_blocking=True)
                    flow_tps_composite(slot["fg"][:m], slot["fb"][:m], slot["ff"][:m], slot["bg"][:m],
                                       ctrl_d, coef_d, plan=self.plan, out=slot["out"][:m], status=self.status)
                slot["computed"].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot["computed"])
                out[lo:hi].copy_(slot["out"][:m], non_blocking=True)
                self.d2h_bytes += slot["out"][:m].numel() * 4
                slot["drained"].record(self.s_out)
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)
        return out


_runners = {}


def flow_tps_composite_host(fg, backward, forward, bg, grids, out=None, chunk=4, pool=None):
    """NumPy/host-tensor front end of the C4 pipeline (H2D, kernels, D2H inside).  Returns the
    host (n,H,W,4) float32 result; synchronises before returning.  ``pool``: optional SolverPool for the
    host TPS solves."""
    n, h, w = fg.shape[:3]
    key = (h, w, chunk, torch.cuda.current_device())
    if key not in _runners:
        _runners[key] = HostClipRunner(h, w, chunk)
    if out is None:
        out = torch.empty((n, h, w, 4), dtype=torch.float32).pin_memory()
    res = _runners[key].run(fg, backward, forward, bg, grids, out, pool=pool)
    torch.cuda.current_stream().synchronize()
    return res

"""ctypes binding of libvm_sm100a.so (the C ABI declared in include/vm_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or
no CUDA device is present when a kernel is requested, the call raises.
"""
import ctypes
import os

import numpy as np
import torch

from . import _build

VM_U8, VM_F32, VM_F64 = 0, 1, 2
STATUS_WORDS = 8
STATUS_INDEX_ERR, STATUS_NAN_ERR, STATUS_MASKED, STATUS_TPS_OUTSIDE, STATUS_SLOW_TILES, STATUS_BAD_TABLE, STATUS_NEAR_KNIFE = 0, 1, 2, 3, 4, 5, 6

_c = ctypes
_P, _I, _L, _D = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double

# name -> argtypes (restype int unless stated); must list every symbol of include/vm_b200.h
SIGNATURES = {
    "vm_version": [],
    "vm_last_error_string": [],
    "vm_init": [],
    "vm_flow_warp": [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P],
    "vm_occlusion_mask": [_P, _P, _I, _I, _P, _P, _P],
    "vm_apply_mask": [_P, _I, _P, _L, _P],
    "vm_composite": [_P, _I, _P, _I, _P, _I, _I, _P, _P],
    "vm_tps_coarse": [_P, _P, _I, _I, _I, _I, _D, _D, _D, _D, _P, _P],
    "vm_tps_upsample": [_P, _I, _I, _P, _P, _I, _I, _P, _P],
    "vm_tps_warp": [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P],
    "vm_map_coordinates": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P],
    "vm_tps_warp_order": [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P],
    "vm_map_coordinates_order": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P],
    "vm_warp_affine": [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P],
    "vm_change_illumination": [_P, _L, _D, _D, _D, _P, _P],
    "vm_illumination_lut": [_P, _L, _P, _P, _P],
    "vm_illumination_lut_rows": [_P, _L, _I, _P, _I, _P, _P],
    "vm_resize_u8": [_P, _I, _I, _I, _I, _P, _I, _I, _P],
    "vm_alpha_stats": [_P, _I, _I, _I, _P, _P],
    "vm_flow_warp_mask_bgra": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "vm_fused_scratch_bytes": [_I, _I, _I],
    "vm_tps_composite_bgra": [_P, _P, _I, _P, _P, _I, _I, _I, _D, _D, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "vm_flow_tps_composite_bgra": [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _D, _D, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "vm_set_option": [_c.c_char_p, _I],
    "vm_tps_coarse_packed": [_P, _P, _I, _I, _I, _I, _D, _D, _P, _P, _P],
    "vm_aug_tps": [_P, _P, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "vm_alpha_stats_bgra": [_P, _I, _I, _I, _P, _P],
    "vm_aug_affine": [_I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _P],
    "vm_loader_batch": [_P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P],
    "vm_sq_err_sum": [_P, _P, _I, _L, _P, _P],
    "vm_trimap_from_matte": [_P, _I, _I, _I, _I, _P, _P],
    "vm_fg_from_u16": [_P, _L, _P, _P],
    "vm_lean_stage_ms": [_P],
    "vm_lean_launch_count": [],
    "vm_fuse_launch_count": [],
}

_lib = None


class VmError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first if needed) the shared library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # a stale .so is better than nothing only if it exists
            if not os.path.exists(_build.LIB):
                raise VmError(f"libvm_sm100a.so is missing and could not be built: {e}") from e
    if not os.path.exists(_build.LIB):
        raise VmError("libvm_sm100a.so is missing (run __graft_entry__.build())")
    lib = ctypes.CDLL(_build.LIB)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = header / library mismatch: be loud
        fn.argtypes = argtypes
        fn.restype = _c.c_int
    lib.vm_last_error_string.restype = _c.c_char_p
    lib.vm_fused_scratch_bytes.restype = _c.c_int64
    lib.vm_lean_launch_count.restype = _c.c_longlong
    lib.vm_fuse_launch_count.restype = _c.c_longlong
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VmError(f"libvm_sm100a: error {rc}: {load().vm_last_error_string().decode()}")


def require_cuda():
    if not torch.cuda.is_available():
        raise VmError("video_matting_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_DT = {torch.uint8: VM_U8, torch.float32: VM_F32, torch.float64: VM_F64}


def dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype} (uint8, float32, float64 only)") from None


def to_device(x, dtype=None):
    """(CUDA tensor, kind) for a numpy array, CPU tensor or CUDA tensor; kind in
    {'numpy', 'cpu', 'cuda'} tells the caller what to hand back."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        kind = "cuda" if x.is_cuda else "cpu"
        t = x
    else:
        kind = "numpy"
        a = np.ascontiguousarray(x)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    return t.contiguous(), kind


def upload(arr, device):
    """Small host array -> CUDA tensor without blocking the host: a copy from PAGEABLE memory first waits for
    everything already queued on the stream (CUDA's rule for pageable transfers), which serialises the host
    work of a call with the kernels of the previous one.  Staged through torch's caching pinned allocator the
    copy is only enqueued (the allocator keeps the block until the copy has run)."""
    require_cuda()
    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.pin_memory().to(device, non_blocking=True)


_rings = {}


def upload_many(arrays, device, slots=8):
    """Several small host arrays -> CUDA tensors with ONE asynchronous copy from a persistent ring of pinned
    staging buffers (no allocator call, no implicit synchronisation; slot k is reused `slots` calls later, after the
    event recorded behind its copy).  Returns tensors of the arrays' dtypes and shapes (views of one device buffer)."""
    require_cuda()
    device = torch.device(device)
    arrs = [np.ascontiguousarray(a) for a in arrays]
    offs, total = [], 0
    for a in arrs:
        offs.append(total)
        total += (a.nbytes + 15) & ~15
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ring = _rings.get(key)
    if ring is None or ring["bytes"] < total:
        size = max(1 << 16, 1 << (max(total, 1) - 1).bit_length())
        ring = {"bytes": size, "next": 0,
                "slots": [(torch.empty(size, dtype=torch.uint8).pin_memory(), torch.cuda.Event()) for _ in range(slots)], "used": [False] * slots}
        _rings[key] = ring
    k = ring["next"]
    ring["next"] = (k + 1) % len(ring["slots"])
    pinned, event = ring["slots"][k]
    if ring["used"][k]:
        event.synchronize()
    host = pinned.numpy()
    for a, o in zip(arrs, offs):
        host[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
    dev = torch.empty(max(total, 16), dtype=torch.uint8, device=device)
    dev[:total].copy_(pinned[:total], non_blocking=True)
    event.record(torch.cuda.current_stream(device))
    ring["used"][k] = True
    out = []
    for a, o in zip(arrs, offs):
        t = dev[o:o + a.nbytes]
        out.append(t.view(torch.from_numpy(a[:0].reshape(-1)).dtype).view(a.shape) if a.dtype != np.uint8 else t.view(a.shape))
    return out


def from_device(t, kind):
    if kind == "cuda":
        return t
    if kind == "cpu":
        return t.cpu()
    return t.cpu().numpy()


def new_status(device=None):
    return torch.zeros(STATUS_WORDS, dtype=torch.int32, device=device or "cuda")


_hsv_vec = [None]


def hsv_vec():
    """Pixels cv2's HSV2BGR converts per SIMD step on this host (its results are truncated in the SIMD body of a
    row and rounded in the scalar tail, include/vm_b200.h).  Probed once with cv2 itself: the pixel (H, S, V) =
    (0, 1, 1) becomes (0, 0, 1) in the body and (1, 1, 1) in the tail; a row of 255 pixels has 255 % vec tail
    pixels.  32 (AVX2) when cv2 cannot be asked."""
    if _hsv_vec[0] is None:
        try:
            import cv2
            row = np.tile(np.array([0, 1, 1], np.uint8), (1, 255, 1))
            tail = int((cv2.cvtColor(row, cv2.COLOR_HSV2BGR)[0, :, 0] == 1).sum())
            _hsv_vec[0] = tail + 1 if tail + 1 in (16, 32, 64, 128, 256) else 32
        except Exception:
            _hsv_vec[0] = 32
    return _hsv_vec[0]


def set_option(key, value):
    """Tuning / test switches of the library (see vm_set_option in include/vm_b200.h)."""
    check(load().vm_set_option(key.encode(), int(value)))

"""Drop-in replacement for the I/O + compositing half of the reference ``reader`` module
(reference reader.py:10-79).  File decoding stays on the host (cv2 / numpy, as in the
reference); compositing runs on the GPU."""
import ctypes
import os

import numpy as np
import torch

from . import _native as N
from . import pipeline as P

FLO_MAGIC = 202021.25


def read_bgra(img_path):
    """Decoded (H, W, 4) uint8 BGRA image after the uint16 -> uint8 conversion of reference
    reader.py:12-15 (alpha = A / 255, foreground = B,G,R): the device layout of a foreground."""
    import cv2
    img = cv2.imread(img_path, cv2.IMREAD_UNCHANGED)
    if img.dtype == np.uint16:
        # ((img+1)/256 - 1).astype(uint8) with img+1 wrapping in uint16 and -1.0 wrapping to 255
        t = ((img.astype(np.uint32) + 1) & 0xFFFF) / 256. - 1.
        img = (np.trunc(t).astype(np.int64) & 0xFF).astype(np.uint8)
    if img.ndim != 3 or img.shape[2] < 4:
        raise IndexError("index 3 is out of bounds for axis 2 (the foreground must be RGBA)")
    return img


def read_fg_img(img_path):
    """reads a foreground RGBA image -> (alpha float64 (H,W), bgr uint8 (H,W,3)) - reference
    reader.py:10-18, including the uint16 -> uint8 conversion quirk."""
    img = read_bgra(img_path)
    alpha = img[:, :, 3] / 255.
    bgr = img[:, :, :3]
    return alpha, bgr


def read_flow(flow_path):
    """read a Middlebury .flo optical-flow file -> float32 (h, w, 2) - reference reader.py:21-30.
    A bad magic number is reported on stdout and parsing continues, as in the reference."""
    with open(flow_path, 'rb') as f:
        key = np.fromfile(f, dtype=np.float32, count=1)
        if FLO_MAGIC != key:
            print('ERROR: invalid key ({})'.format(key))
        w = np.fromfile(f, dtype=np.int32, count=1)[0]
        h = np.fromfile(f, dtype=np.int32, count=1)[0]
        return np.fromfile(f, dtype=np.float32, count=2 * h * w).reshape((h, w, 2))


def create_composite_image(fg, bg, alpha):
    """alpha*fg + (1-alpha)*bg as float64 (H,W,3) - reference reader.py:72-79."""
    f, kind = N.to_device(fg)
    b, _ = N.to_device(bg)
    a, _ = N.to_device(alpha, torch.float64)
    if f.dtype != b.dtype and torch.float64 in (f.dtype, b.dtype):
        pass                                            # mixed uint8/float64 handled natively
    return N.from_device(P.composite(f, b, a), kind)


def resize_background(bg, h, w):
    """cv2.resize(bg, dsize=(w, h), interpolation=cv2.INTER_LINEAR) (reference reader.py:40-41, 52-53,
    augmentation.py:159-160) on the device - bit-exact for uint8 images (vm_resize_u8); NumPy in, NumPy out, CUDA
    tensors stay on the device."""
    src, kind = N.to_device(bg)
    if src.dtype != torch.uint8:
        raise TypeError("resize_background expects a uint8 image (as cv2.imread returns)")
    return N.from_device(P.resize_u8(src, (w, h)), kind)


def load_test_image(filename='in0062.png', bg_name='sea.jpg'):
    """loads a test image - reference reader.py:33-43 (paths relative to the cwd)."""
    import cv2
    alpha, fg = read_fg_img(os.path.join('test_data', filename))
    bg = cv2.imread(os.path.join('test_data', bg_name))
    h, w = fg.shape[:2]
    if bg.shape[0] != h or bg.shape[1] != w:
        bg = resize_background(bg, h, w)
    return fg, bg, create_composite_image(fg, bg, alpha), alpha


def load_test_video(folder_name='hairball2', bg_name='grass.jpg'):
    """loads a test video - reference reader.py:46-64."""
    import cv2
    print('Loading test video...')
    names = sorted(os.listdir(os.path.join('test_data', folder_name)))
    h, w = cv2.imread(os.path.join('test_data', folder_name, names[0])).shape[:2]
    bg = cv2.imread(os.path.join('test_data', bg_name))
    if bg.shape[0] != h or bg.shape[1] != w:
        bg = resize_background(bg, h, w)
    fg_list, alpha_list, cmp_list = [], [], []
    for name in names:
        alpha, fg = read_fg_img(os.path.join('test_data', folder_name, name))
        fg_list.append(fg)
        alpha_list.append(alpha)
        cmp_list.append(create_composite_image(fg, bg, alpha).astype(np.uint8))
    return fg_list, alpha_list, cmp_list, bg


# ----------------------------------------------------------------------------------------
# clip ingest: files -> device-resident clip in the canonical layouts
# ----------------------------------------------------------------------------------------

def _read_flo_into(path, dst):
    """reader.read_flow (reference reader.py:21-30) straight into ``dst``, a float32 (h, w, 2) view of
    pinned memory: same bad-magic message, ValueError for a short or differently sized file."""
    with open(path, 'rb') as f:
        head = f.read(12)
        if len(head) < 12:
            raise ValueError("cannot reshape array of size 0 into a flow field (truncated header)")
        key = np.frombuffer(head, dtype=np.float32, count=1)
        if FLO_MAGIC != key:
            print('ERROR: invalid key ({})'.format(key))
        w, h = (int(v) for v in np.frombuffer(head, dtype=np.int32, count=2, offset=4))
        if (h, w) != dst.shape[:2]:
            raise ValueError(f"flow {path} is {h}x{w}, the clip is {dst.shape[0]}x{dst.shape[1]}")
        got = f.readinto(memoryview(dst.reshape(-1).view(np.uint8)))
        if got != dst.nbytes:
            raise ValueError(f"cannot reshape array of size {got // 4} into shape ({h},{w},2)")


def load_clip(fg_paths, backward_paths=None, forward_paths=None, bg_paths=None, chunk=8, threads=None):
    """Decode a clip's files and return it device-resident in the layouts of ``pipeline``:
    ``{'fg': (n,H,W,4) uint8 BGRA, 'backward' / 'forward': (n,H,W,2) float32 or None,
    'bg': (n_bg,H,W,3) uint8 or None}``.

    Every file goes through the reference's rules: RGBA PNGs as ``read_fg_img`` reads them
    (reader.py:10-18; 16-bit PNGs are uploaded as decoded and converted by ``vm_fg_from_u16`` on the
    device), ``.flo`` files as ``read_flow`` (reader.py:21-30), backgrounds resized to the frame size
    with ``cv2.resize(INTER_LINEAR)`` when they differ (reader.py:39-41).  Host threads decode
    directly into pinned buffers; each chunk of ``chunk`` frames is copied to the device on a side
    stream as soon as its files are done, while later files are still decoding."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    N.require_cuda()
    lib = N.load()
    n = len(fg_paths)
    for other in (backward_paths, forward_paths):
        if other is not None and len(other) != n:
            raise ValueError("one flow file per frame is required")
    if n == 0:
        raise ValueError("empty clip")
    first = cv2.imread(fg_paths[0], cv2.IMREAD_UNCHANGED)
    if first is None or first.ndim != 3 or first.shape[2] < 4:
        raise IndexError("index 3 is out of bounds for axis 2 (the foreground must be RGBA)")
    h, w = first.shape[:2]
    wide = first.dtype == np.uint16
    dev = torch.device("cuda", torch.cuda.current_device())
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    host = {"fg": pin((n, h, w, 4), torch.uint16 if wide else torch.uint8)}
    out = {"fg": torch.empty((n, h, w, 4), dtype=torch.uint8, device=dev), "backward": None, "forward": None, "bg": None}
    raw16 = torch.empty((min(chunk, n), h, w, 4), dtype=torch.uint16, device=dev) if wide else None
    for key, paths in (("backward", backward_paths), ("forward", forward_paths)):
        if paths is not None:
            host[key] = pin((n, h, w, 2), torch.float32)
            out[key] = torch.empty((n, h, w, 2), dtype=torch.float32, device=dev)
    if bg_paths is not None:
        host["bg"] = pin((len(bg_paths), h, w, 3), torch.uint8)
        out["bg"] = torch.empty((len(bg_paths), h, w, 3), dtype=torch.uint8, device=dev)
    views = {k: v.numpy() if v.dtype != torch.uint16 else v.view(torch.int16).numpy().view(np.uint16) for k, v in host.items()}

    def frame_job(k):
        img = first if k == 0 else cv2.imread(fg_paths[k], cv2.IMREAD_UNCHANGED)
        if img is None or img.shape != (h, w, 4) or img.dtype != views["fg"].dtype:
            raise ValueError(f"{fg_paths[k]}: every frame of a clip must be {h}x{w} RGBA of one bit depth")
        np.copyto(views["fg"][k], img)
        if backward_paths is not None:
            _read_flo_into(backward_paths[k], views["backward"][k])
        if forward_paths is not None:
            _read_flo_into(forward_paths[k], views["forward"][k])

    def bg_job(k):
        img = cv2.imread(bg_paths[k])
        if img is None:
            raise AttributeError("'NoneType' object has no attribute 'shape'")
        if img.shape[0] != h or img.shape[1] != w:
            img = cv2.resize(img, dsize=(w, h), interpolation=cv2.INTER_LINEAR)      # reader.py:39-41
        np.copyto(views["bg"][k], img)

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    workers = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=workers) as pool:
        bg_f = [pool.submit(bg_job, k) for k in range(len(bg_paths))] if bg_paths is not None else []
        frame_f = [pool.submit(frame_job, k) for k in range(n)]
        for lo in range(0, n, chunk):
            hi = min(lo + chunk, n)
            for f in frame_f[lo:hi]:
                f.result()
            with torch.cuda.stream(side):
                if wide:
                    raw16[:hi - lo].copy_(host["fg"][lo:hi], non_blocking=True)
                    N.check(lib.vm_fg_from_u16(N.ptr(raw16), (hi - lo) * h * w * 4, N.ptr(out["fg"][lo:hi]),
                                               ctypes.c_void_p(side.cuda_stream)))
                else:
                    out["fg"][lo:hi].copy_(host["fg"][lo:hi], non_blocking=True)
                for key in ("backward", "forward"):
                    if out[key] is not None:
                        out[key][lo:hi].copy_(host[key][lo:hi], non_blocking=True)
        for f in bg_f:
            f.result()
        if bg_paths is not None:
            with torch.cuda.stream(side):
                out["bg"].copy_(host["bg"], non_blocking=True)
    side.synchronize()               # every copy has landed: the staging buffers can go back to torch's pools
    return out

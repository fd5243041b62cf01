"""Drop-in replacement for the reference ``loader`` module (reference loader.py), the production
caller of ``flow.warp_img`` and ``reader.create_composite_image``.

Same names, argument meaning, return shapes/dtypes and ``np.random`` draw order as the reference.
The split of work is different: the host only decodes files (in a thread pool) and *plans* each
sample - the crop type, the padding offsets of ``get_padded_img`` and the crop origins are drawn
from the global ``np.random`` stream in the reference's order and turned into two view
descriptors; one launch of ``vm_loader_batch`` (csrc/vm_loader.cu) then does the flow warp, padding,
crop, float64 ``cv2.resize``, composite, mean subtraction and mirror for the whole batch.  The
decoded uint8 files are the only bytes that cross PCIe on the way in.  There is no CPU fallback.

Extra keyword arguments (not in the reference): ``device=True`` returns CUDA tensors instead of
NumPy arrays (no device-to-host copy, no synchronisation); ``dtype`` selects float32 outputs.
"""
import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _native as N
from . import reader

VGG_MEAN = [103.939, 116.779, 123.68]                  # reference params.py:10
CROP_TYPES = [(320, 320), (480, 480), (640, 640)]      # reference loader.py:47, 124, 295
DECODE_THREADS = min(32, os.cpu_count() or 1)

VIEW_DTYPE = np.dtype([("win_h", "<i4"), ("win_w", "<i4"), ("wi", "<i4"), ("wj", "<i4"),
                       ("vi0", "<i4"), ("vi1", "<i4"), ("vj0", "<i4"), ("vj1", "<i4"),
                       ("si", "<i4"), ("sj", "<i4"), ("mode", "<i4"), ("reserved", "<i4"),
                       ("scale_y", "<f8"), ("scale_x", "<f8")], align=True)
SAMPLE_DTYPE = np.dtype([("fg", "<u8"), ("prev", "<u8"), ("flow", "<u8"), ("bg", "<u8"),
                         ("fh", "<i4"), ("fw", "<i4"), ("bh", "<i4"), ("bw", "<i4"),
                         ("oy", "<i4"), ("ox", "<i4"), ("ph", "<i4"), ("pw", "<i4"), ("prev_stride", "<i4"),
                         ("flip", "<i4"), ("fgv", VIEW_DTYPE), ("bgv", VIEW_DTYPE)], align=True)
assert VIEW_DTYPE.itemsize == 64 and SAMPLE_DTYPE.itemsize == 200     # vm_loader_view / vm_loader_sample

_ALIGN = 256


# ----------------------------------------------------------------------------------------
# host-side planning: the reference's random decisions as view descriptors
# ----------------------------------------------------------------------------------------

class _Canvas:
    """An image seen through the zero canvas of get_padded_img (loader.py:10-36): canvas size and
    the rectangle [vi0,vi1) x [vj0,vj1) that holds image pixels starting at image (si, sj)."""

    def __init__(self, h, w):
        self.h, self.w = h, w
        self.vi0, self.vi1, self.vj0, self.vj1, self.si, self.sj = 0, h, 0, w, 0, 0

    def pad(self, crop_h, crop_w):
        """get_padded_img on an un-padded image: rows are drawn before columns; an axis shorter than
        the crop is placed at a random offset, a longer one is cut to a crop-sized window at the canvas
        origin (the rest of that axis stays zero)."""
        h, w = self.h, self.w
        if crop_h > h:
            o = int(np.random.randint(0, crop_h - h + 1))
            self.vi0, self.vi1, self.si = o, o + h, 0
        else:
            self.vi0, self.vi1, self.si = 0, crop_h, int(np.random.randint(0, h - crop_h + 1))
        if crop_w > w:
            o = int(np.random.randint(0, crop_w - w + 1))
            self.vj0, self.vj1, self.sj = o, o + w, 0
        else:
            self.vj0, self.vj1, self.sj = 0, crop_w, int(np.random.randint(0, w - crop_w + 1))
        self.h, self.w = max(crop_h, h), max(crop_w, w)
        return self

    def window(self, i, j, nh, nw, out_h, out_w):
        """View record of canvas[i:i+nh, j:j+nw] resized to (out_h, out_w)."""
        win_h, win_w = max(0, min(nh, self.h - i)), max(0, min(nw, self.w - j))
        if win_h == 0 or win_w == 0:
            raise ValueError("empty crop window (cv2.resize would fail on an empty image)")
        v = np.zeros((), dtype=VIEW_DTYPE)
        v["win_h"], v["win_w"], v["wi"], v["wj"] = win_h, win_w, i, j
        v["vi0"], v["vi1"], v["vj0"], v["vj1"] = self.vi0, self.vi1, self.vj0, self.vj1
        v["si"], v["sj"] = self.si, self.sj
        v["mode"] = 1 if (win_w == 2 * out_w and win_h == 2 * out_h) else 0     # cv2: INTER_LINEAR -> INTER_AREA
        v["scale_y"], v["scale_x"] = 1.0 / (out_h / win_h), 1.0 / (out_w / win_w)
        return v


def _plan_sample(fh, fw, bh, bw, input_size):
    """The np.random.randint sequence of load_and_crop / simple_load_crop / video_load_crop after the
    files are read (loader.py:47-69, 124-145, 295-315) -> (fg view, bg view)."""
    out_w, out_h = int(input_size[0]), int(input_size[1])                      # cv2 dsize = (width, height)
    crop_h, crop_w = CROP_TYPES[np.random.randint(0, len(CROP_TYPES))]
    fgc = _Canvas(fh, fw)
    if fh < crop_h or fw < crop_w:
        fgc.pad(crop_h, crop_w)
    i = int(np.random.randint(0, fgc.h - crop_h + 1))
    j = int(np.random.randint(0, fgc.w - crop_w + 1))
    fgv = fgc.window(i, j, crop_h, crop_h, out_h, out_w)                       # sic: crop_h for both axes
    bch = int(np.ceil(crop_h * bh / int(fgv["win_h"])))
    bcw = int(np.ceil(crop_w * bw / int(fgv["win_w"])))
    bgc = _Canvas(bh, bw).pad(bch, bcw)
    i = int(np.random.randint(0, bh - bch + 1))
    j = int(np.random.randint(0, bw - bcw + 1))
    return fgv, bgc.window(i, j, bch, bcw, out_h, out_w)


# ----------------------------------------------------------------------------------------
# decode (host, parallel) and staging (one pinned buffer, one H2D copy per batch)
# ----------------------------------------------------------------------------------------

def _imread_bgr(path):
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise AttributeError("'NoneType' object has no attribute 'astype'")    # what loader.py:42 raises
    return img


def _decode(kind, entry):
    """Files of one list entry -> dict of contiguous uint8 / float32 arrays."""
    if kind == "video":
        fg_path, bg_path, prev_path, flo_path = entry
        d = {"fg": reader.read_bgra(fg_path), "bg": _imread_bgr(bg_path), "flow": reader.read_flow(flo_path),
             "prev": reader.read_bgra(prev_path)}
        if d["prev"].shape[:2] != d["flow"].shape[:2] or d["fg"].shape[:2] != d["flow"].shape[:2]:
            raise ValueError("foreground, previous frame and flow must have the same size")
        return d
    fg_path, _tr_path, bg_path = entry
    # the trimap of load_and_crop is padded/cropped/resized by the reference but never returned
    # (loader.py:80-83), so it is not decoded here
    return {"fg": reader.read_bgra(fg_path), "bg": _imread_bgr(bg_path)}


def _decode_all(kind, entries):
    if len(entries) <= 1 or DECODE_THREADS <= 1:
        return [_decode(kind, e) for e in entries]
    with ThreadPoolExecutor(max_workers=DECODE_THREADS) as pool:
        return list(pool.map(lambda e: _decode(kind, e), entries))


def _touched(view, rows, cols):
    """Image rectangle (r0, r1, c0, c1) a view can read, and the view re-based onto it."""
    v = view.copy()
    r0 = int(v["si"]) + max(0, int(v["wi"]) - int(v["vi0"]))
    c0 = int(v["sj"]) + max(0, int(v["wj"]) - int(v["vj0"]))
    nr = min(int(v["vi1"]), int(v["wi"]) + int(v["win_h"])) - max(int(v["vi0"]), int(v["wi"]))
    nc = min(int(v["vj1"]), int(v["wj"]) + int(v["win_w"])) - max(int(v["vj0"]), int(v["wj"]))
    if nr <= 0 or nc <= 0:                                   # the window sees padding only
        return (0, 1, 0, 1), v
    # image pixel = (canvas - v0) + s  ->  rectangle pixel = image pixel - (r0, c0)
    v["si"], v["sj"] = int(v["si"]) - r0, int(v["sj"]) - c0
    return (r0, min(rows, r0 + nr), c0, min(cols, c0 + nc)), v


def _stage(decoded, records):
    """Copy what the kernel can touch - the foreground / flow rectangle under the crop window, the
    previous frame's alpha plane, the background rectangle - and the descriptor table into one
    pinned buffer, send it to the device with one copy, and patch the device addresses into the
    descriptors.  Returns (device buffer, pinned buffer, device pointer of the descriptor table)."""
    jobs, total = [], 0

    def reserve(arr_view):
        nonlocal total
        off = total
        total += (arr_view.size * arr_view.itemsize + _ALIGN - 1) // _ALIGN * _ALIGN
        return off

    for d, r in zip(decoded, records):
        (r0, r1, c0, c1), r["fgv"] = _touched(r["fgv"], *d["fg"].shape[:2])
        r["fh"], r["fw"], r["oy"], r["ox"] = r1 - r0, c1 - c0, r0, c0
        jobs.append((r, "fg", d["fg"][r0:r1, c0:c1], None))
        if "prev" in d:
            jobs.append((r, "flow", d["flow"][r0:r1, c0:c1], None))
            jobs.append((r, "prev", d["prev"][:, :, 3], None))
            r["ph"], r["pw"], r["prev_stride"] = d["prev"].shape[0], d["prev"].shape[1], 1
        (r0, r1, c0, c1), r["bgv"] = _touched(r["bgv"], *d["bg"].shape[:2])
        r["bh"], r["bw"] = r1 - r0, c1 - c0
        jobs.append((r, "bg", d["bg"][r0:r1, c0:c1], None))
    jobs = [(r, k, a, reserve(a)) for r, k, a, _ in jobs]
    table_off = total
    total += (records.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(total, dtype=torch.uint8, device="cuda")
    hv = host.numpy()
    base = dev.data_ptr()

    def put(job):
        r, k, a, off = job
        dst = hv[off:off + a.size * a.itemsize].view(a.dtype).reshape(a.shape)
        np.copyto(dst, a)                                    # one strided copy straight into pinned memory

    if len(jobs) > 4 and DECODE_THREADS > 1:
        with ThreadPoolExecutor(max_workers=DECODE_THREADS) as pool:
            list(pool.map(put, jobs))
    else:
        for job in jobs:
            put(job)
    for r, k, a, off in jobs:
        r[k] = base + off
    hv[table_off:table_off + records.nbytes] = records.view(np.uint8).reshape(-1)
    dev.copy_(host, non_blocking=True)
    return dev, host, base + table_off


def _run(kind, entries, input_size, mirror=False, device=False, dtype=np.float64):
    """Shared driver: decode -> plan (reference RNG order) -> stage -> one kernel launch."""
    N.require_cuda()
    lib = N.load()
    out_w, out_h = int(input_size[0]), int(input_size[1])
    decoded = _decode_all(kind, entries)
    records = np.zeros(len(entries), dtype=SAMPLE_DTYPE)
    for d, r in zip(decoded, records):
        r["fgv"], r["bgv"] = _plan_sample(d["fg"].shape[0], d["fg"].shape[1], d["bg"].shape[0], d["bg"].shape[1],
                                          input_size)
        if mirror:
            r["flip"] = 1 if np.random.uniform(0., 1.) > 0.5 else 0            # loader.py:106-110
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    n = len(entries)
    mk = lambda c: torch.empty((n, out_h, out_w, c), dtype=tdt, device="cuda")
    out = {"cmp": mk(3), "bg": mk(3), "label": mk(1), "fg": mk(3), "warped": mk(3) if kind == "video" else None}
    if n:
        dev, host, table = _stage(decoded, records)
        mean = (ctypes.c_double * 3)(*VGG_MEAN)
        N.check(lib.vm_loader_batch(ctypes.c_void_p(table), n, out_h, out_w, mean, N.dtype_code(out["cmp"]),
                                    N.ptr(out["cmp"]), N.ptr(out["bg"]), N.ptr(out["label"]), N.ptr(out["warped"]),
                                    N.ptr(out["fg"]), N.stream_ptr()))
        del host, dev            # both were used on the current stream only: torch's allocators order their reuse
    if device:
        return out
    host_out = {k: (torch.empty(v.shape, dtype=v.dtype, pin_memory=True).copy_(v, non_blocking=True)
                    if v is not None else None) for k, v in out.items()}
    torch.cuda.current_stream().synchronize()
    return {k: (v.numpy() if v is not None else None) for k, v in host_out.items()}


def device_batch(fg, bg, input_size, prev=None, flow=None, mirror=False, dtype=np.float64):
    """Training batch from frames that are already on the device (no files, no PCIe): the same
    planning (reference np.random order, loader.py:47-69 / 295-315) and the same kernel as
    ``video_batch`` / ``simple_batch``.

    fg: list of (H,W,4) uint8 BGRA CUDA tensors (or one (n,H,W,4) tensor); bg: list of (h,w,3) uint8 CUDA
    tensors; prev + flow: previous BGRA frames and (H,W,2) float32 flows for the video variant.
    Returns a dict of CUDA tensors {cmp, bg, label, fg, warped}; nothing is synchronised."""
    N.require_cuda()
    lib = N.load()
    out_w, out_h = int(input_size[0]), int(input_size[1])
    n = len(fg)
    video = prev is not None
    if video and (flow is None or len(prev) != n or len(flow) != n):
        raise ValueError("prev and flow must hold one entry per sample")
    if len(bg) != n:
        raise ValueError("one background per sample is required")
    keep = []                                                   # contiguous views must outlive the launch
    records = np.zeros(n, dtype=SAMPLE_DTYPE)
    for k, r in enumerate(records):
        f, b = fg[k].contiguous(), bg[k].contiguous()
        if f.dtype != torch.uint8 or f.dim() != 3 or f.shape[2] != 4 or b.dtype != torch.uint8 or b.shape[2] != 3:
            raise TypeError("fg must be (H,W,4) uint8 BGRA and bg (h,w,3) uint8")
        keep += [f, b]
        r["fg"], r["bg"] = f.data_ptr(), b.data_ptr()
        r["fh"], r["fw"], r["bh"], r["bw"] = f.shape[0], f.shape[1], b.shape[0], b.shape[1]
        if video:
            p, fl = prev[k].contiguous(), flow[k].contiguous()
            if p.shape != f.shape or fl.shape != (f.shape[0], f.shape[1], 2) or fl.dtype != torch.float32:
                raise ValueError("foreground, previous frame and flow must have the same size")
            keep += [p, fl]
            r["prev"], r["flow"] = p.data_ptr() + 3, fl.data_ptr()
            r["ph"], r["pw"], r["prev_stride"] = p.shape[0], p.shape[1], 4
        r["fgv"], r["bgv"] = _plan_sample(f.shape[0], f.shape[1], b.shape[0], b.shape[1], input_size)
        if mirror:
            r["flip"] = 1 if np.random.uniform(0., 1.) > 0.5 else 0
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    mk = lambda c: torch.empty((n, out_h, out_w, c), dtype=tdt, device="cuda")
    out = {"cmp": mk(3), "bg": mk(3), "label": mk(1), "fg": mk(3), "warped": mk(3) if video else None}
    if n:
        table = torch.from_numpy(records.view(np.uint8).reshape(-1).copy()).pin_memory().cuda(non_blocking=True)
        mean = (ctypes.c_double * 3)(*VGG_MEAN)
        N.check(lib.vm_loader_batch(N.ptr(table), n, out_h, out_w, mean, N.dtype_code(out["cmp"]),
                                    N.ptr(out["cmp"]), N.ptr(out["bg"]), N.ptr(out["label"]), N.ptr(out["warped"]),
                                    N.ptr(out["fg"]), N.stream_ptr()))
    return out


def _square(input_size):
    # the reference's batch arrays are (B, input_size[0], input_size[1], C) while cv2.resize makes
    # (input_size[1], input_size[0]) samples: a non-square size fails in the assignment (loader.py:344)
    if int(input_size[0]) != int(input_size[1]):
        raise ValueError("could not broadcast input array: batch loaders need a square input_size")


# ----------------------------------------------------------------------------------------
# reference API
# ----------------------------------------------------------------------------------------

def get_padded_img(img, crop_h, crop_w):
    """returns padded image, the original image being randomly placed in the output window -
    reference loader.py:10-36 (host array in, host array out; same two np.random draws)."""
    img = np.asarray(img)
    c = _Canvas(img.shape[0], img.shape[1]).pad(crop_h, crop_w)
    out = np.zeros((c.h, c.w, img.shape[2]), dtype=img.dtype)
    out[c.vi0:c.vi1, c.vj0:c.vj1] = img[c.si:c.si + c.vi1 - c.vi0, c.sj:c.sj + c.vj1 - c.vj0]
    return out


def load_and_crop(entry, input_size, device=False, dtype=np.float64):
    """loads input/label from training list entry (fg, trimap, bg) -> (inp (h,w,6), label (h,w,1),
    fg (h,w,3)) - reference loader.py:39-85."""
    o = _run("trimap", [entry], input_size, device=device, dtype=dtype)
    cat = torch.cat if device else np.concatenate
    return cat((o["cmp"][0], o["bg"][0]), 2), o["label"][0], o["fg"][0]


def random_scale(input, label, raw_fg):
    """reference loader.py:88-90 (a TODO stub there as well)."""
    return [], [], []


def get_batch(file_list, input_size, rd_scale=False, rd_mirror=False, device=False, dtype=np.float64):
    """returns normalized batch of cropped images - reference loader.py:93-116."""
    if rd_scale:
        raise ValueError("could not broadcast input array: random_scale is an empty stub in the reference")
    _square(input_size)
    o = _run("trimap", file_list, input_size, mirror=rd_mirror, device=device, dtype=dtype)
    cat = torch.cat if device else np.concatenate
    return cat((o["cmp"], o["bg"]), 3), o["label"], o["fg"]


def simple_load_crop(entry, input_size, device=False, dtype=np.float64):
    """(fg, trimap, bg) entry -> (cmp, bg, label, fg) - reference loader.py:119-157."""
    o = _run("simple", [entry], input_size, device=device, dtype=dtype)
    return o["cmp"][0], o["bg"][0], o["label"][0], o["fg"][0]


def simple_batch(file_list, input_size, device=False, dtype=np.float64):
    """reference loader.py:160-171 -> (cmps, bgs, label, raw_fgs)."""
    _square(input_size)
    o = _run("simple", file_list, input_size, device=device, dtype=dtype)
    return o["cmp"], o["bg"], o["label"], o["fg"]


def video_load_crop(entry, input_size, device=False, dtype=np.float64):
    """(fg, bg, previous fg, flow) entry -> (cmp, bg, label, warped_alpha, fg) - reference
    loader.py:285-330."""
    o = _run("video", [entry], input_size, device=device, dtype=dtype)
    return o["cmp"][0], o["bg"][0], o["label"][0], o["warped"][0], o["fg"][0]


def video_batch(file_list, input_size, device=False, dtype=np.float64):
    """reference loader.py:333-345 -> (cmps, bgs, label, warped, raw_fgs)."""
    _square(input_size)
    o = _run("video", file_list, input_size, device=device, dtype=dtype)
    return o["cmp"], o["bg"], o["label"], o["warped"], o["fg"]


def get_file_list(root_dir, list_path):
    """reads file list - reference loader.py:192-200 (every line: 'fg trimap bg', newline-terminated)."""
    with open(list_path, 'r') as f:
        return [[os.path.join(root_dir, rel) for rel in line[:-1].split(' ')] for line in f]


def get_batch_list(file_list, batch_size):
    """returns file list for current batch (pops from the end) - reference loader.py:203-208."""
    return [file_list.pop() for _ in range(batch_size)]


def epoch_is_over(file_list, batch_size):
    """reference loader.py:211-213."""
    return len(file_list) < batch_size


def psnr(img, img_ref):
    """peak signal to noise ratio of [0, 1] float images - reference loader.py:214-227."""
    lib = N.load()
    a, _ = N.to_device(img)
    b, _ = N.to_device(img_ref)
    if a.dtype not in (torch.float32, torch.float64) or a.dtype != b.dtype:
        a, b = a.to(torch.float64), b.to(torch.float64)
    if a.shape != b.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(a.shape)} {tuple(b.shape)}")
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    N.check(lib.vm_sq_err_sum(N.ptr(a), N.ptr(b), N.dtype_code(a), a.numel(), N.ptr(acc), N.stream_ptr()))
    eqm = float(acc.item()) / (a.shape[0] * a.shape[1])
    return 10. * np.log10(1. / (1e-6 + eqm))


def add_noise(img, var=0.1):
    """reference loader.py:230-237: gaussian noise from the global np.random stream, clipped to [0,1]
    (a host helper of the reference's demo; the draw has to come from the host stream)."""
    noised = np.array(img, dtype=np.float64, copy=True)
    if noised.ndim == 2:
        noised = noised.reshape(noised.shape + (1,))
    return np.clip(noised + np.random.normal(0., var, noised.shape), 0, 1)


def video_file_list(params=None):
    """(train, test) lists of (fg, bg, previous fg, flow) paths under ./flow and ./SYNTHETIC -
    reference loader.py:240-282.  The sequence-name lists come from the caller's ``params`` module
    (TRAIN_AUGMENTED, TEST_AUGMENTED, TRAIN_SYNTHETIC, TEST_SYNTHETIC), imported by its bare name
    as the reference does when none is passed."""
    if params is None:
        import params                                                    # the user's params.py
    lists = {k: [] for k in ("atr", "ate", "str", "ste")}
    flow_dir = os.path.join('flow', 'augmented', 'flow')
    for filename in os.listdir(flow_dir):
        stem = filename.split('.')[0].split('_')
        basename, id_ = '_'.join(stem[:-1]), int(stem[-1])
        fg_dir, bg_dir = os.path.join('flow', 'augmented', 'fg'), os.path.join('flow', 'augmented', 'bg')
        item = (os.path.join(fg_dir, '{}_fg_{:04d}.png'.format(basename, id_)),
                os.path.join(bg_dir, '{}_bg_{:04d}.png'.format(basename, id_)),
                os.path.join(fg_dir, '{}_fg_ref.png'.format(basename)),
                os.path.join(flow_dir, filename))
        if not all(os.path.isfile(p) for p in item[:3]):
            print('ERROR LOADING FILE {} FOR ID {}'.format(basename, id_))
            print(item[0]); print(item[1]); print(item[2])
            continue
        if basename in params.TRAIN_AUGMENTED:
            lists["atr"].append(item)
        elif basename in params.TEST_AUGMENTED:
            lists["ate"].append(item)
        else:
            print('ERROR, CANT FIND {}'.format(basename))
    syn = os.path.join('flow', 'synthetic')
    for vid in os.listdir(syn):
        for filename in sorted(os.listdir(os.path.join(syn, vid))):
            id_ = int(filename.split('.')[0][2:])
            item = (os.path.join('SYNTHETIC', 'fg', vid, 'in{:04d}.png'.format(id_ + 1)),
                    os.path.join('SYNTHETIC', 'bg', vid, 'in{:04d}.png'.format(id_ + 1)),
                    os.path.join('SYNTHETIC', 'fg', vid, 'in{:04d}.png'.format(id_)),
                    os.path.join(syn, vid, filename))
            if not os.path.isfile(item[2]):
                continue
            if not os.path.isfile(item[1]) or not os.path.isfile(item[0]):
                print('ERROR LOADING INPUT (ID: {} / VIDEO: {})'.format(id_, vid))
            if vid in params.TRAIN_SYNTHETIC:
                lists["str"].append(item)
            elif vid in params.TEST_SYNTHETIC:
                lists["ste"].append(item)
    return lists["atr"] + lists["str"], lists["ate"] + lists["ste"]

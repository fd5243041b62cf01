"""Drop-in replacement for the I/O + compositing half of the reference ``reader`` module
(reference reader.py:10-79).  File decoding stays on the host (cv2 / numpy, as in the
reference); compositing runs on the GPU."""
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


def load_test_image(filename='in0062.png', bg_name='sea.jpg'):
    """loads a test image - reference reader.py:33-43 (paths relative to the cwd)."""
    import cv2
    alpha, fg = read_fg_img(os.path.join('test_data', filename))
    bg = cv2.imread(os.path.join('test_data', bg_name))
    h, w = fg.shape[:2]
    if bg.shape[0] != h or bg.shape[1] != w:
        bg = cv2.resize(bg, dsize=(w, h), interpolation=cv2.INTER_LINEAR)
    return fg, bg, create_composite_image(fg, bg, alpha), alpha


def load_test_video(folder_name='hairball2', bg_name='grass.jpg'):
    """loads a test video - reference reader.py:46-64."""
    import cv2
    print('Loading test video...')
    names = sorted(os.listdir(os.path.join('test_data', folder_name)))
    h, w = cv2.imread(os.path.join('test_data', folder_name, names[0])).shape[:2]
    bg = cv2.imread(os.path.join('test_data', bg_name))
    if bg.shape[0] != h or bg.shape[1] != w:
        bg = cv2.resize(bg, dsize=(w, h), interpolation=cv2.INTER_LINEAR)
    fg_list, alpha_list, cmp_list = [], [], []
    for name in names:
        alpha, fg = read_fg_img(os.path.join('test_data', folder_name, name))
        fg_list.append(fg)
        alpha_list.append(alpha)
        cmp_list.append(create_composite_image(fg, bg, alpha).astype(np.uint8))
    return fg_list, alpha_list, cmp_list, bg

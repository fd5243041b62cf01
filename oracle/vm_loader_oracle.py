"""CPU oracle for the batch loader.  TEST INFRASTRUCTURE ONLY.

Same rules as ``vm_oracle.py``: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU
legs of the benchmark scripts may import this module; nothing under ``video-matting_b200/``
does.  It restates, in NumPy, what the reference's ``loader.py`` computes for one training
sample *after* the files are decoded, including the arithmetic of the one third-party kernel
involved, ``cv2.resize(INTER_LINEAR)`` on float64 images (opencv-python-headless 4.13.0; the
reference converts every image to float64 before resizing, loader.py:41-42, 121-122,
287-288, so the 11-bit fixed-point uint8 path of ``cv2.resize`` is never taken here).

Pinning: ``tests/golden/loader_golden.npz`` holds outputs of the *unmodified* reference
``loader.video_load_crop`` / ``simple_load_crop`` / ``load_and_crop`` on small seeded files
(``tests/golden/make_loader_golden.py``); ``tests/test_loader_cpu.py`` checks this module
against them and, where cv2 is importable, ``resize_linear_f64`` against ``cv2.resize``.
"""
import numpy as np

import vm_oracle as O

VGG_MEAN = [103.939, 116.779, 123.68]                 # reference params.py:10
CROP_TYPES = [(320, 320), (480, 480), (640, 640)]     # reference loader.py:47, 124, 295


# --------------------------------------------------------------------------------------
# cv2.resize(src float64, dsize, INTER_LINEAR)   (loader.py:70-73, 146-148, 316-319)
# --------------------------------------------------------------------------------------

def _linear_axis(n_src, n_dst):
    """Source index and fraction of every destination index (OpenCV resize.cpp: pixel
    centres aligned, ``scale = 1 / (n_dst / n_src)``, double-precision coefficients for
    64-bit images [probed: a ramp image is reproduced exactly])."""
    scale = 1.0 / (n_dst / n_src)
    f = (np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5
    s = np.floor(f)
    return s.astype(np.int64), f - s


def resize_linear_f64(src, dsize):
    """``cv2.resize(src.astype(float64), dsize=(width, height), interpolation=INTER_LINEAR)``.

    * same size: a copy;
    * exactly 2x smaller along both axes: OpenCV switches to INTER_AREA, the mean of each
      2x2 block (640 -> 320 crops take this branch);
    * otherwise separable linear interpolation, horizontal pass first; columns left of the
      first / right of the last sample centre replicate the border (fraction forced to 0),
      rows are clamped with the fraction kept.
    A trailing singleton channel is dropped, as cv2 does.  Agreement with cv2 4.13:
    <= 3e-11 absolute on 0..255 data (operation order inside cv2's SIMD loops is not
    restated; the parity bar for floats is 1e-5 relative).
    """
    dw, dh = int(dsize[0]), int(dsize[1])
    src = np.asarray(src, dtype=np.float64)
    sh, sw = src.shape[:2]
    a3 = src.reshape(sh, sw, -1)
    if (sh, sw) == (dh, dw):
        out = a3.copy()
    elif sw == 2 * dw and sh == 2 * dh:
        out = (a3[0::2, 0::2] + a3[0::2, 1::2] + a3[1::2, 0::2] + a3[1::2, 1::2]) * 0.25
    else:
        sx, fx = _linear_axis(sw, dw)
        lo, hi = sx < 0, sx >= sw - 1
        sx[lo], fx[lo] = 0, 0.
        sx[hi], fx[hi] = sw - 1, 0.
        sx1 = np.minimum(sx + 1, sw - 1)
        hor = a3[:, sx] * (1. - fx)[None, :, None] + a3[:, sx1] * fx[None, :, None]
        sy, fy = _linear_axis(sh, dh)
        y0, y1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
        out = hor[y0] * (1. - fy)[:, None, None] + hor[y1] * fy[:, None, None]
    return out[:, :, 0] if out.shape[2] == 1 else out


# --------------------------------------------------------------------------------------
# loader.get_padded_img                          (reference loader.py:10-36)
# --------------------------------------------------------------------------------------

def _axis_window(n, crop, rng):
    """(out_begin, out_end, in_begin, in_end) along one axis; one ``randint`` draw."""
    if crop > n:
        o = int(rng.randint(0, crop - n + 1))
        return o, o + n, 0, n
    i = int(rng.randint(0, n - crop + 1))
    return 0, crop, i, i + crop


def get_padded_img(img, crop_h, crop_w, rng=np.random):
    """Canvas of max(crop, size) per axis; an axis shorter than the crop is placed at a random
    offset, a longer one is cut to a random crop-sized window and the rest of the canvas stays
    zero.  Rows are drawn before columns."""
    h, w = img.shape[:2]
    canvas = np.zeros((max(crop_h, h), max(crop_w, w), img.shape[2]), dtype=img.dtype)
    oi0, oi1, ii0, ii1 = _axis_window(h, crop_h, rng)
    oj0, oj1, ij0, ij1 = _axis_window(w, crop_w, rng)
    canvas[oi0:oi1, oj0:oj1] = img[ii0:ii1, ij0:ij1]
    return canvas


# --------------------------------------------------------------------------------------
# loader.load_and_crop / simple_load_crop / video_load_crop on decoded arrays
# (reference loader.py:39-85, 119-157, 285-330)
# --------------------------------------------------------------------------------------

def _crop_sample(fg_bgra, bg_bgr, extra, input_size, rng):
    """Shared body of the three loaders.  ``extra`` is None or a float64 (H,W,C) plane set
    that is padded / cropped / resized with the foreground (trimap: C=1; warped alpha: C=3).
    Returns (cmp, bg, alpha, extra, fg) after the resize, before mean subtraction."""
    alpha, fg = O.split_fg(fg_bgra)                               # reader.py:16-18
    fg = fg.astype(np.float64)
    bg = bg_bgr.astype(np.float64)
    crop_h, crop_w = CROP_TYPES[int(rng.randint(0, len(CROP_TYPES)))]
    if fg.shape[0] < crop_h or fg.shape[1] < crop_w:
        planes = [fg, alpha[:, :, None]] + ([extra] if extra is not None else [])
        canvas = get_padded_img(np.concatenate(planes, axis=2), crop_h, crop_w, rng)
        fg, alpha = canvas[:, :, :3], canvas[:, :, 3:4]
        if extra is not None:
            extra = canvas[:, :, 4:]
    i = int(rng.randint(0, fg.shape[0] - crop_h + 1))
    j = int(rng.randint(0, fg.shape[1] - crop_w + 1))
    win = (slice(i, i + crop_h), slice(j, j + crop_h))            # sic: crop_h for both axes
    fg, alpha = fg[win], alpha[win]
    if extra is not None:
        extra = extra[win]
    bch = int(np.ceil(crop_h * bg.shape[0] / fg.shape[0]))
    bcw = int(np.ceil(crop_w * bg.shape[1] / fg.shape[1]))
    padded = get_padded_img(bg, bch, bcw, rng)
    i = int(rng.randint(0, bg.shape[0] - bch + 1))
    j = int(rng.randint(0, bg.shape[1] - bcw + 1))
    bg = resize_linear_f64(padded[i:i + bch, j:j + bcw], input_size)
    fg = resize_linear_f64(fg, input_size)
    alpha = resize_linear_f64(alpha, input_size)
    if extra is not None:
        extra = resize_linear_f64(extra, input_size)
    cmp = O.create_composite_image(fg, bg, alpha)
    return cmp, bg, alpha, extra, fg


def video_sample(fg_bgra, bg_bgr, prev_bgra, flo, input_size, rng=np.random):
    """loader.video_load_crop (285-330) -> (cmp, bg, label, warped_alpha, fg)."""
    warped = O.warp_img(prev_bgra[:, :, 3] / 255., flo)          # loader.py:291-292
    warped = np.repeat(warped[:, :, None], 3, axis=2)
    cmp, bg, alpha, warped, fg = _crop_sample(fg_bgra, bg_bgr, warped, input_size, rng)
    return cmp - VGG_MEAN, bg - VGG_MEAN, alpha[:, :, None], warped, fg


def simple_sample(fg_bgra, bg_bgr, input_size, rng=np.random):
    """loader.simple_load_crop (119-157) -> (cmp, bg, label, fg)."""
    cmp, bg, alpha, _, fg = _crop_sample(fg_bgra, bg_bgr, None, input_size, rng)
    return cmp - VGG_MEAN, bg - VGG_MEAN, alpha[:, :, None], fg


def trimap_sample(fg_bgra, trimap_u8, bg_bgr, input_size, rng=np.random):
    """loader.load_and_crop (39-85) -> (inp (h,w,6), label, fg).  The trimap is padded,
    cropped and resized like the reference does but is not part of the returned input
    (loader.py:80-83 comment it out)."""
    tri = (trimap_u8 / 255.)[:, :, None]
    cmp, bg, alpha, _, fg = _crop_sample(fg_bgra, bg_bgr, tri, input_size, rng)
    return np.concatenate((cmp - VGG_MEAN, bg - VGG_MEAN), axis=2), alpha[:, :, None], fg


def psnr(img, img_ref):
    """loader.psnr (214-227): 10 log10(1 / (1e-6 + mean_px ||img - ref||_2^2))."""
    a = img.reshape(img.shape[0], img.shape[1], -1).astype(np.float64)
    b = img_ref.reshape(img_ref.shape[0], img_ref.shape[1], -1).astype(np.float64)
    return 10. * np.log10(1. / (1e-6 + np.mean(np.sum(np.square(a - b), axis=2))))


# --------------------------------------------------------------------------------------
# data.trimap_from_matte                          (reference data.py:37-67)
# --------------------------------------------------------------------------------------

def trimap_from_matte_loop(matte):
    """The reference's raster-order loop, restated literally (small images only): every pixel
    first receives its own class (255 / 0 / 128); a fractional pixel then paints 128 over the
    alpha==1 pixels within +-3 and the alpha==0 pixels within +-1 of it.  Because a pixel's own
    class is written when the scan reaches it, paint applied *before* that moment is lost."""
    assert matte.dtype == np.float64
    h, w = matte.shape
    out = np.zeros((h, w), dtype=np.uint8)
    for i in range(h):
        for j in range(w):
            m = matte[i, j]
            if m == 1.:
                out[i, j] = 255
            elif m == 0.:
                out[i, j] = 0
            else:
                out[i, j] = 128
                for k in range(max(0, i - 3), min(h, i + 4)):
                    for l in range(max(0, j - 3), min(w, j + 4)):
                        if matte[k, l] == 1.:
                            out[k, l] = 128
                        elif matte[k, l] == 0. and abs(k - i) <= 1 and abs(l - j) <= 1:
                            out[k, l] = 128
    return out


def trimap_from_matte(matte):
    """Closed form of the loop above: a pixel with alpha == 1 (== 0) becomes 128 iff a fractional
    pixel lies within Chebyshev distance 3 (1) of it AND comes later in raster order."""
    assert matte.dtype == np.float64
    h, w = matte.shape
    frac = (matte != 1.) & (matte != 0.)
    out = np.where(matte == 1., 255, np.where(matte == 0., 0, 128)).astype(np.uint8)
    pad = np.zeros((h + 6, w + 6), dtype=bool)
    pad[3:h + 3, 3:w + 3] = frac
    later = {1: np.zeros((h, w), dtype=bool), 3: np.zeros((h, w), dtype=bool)}
    for dk in range(-3, 4):
        for dl in range(-3, 4):
            if dk < 0 or (dk == 0 and dl <= 0):
                continue                                        # not later in raster order
            hit = pad[3 + dk:3 + dk + h, 3 + dl:3 + dl + w]
            later[3] |= hit
            if abs(dk) <= 1 and abs(dl) <= 1:
                later[1] |= hit
    out[(matte == 1.) & later[3]] = 128
    out[(matte == 0.) & later[1]] = 128
    return out

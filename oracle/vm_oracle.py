"""CPU oracle for the video-matting hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it (the
measurement helpers under ``scripts/`` and ``bench.py`` also take their seeded synthetic inputs
from the generators at the end of this file and time its functions as CPU baselines - never as
the thing shipped).  Nothing under ``video-matting_b200/`` imports it and the product path has
no CPU fallback.

It is a NumPy restatement of the per-frame pipeline of the reference
(``/root/reference/{flow,tps,augmentation,reader}.py``) *including* the arithmetic of the
third-party kernels those files bottom out in, none of which is vendored in the reference
and none of which is pinned by it (no requirements file).  The versions that define the
oracle in this image are: opencv-python-headless 4.13.0 (``cv2.remap``, ``cv2.warpAffine``,
``cv2.getRotationMatrix2D``, ``cv2.cvtColor``), scipy 1.18.1 (``ndimage.map_coordinates``),
numpy 2.3.5 (``linalg.pinv``, scalar promotion rules).  Their published algorithms are
restated here in integer / IEEE arithmetic (no call into cv2 or scipy from this file).

Pinning status: the reference has no tests and its ``forward.flo``/``backward.flo`` are
missing, so the only reference-owned known answers are ``test_data/cmp1.png``/``cmp2.png``
(pins read_fg_img's uint16 branch + create_composite_image, see
``tests/test_oracle_golden.py``).  Everything else is pinned by golden vectors generated
in the build container by importing the *unmodified* reference modules
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``) and, wherever cv2 / scipy are
importable, by differential tests against those libraries.

Every function cites the reference lines it follows.
"""
import numpy as np

INT_MIN = -(1 << 31)
FLO_MAGIC = np.float32(202021.25)

# --------------------------------------------------------------------------------------
# A.1  tap rule shared by cv2.remap(INTER_LINEAR) and cv2.warpAffine (legacy fixed point)
# --------------------------------------------------------------------------------------

def _gather_tap(src, iy, ix):
    """src[iy, ix] with BORDER_CONSTANT 0 for every index outside the image."""
    H, W = src.shape[:2]
    ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
    v = src[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)]
    if src.ndim == 3:
        ok = ok[..., None]
    return np.where(ok, v, np.zeros((), dtype=src.dtype))


def sample_fixed32(src, SX, SY):
    """Bilinear sample of ``src`` at fixed-point positions (SX, SY) in 1/32 px.

    OpenCV ``remapBilinear`` (imgwarp.cpp): integer pixel = value >> 5 saturated to int16,
    5-bit fraction selects a weight-table row.  uint8: int16 weights summing to 32768 and
    ``(sum + 16384) >> 15``; float/double: float32 table weights, products and the
    left-to-right sum in the source precision.
    """
    SX = np.asarray(SX, dtype=np.int64)
    SY = np.asarray(SY, dtype=np.int64)
    ix = np.clip(SX >> 5, -32768, 32767)
    iy = np.clip(SY >> 5, -32768, 32767)
    fx = SX & 31
    fy = SY & 31
    s00 = _gather_tap(src, iy, ix)
    s01 = _gather_tap(src, iy, ix + 1)
    s10 = _gather_tap(src, iy + 1, ix)
    s11 = _gather_tap(src, iy + 1, ix + 1)
    if src.ndim == 3:
        fx = fx[..., None]
        fy = fy[..., None]
    if src.dtype == np.uint8:
        w00 = (32 - fx) * (32 - fy) * 32
        w01 = fx * (32 - fy) * 32
        w10 = (32 - fx) * fy * 32
        w11 = fx * fy * 32
        acc = (s00.astype(np.int64) * w00 + s01.astype(np.int64) * w01 +
               s10.astype(np.int64) * w10 + s11.astype(np.int64) * w11)
        return ((acc + 16384) >> 15).astype(np.uint8)
    if src.dtype not in (np.float32, np.float64):
        raise TypeError("oracle restates remap for uint8/float32/float64 only")
    f32 = np.float32
    ax = fx.astype(f32) / f32(32)
    ay = fy.astype(f32) / f32(32)
    w00 = (f32(1) - ax) * (f32(1) - ay)          # exact in float32 (5-bit fractions)
    w01 = ax * (f32(1) - ay)
    w10 = (f32(1) - ax) * ay
    w11 = ax * ay
    T = src.dtype.type
    acc = s00 * w00.astype(T)
    acc = acc + s01 * w01.astype(T)
    acc = acc + s10 * w10.astype(T)
    acc = acc + s11 * w11.astype(T)
    return acc.astype(src.dtype)


def _cvround_x32(m):
    """cvRound(m * 32) for a float32 map: float32 product, round-half-even, and the x86
    ``cvtps2dq`` "integer indefinite" (INT_MIN) for NaN / out-of-range."""
    with np.errstate(over='ignore', invalid='ignore'):
        v = (m.astype(np.float32) * np.float32(32)).astype(np.float64)
        r = np.rint(v)
    bad = ~np.isfinite(v) | (r >= 2147483648.0) | (r < -2147483648.0)
    return np.where(bad, float(INT_MIN), r).astype(np.int64)


# --------------------------------------------------------------------------------------
# a-1 / a-2  flow.warp_img, flow.warp_bgr          (reference flow.py:9-18, 21-33)
# --------------------------------------------------------------------------------------

def flow_map(flow):
    """flow.py:12-17: ``(identity_int64 + flow_f32).astype(float32)`` (float64 sum, then
    one rounding to float32)."""
    h, w = flow.shape[:2]
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    mx = (jj.astype(np.float64) + flow[..., 0].astype(np.float64)).astype(np.float32)
    my = (ii.astype(np.float64) + flow[..., 1].astype(np.float64)).astype(np.float32)
    return mx, my


def warp_img(img, flow):
    """flow.py:9-18.  Single-channel bilinear backward warp: out[p] = img[p + flow[p]]."""
    assert img.ndim == 2                                     # flow.py:11
    mx, my = flow_map(flow)
    return sample_fixed32(np.ascontiguousarray(img), _cvround_x32(mx), _cvround_x32(my))


def warp_bgr(img, flow):
    """flow.py:21-33.  Three independent single-channel remaps, re-interleaved."""
    mx, my = flow_map(flow)
    return sample_fixed32(np.ascontiguousarray(img[:, :, :3]), _cvround_x32(mx), _cvround_x32(my))


# --------------------------------------------------------------------------------------
# a-3  flow.correct_alpha                           (reference flow.py:36-65)
# --------------------------------------------------------------------------------------

def occlusion_mask(backward, forward):
    """Boolean (H,W) mask of the pixels flow.py:41-50 zeroes (``err > 15``).

    flow.py:44/47 add a float32 flow component to a Python int (float32 result under
    numpy>=2), truncate toward zero with ``int()``, clamp from above only; negative indices
    wrap Python-style; an index below -H / -W raises IndexError like the reference.
    """
    h, w = backward.shape[:2]
    f32 = np.float32
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    with np.errstate(invalid='ignore', over='ignore'):
        a = backward[..., 0].astype(f32) + jj.astype(f32)
        b = backward[..., 1].astype(f32) + ii.astype(f32)
    if not (np.isfinite(a).all() and np.isfinite(b).all()):
        raise ValueError("cannot convert float NaN/inf to integer")   # int(nan) in the loop
    j0 = np.minimum(np.trunc(a.astype(np.float64)).astype(np.int64), w - 1)
    i0 = np.minimum(np.trunc(b.astype(np.float64)).astype(np.int64), h - 1)
    if (j0 < -w).any() or (i0 < -h).any():
        raise IndexError("index out of bounds in forward[i0, j0]")
    jw = np.where(j0 < 0, j0 + w, j0)
    iw = np.where(i0 < 0, i0 + h, i0)
    f = forward[iw, jw]
    with np.errstate(invalid='ignore', over='ignore'):
        c = f[..., 0].astype(f32) + j0.astype(f32)
        d = f[..., 1].astype(f32) + i0.astype(f32)
    if not (np.isfinite(c).all() and np.isfinite(d).all()):
        raise ValueError("cannot convert float NaN/inf to integer")
    j1 = np.minimum(np.trunc(c.astype(np.float64)).astype(np.int64), w - 1)
    i1 = np.minimum(np.trunc(d.astype(np.float64)).astype(np.int64), h - 1)
    di = (i1 - ii).astype(np.float64)
    dj = (j1 - jj).astype(np.float64)
    return (di * di + dj * dj) > 225.0


def correct_alpha(backward, forward, alpha):
    """flow.py:36-65 without the print/imshow side effects.  In place; returns ``alpha``."""
    alpha[occlusion_mask(backward, forward)] = 0.
    return alpha


def correct_alpha_loop(backward, forward, alpha):
    """Literal per-pixel restatement of flow.py:41-50 (pure-Python; small inputs only)."""
    h, w = backward.shape[:2]
    err = np.zeros((h, w))
    for i in range(h):
        for j in range(w):
            bx, by = backward[i, j]
            j0, i0 = min(int(bx + j), w - 1), min(int(by + i), h - 1)
            fx, fy = forward[i0, j0]
            j1, i1 = min(int(fx + j0), w - 1), min(int(fy + i0), h - 1)
            err[i, j] = np.sqrt(float((i1 - i) ** 2 + (j1 - j) ** 2))
    alpha[err > 15.] = 0.
    return alpha


# --------------------------------------------------------------------------------------
# a-4  reader.read_flow                             (reference reader.py:21-30)
# --------------------------------------------------------------------------------------

def parse_flo(buf):
    """Middlebury .flo from bytes: float32 magic, int32 w, int32 h, 2*h*w float32.
    Returns (flow, magic_ok).  A short payload raises ValueError (reshape), reader.py:29."""
    key = np.frombuffer(buf, dtype=np.float32, count=1, offset=0)
    w = int(np.frombuffer(buf, dtype=np.int32, count=1, offset=4)[0])
    h = int(np.frombuffer(buf, dtype=np.int32, count=1, offset=8)[0])
    avail = (len(buf) - 12) // 4
    data = np.frombuffer(buf, dtype=np.float32, count=min(avail, 2 * h * w), offset=12)
    return data.reshape((h, w, 2)), bool(key[0] == FLO_MAGIC)


def write_flo(path, flow):
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    h, w = flow.shape[:2]
    with open(path, 'wb') as f:
        np.array([FLO_MAGIC], dtype=np.float32).tofile(f)
        np.array([w, h], dtype=np.int32).tofile(f)
        flow.tofile(f)


# --------------------------------------------------------------------------------------
# a-13 reader.read_fg_img uint16 branch             (reference reader.py:13-18)
# --------------------------------------------------------------------------------------

def fg_from_uint16(img16):
    """reader.py:13-15: ``(((img+1)/256.)-1).astype(uint8)``; ``img+1`` wraps in uint16 and
    the float->uint8 cast of -1.0 wraps to 255 (x86)."""
    t = ((img16.astype(np.uint32) + 1) & 0xFFFF).astype(np.float64) / 256. - 1.
    return (np.trunc(t).astype(np.int64) & 0xFF).astype(np.uint8)


def split_fg(img8):
    """reader.py:16-18."""
    return img8[:, :, 3] / 255., img8[:, :, :3]


# --------------------------------------------------------------------------------------
# a-12 reader.create_composite_image                (reference reader.py:72-79)
# --------------------------------------------------------------------------------------

def create_composite_image(fg, bg, alpha):
    tri = np.zeros(fg.shape, dtype=np.float64)
    tri[:, :, 0] = alpha
    tri[:, :, 1] = alpha
    tri[:, :, 2] = alpha
    return tri * fg + (1. - tri) * bg


# --------------------------------------------------------------------------------------
# a-5  TPS coefficients                             (reference tps.py:78-98, 113-119)
# --------------------------------------------------------------------------------------

def _tps_U(r):
    """tps.py:78-82."""
    with np.errstate(divide='ignore', invalid='ignore'):
        return (r ** 2) * np.where(r < 1e-100, 0, np.log(r))


def tps_system(points):
    """tps.py:85-98: L = [[K, P], [P^T, 0]] with K_ab = U(|P_a - P_b|), P = [1 | points]."""
    points = np.asarray(points, dtype=np.float64)
    n = len(points)
    xd = np.subtract.outer(points[:, 0], points[:, 0])
    yd = np.subtract.outer(points[:, 1], points[:, 1])
    K = _tps_U(np.sqrt(xd ** 2 + yd ** 2))
    L = np.zeros((n + 3, n + 3))
    L[:n, :n] = K
    L[:n, n] = 1.
    L[:n, n + 1:] = points
    L[n, :n] = 1.
    L[n + 1:, :n] = points.T
    return L


def tps_coefficients(from_points, to_points):
    """tps.py:113-119 for ``_make_warp(from_points, to_points, ...)``: (N+3, 2) float64.
    The truncated pseudo-inverse is numpy's own (LAPACK SVD, rcond 1e-15), as in the
    reference - it is called, not restated."""
    from_points = np.asarray(from_points, dtype=np.float64)
    to_points = np.asarray(to_points, dtype=np.float64)
    n = len(to_points)
    v = np.zeros((n + 3, 2))
    v[:n] = to_points                                        # np.resize then v[-3:] = 0
    return np.dot(np.linalg.pinv(tps_system(from_points)), v)


def tps_eval(coeffs, points, x, y):
    """tps.py:101-110: a1 + ax*x + ay*y + sum_i w_i U(|(x,y) - P_i|), sequential in i."""
    w = coeffs[:-3]
    a1, ax, ay = coeffs[-3:]
    s = np.zeros(x.shape)
    for wi, Pi in zip(w, points):
        s += wi * _tps_U(np.sqrt((x - Pi[0]) ** 2 + (y - Pi[1]) ** 2))
    return a1 + ax * x + ay * y + s


# --------------------------------------------------------------------------------------
# a-6  tps._make_inverse_warp / tps.warp_images     (reference tps.py:14-34, 41-75)
# --------------------------------------------------------------------------------------

def tps_coarse_axes(h, w, approximate_grid=2):
    """tps.py:45-47: ``np.mgrid[0:h:xs*1j, 0:w:ys*1j]`` -> int(xs) x int(ys) points,
    both ends inclusive (np.linspace semantics of a complex step)."""
    xs = h / approximate_grid
    ys = w / approximate_grid
    nx, ny = int(xs), int(ys)
    cx = np.mgrid[0:h:xs * 1j]
    cy = np.mgrid[0:w:ys * 1j]
    assert cx.shape == (nx,) and cy.shape == (ny,)
    return xs, ys, cx, cy


def tps_inverse_transform(from_points, to_points, h, w, approximate_grid=2):
    """tps.py:41-75 for output_region (0, 0, h, w): returns [t_row, t_col], each (h+1, w+1),
    plus the coarse transform for inspection."""
    xs, ys, cx, cy = tps_coarse_axes(h, w, approximate_grid)
    X, Y = np.meshgrid(cx, cy, indexing='ij')
    # tps.py:51 calls _make_warp(to_points, from_points, ...): the system is built from the
    # *deformed* grid and maps back onto the regular one.
    C = tps_coefficients(to_points, from_points)
    P = np.asarray(to_points, dtype=np.float64)
    T0 = tps_eval(C[:, 0], P, X, Y)
    T1 = tps_eval(C[:, 1], P, X, Y)
    ni, nj = np.mgrid[0:h + 1, 0:w + 1]
    xf, xi = np.modf((xs - 1) * ni / float(h))
    yf, yi = np.modf((ys - 1) * nj / float(w))
    xi = xi.astype(int)
    yi = yi.astype(int)
    x1 = 1 - xf
    y1 = 1 - yf
    ix1 = (xi + 1).clip(0, xs - 1).astype(int)
    iy1 = (yi + 1).clip(0, ys - 1).astype(int)
    out = []
    for T in (T0, T1):
        t00, t01, t10, t11 = T[xi, yi], T[xi, iy1], T[ix1, yi], T[ix1, iy1]
        out.append(t00 * x1 * y1 + t01 * x1 * yf + t10 * xf * y1 + t11 * xf * yf)
    return out, (T0, T1), C


def map_coordinates_linear(img, t0, t1):
    """scipy.ndimage.map_coordinates(order=1, mode='constant', cval=0) (tps.py:34).

    Any coordinate outside [0, n-1] (exact comparison) -> 0.  float64 bilinear
    ``s00(1-a)(1-b) + s01(1-a)b + s10 a(1-b) + s11 ab``; uint8 output = floor(v + 0.5)
    clamped (round half up); output dtype follows the input.
    """
    img = np.asarray(img)
    H, W = img.shape
    with np.errstate(invalid='ignore'):
        inside = (t0 >= 0) & (t0 <= H - 1) & (t1 >= 0) & (t1 <= W - 1)
    t0c = np.where(inside, t0, 0.)
    t1c = np.where(inside, t1, 0.)
    i0 = np.floor(t0c).astype(np.int64)
    j0 = np.floor(t1c).astype(np.int64)
    a = t0c - i0
    b = t1c - j0
    i1 = np.minimum(i0 + 1, H - 1)
    j1 = np.minimum(j0 + 1, W - 1)
    src = img.astype(np.float64)
    # scipy's get_spline_interpolation_weights (order 1): w0 = 1 - a, w1 = 1 - w0 (which is
    # not bit-equal to a when a < 0.5 carries more than 53 bits below 1.0, i.e. in row/col 0)
    a0 = 1 - a
    a1 = 1 - a0
    b0 = 1 - b
    b1 = 1 - b0
    v = (src[i0, j0] * a0 * b0 + src[i0, j1] * a0 * b1 +
         src[i1, j0] * a1 * b0 + src[i1, j1] * a1 * b1)
    v = np.where(inside, v, 0.)
    if img.dtype == np.uint8:
        return np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)
    return v.astype(img.dtype)


def map_coordinates_nearest(img, t0, t1):
    """scipy.ndimage.map_coordinates(order=0, mode='constant', cval=0) (tps.py:34 with interpolation_order=0,
    "if 0 then use nearest-neighbor", tps.py:22): outside [0, n-1] -> 0, else the sample at floor(t + 1/2)
    (probed against scipy 1.18: 0.5 -> 1, 1.5 -> 2, 2.5 -> 3, -1e-4 -> cval, n-1+1e-7 -> cval)."""
    img = np.asarray(img)
    H, W = img.shape
    with np.errstate(invalid='ignore'):
        inside = (t0 >= 0) & (t0 <= H - 1) & (t1 >= 0) & (t1 <= W - 1)
    i = np.floor(np.where(inside, t0, 0.) + 0.5).astype(np.int64)
    j = np.floor(np.where(inside, t1, 0.) + 0.5).astype(np.int64)
    return np.where(inside, img[i, j], 0).astype(img.dtype)


def tps_warp_images(from_points, to_points, images, output_region, interpolation_order=1,
                    approximate_grid=2):
    """tps.py:14-34 (nearest or linear interpolation, region (0,0,h,w))."""
    x_min, y_min, h, w = output_region
    assert x_min == 0 and y_min == 0 and interpolation_order in (0, 1)
    (t0, t1), _, _ = tps_inverse_transform(from_points, to_points, h, w, approximate_grid)
    fn = map_coordinates_linear if interpolation_order == 1 else map_coordinates_nearest
    return [fn(im, t0, t1) for im in images]


# --------------------------------------------------------------------------------------
# a-7  augmentation.deform_grid                     (reference augmentation.py:24-41)
# --------------------------------------------------------------------------------------

def deform_grid(h, w, n=5, rng=None):
    """Regular n x n grid over [0,h]x[0,w] (rows are (y, x)); per point, in row-major
    order: one uniform(-b, b) draw for x if 0<x<w, then one for y if 0<y<h.
    ``rng`` defaults to the global ``np.random`` stream like the reference."""
    rng = np.random if rng is None else rng
    bound = min(w, h) * 0.05
    vy = (h / (n - 1)) * np.arange(n)
    vx = (w / (n - 1)) * np.arange(n)
    grid = np.transpose([np.repeat(vy, n), np.tile(vx, n)])
    new = grid.copy()
    for k in range(n * n):
        y, x = grid[k]
        if 0. < x < w:
            new[k, 1] += rng.uniform(-bound, bound)
        if 0. < y < h:
            new[k, 0] += rng.uniform(-bound, bound)
    return grid, new


# --------------------------------------------------------------------------------------
# a-8  cv2.warpAffine / getRotationMatrix2D / augmentation.warp_image
#                                                   (reference augmentation.py:44-63)
# --------------------------------------------------------------------------------------

def rotation_matrix_2d(center, angle_deg, scale):
    """cv2.getRotationMatrix2D: float64 2x3."""
    ang = angle_deg * (np.pi / 180.)
    al = np.cos(ang) * scale
    be = np.sin(ang) * scale
    cx, cy = float(center[0]), float(center[1])
    return np.array([[al, be, (1 - al) * cx - be * cy],
                     [-be, al, be * cx + (1 - al) * cy]], dtype=np.float64)


def affine_inverse(M):
    """OpenCV's closed-form inverse used by warpAffine when WARP_INVERSE_MAP is not set."""
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1. / D if D != 0 else 0.
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    i00, i01, i10, i11 = A11, M[0, 1] * (-D), M[1, 0] * (-D), A22
    b0 = -i00 * M[0, 2] - i01 * M[1, 2]
    b1 = -i10 * M[0, 2] - i11 * M[1, 2]
    return np.array([[i00, i01, b0], [i10, i11, b1]], dtype=np.float64)


def _sat_i32_round(v):
    r = np.rint(np.asarray(v, dtype=np.float64))
    bad = ~np.isfinite(r) | (r >= 2147483648.0) | (r < -2147483648.0)
    return np.where(bad, float(INT_MIN), r).astype(np.int64)


def affine_fixed_coords(M, dh, dw):
    """Fixed-point (1/32 px) source coordinates of cv2.warpAffine's legacy path:
    AB_BITS = 10, adelta/bdelta per column, X0/Y0 per row with the +16 rounding offset."""
    I = affine_inverse(M)
    x = np.arange(dw, dtype=np.float64)
    y = np.arange(dh, dtype=np.float64)
    adelta = _sat_i32_round(I[0, 0] * x * 1024)
    bdelta = _sat_i32_round(I[1, 0] * x * 1024)
    X0 = _sat_i32_round((I[0, 1] * y + I[0, 2]) * 1024) + 16
    Y0 = _sat_i32_round((I[1, 1] * y + I[1, 2]) * 1024) + 16
    SX = (X0[:, None] + adelta[None, :]) >> 5
    SY = (Y0[:, None] + bdelta[None, :]) >> 5
    return SX, SY


def warp_affine(src, M, dsize):
    """cv2.warpAffine(src, M, (dw, dh)) with default flags (INTER_LINEAR, BORDER_CONSTANT 0)."""
    dw, dh = dsize
    SX, SY = affine_fixed_coords(M, dh, dw)
    return sample_fixed32(np.ascontiguousarray(src), SX, SY)


def warp_image(img, params, thin=None):
    """augmentation.py:44-63: optional TPS, then integer translate, then rotate/scale."""
    (tu, tv), rot, scale, center = params
    h, w = img.shape[:2]
    if thin is not None:
        grid, def_grid = thin
        if img.ndim == 3 and img.shape[2] == 3:
            chans = tps_warp_images(grid, def_grid, [img[:, :, 0], img[:, :, 1], img[:, :, 2]],
                                    (0, 0, h, w))
            img = np.transpose(chans, axes=(1, 2, 0)).copy()
        else:
            img = tps_warp_images(grid, def_grid, [img], (0, 0, h, w))[0]
    mt = np.float32([[1, 0, tu], [0, 1, tv]])
    translated = warp_affine(img, mt, (w, h))
    return warp_affine(translated, rotation_matrix_2d(center, rot, scale), (w, h))


# --------------------------------------------------------------------------------------
# a-9  augmentation.change_illumination             (reference augmentation.py:88-99)
# --------------------------------------------------------------------------------------

def _div_table(scale, n=256):
    t = np.zeros(n, dtype=np.int64)
    k = np.arange(1, n)
    t[1:] = np.rint(scale / k.astype(np.float64)).astype(np.int64)
    return t


_SDIV = _div_table(255 * 4096.)
_HDIV180 = _div_table(180 * 4096. / 6.)


def bgr2hsv_u8(bgr):
    """cv2.cvtColor(COLOR_BGR2HSV) for uint8 (hrange 180): OpenCV's integer table model
    (hsv_shift = 12), exact over all 2^24 colours."""
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    d = v - vmin
    vr = (v == r)
    vg = (v == g)
    hh = np.where(vr, g - b, np.where(vg, b - r + 2 * d, r - g + 4 * d))
    s = (d * _SDIV[v] + (1 << 11)) >> 12
    hh = (hh * _HDIV180[d] + (1 << 11)) >> 12
    hh = np.where(hh < 0, hh + 180, hh)
    return np.stack([hh, s, v], axis=-1).astype(np.uint8)


def _fma32(a, b, c):
    """float32 fused multiply-add (one rounding): exact product in float64, one add, one cast."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


HSV_VEC = 32      # pixels per SIMD step of cv2's HSV2BGR on the machine that made the fixtures (AVX2: v_uint8x32)


def hsv2bgr_u8(hsv, vec=HSV_VEC):
    """cv2.cvtColor(COLOR_HSV2BGR) for uint8 images (H, W, 3), bit-exact for opencv-python-headless 4.13
    [probed over all 180 x 256 x 256 inputs, 0 mismatches on either path]:

        s = S * (1/255.f); v = V * (1/255.f); hh = H * (6.f/180.f); sector = floor(hh); f = hh - sector
        tab = {v, v*(1-s), v*fma(-s, f, 1), v*fma(-s, 1-f, 1)}         (float32; the compiler contracts 1 - s*f)
        {b, g, r} = tab[sector_data[sector]];  s == 0 -> b = g = r = v
        out = tab * 255.f, then  - TRUNCATED toward zero in the SIMD body of a row (pixels x < W - W % vec),
                                 - rounded half-to-even (cvRound) in the scalar tail of the row (the last W % vec pixels).

    ``vec`` is the SIMD width cv2 dispatches to on the host (32 with AVX2, 16 with SSE only, 64 with AVX-512);
    rows are processed independently (cvtColor does not flatten continuous images)."""
    f32 = np.float32
    hsv = np.asarray(hsv)
    assert hsv.ndim == 3 and hsv.shape[2] == 3
    W = hsv.shape[1]
    h = (hsv[..., 0].astype(f32) * f32(f32(6.) / f32(180.))).astype(f32)
    s = (hsv[..., 1].astype(f32) * f32(1. / 255.)).astype(f32)
    v = (hsv[..., 2].astype(f32) * f32(1. / 255.)).astype(f32)
    sector = np.floor(h)
    hf = (h - sector).astype(f32)
    sec = sector.astype(np.int64)
    oob = (sec < 0) | (sec >= 6)
    sec = np.where(oob, 0, sec)
    hf = np.where(oob, f32(0), hf)
    one = np.ones_like(s)
    t0 = v
    t1 = (v * (f32(1) - s)).astype(f32)
    t2 = (v * _fma32(-s, hf, one)).astype(f32)
    t3 = (v * _fma32(-s, (f32(1) - hf).astype(f32), one)).astype(f32)
    tab = np.stack([t0, t1, t2, t3], axis=-1)
    idx = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    sel = idx[sec]
    b = np.take_along_axis(tab, sel[..., 0:1], -1)[..., 0]
    g = np.take_along_axis(tab, sel[..., 1:2], -1)[..., 0]
    r = np.take_along_axis(tab, sel[..., 2:3], -1)[..., 0]
    grey = hsv[..., 1] == 0
    b = np.where(grey, v, b)
    g = np.where(grey, v, g)
    r = np.where(grey, v, r)
    out = (np.stack([b, g, r], axis=-1) * f32(255)).astype(f32)
    body = (np.arange(W) < W - W % int(vec))[None, :, None]
    return np.clip(np.where(body, np.trunc(out), np.rint(out)), 0, 255).astype(np.uint8)


def probe_hsv_vec():
    """SIMD width of cv2's HSV2BGR on this host: the pixel (H, S, V) = (0, 1, 1) becomes (0, 0, 1) in the
    truncating SIMD body and (1, 1, 1) in the rounding tail; a row of 255 pixels has 255 % vec tail pixels."""
    import cv2
    row = np.tile(np.array([0, 1, 1], np.uint8), (1, 255, 1))
    tail = int((cv2.cvtColor(row, cv2.COLOR_HSV2BGR)[0, :, 0] == 1).sum())
    return tail + 1


# --------------------------------------------------------------------------------------
# cv2.resize(uint8, INTER_LINEAR)            (reference reader.py:41,53, augmentation.py:160)
# --------------------------------------------------------------------------------------

def _resize_axis_u8(dn, sn, clamp_frac):
    """OpenCV resize.cpp (8-bit INTER_LINEAR, INTER_RESIZE_COEF_BITS = 11): scale = 1. / (dn / sn) in double,
    f = (float)((d + 0.5) * scale - 0.5), i = floor(f), f -= i; horizontally the fraction is forced to 0 where
    the tap pair would leave the row (i < 0 or i >= sn - 1) - vertically it is NOT (rows are clipped instead,
    both taps then read the same row with their two separately truncated weights); coefficients
    cvRound((1 - f) * 2048), cvRound(f * 2048) from float32 products."""
    scale = 1.0 / (float(dn) / float(sn))
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    i = np.floor(f).astype(np.int64)
    f = (f - i.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = i < 0
        f[lo] = 0
        i[lo] = 0
        hi = i >= sn - 1
        f[hi] = 0
        i[hi] = sn - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    w1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return i, w0, w1


def resize_linear_u8(src, dsize):
    """``cv2.resize(src uint8 (H, W[, C]), dsize=(width, height), interpolation=cv2.INTER_LINEAR)``, bit-exact for
    opencv-python-headless 4.13 [probed: 0 mismatches on random images, up- and down-scaling, 1 and 3 channels,
    optimisations on or off].  Horizontal pass: int32 rows S = s[i]*a0 + s[i+1]*a1 (11-bit coefficients);
    vertical pass: (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  Exact 2x reductions in both
    directions are INTER_AREA in OpenCV (2x2 block mean, (a + b + c + d + 2) >> 2)."""
    src = np.asarray(src)
    assert src.dtype == np.uint8
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = src.shape[:2]
    s = src.reshape(sh, sw, -1).astype(np.int64)
    if sw == 2 * dw and sh == 2 * dh:
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8).reshape((dh, dw) + src.shape[2:])
    xi, a0, a1 = _resize_axis_u8(dw, sw, True)
    yi, b0, b1 = _resize_axis_u8(dh, sh, False)
    x1 = np.minimum(xi + 1, sw - 1)
    rows = s[:, xi] * a0[None, :, None] + s[:, x1] * a1[None, :, None]
    S0, S1 = rows[np.clip(yi, 0, sh - 1)], rows[np.clip(yi + 1, 0, sh - 1)]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + src.shape[2:])


def illumination_sv(x_u8, a, b, c):
    """augmentation.py:91-98 for one S or V plane: clip(a*(x/255)**b + c, 0, 1), then
    ``(255.*new).astype(uint8)`` (truncation)."""
    new = np.clip(a * np.power(x_u8 / 255., b) + c, 0., 1.)
    return (255. * new).astype(np.uint8)


def change_illumination(bgr, a, b, c, vec=HSV_VEC):
    """augmentation.py:88-99, bit-exact for a host whose cv2 converts `vec` pixels per SIMD step."""
    hsv = bgr2hsv_u8(bgr)
    out = hsv.copy()
    out[..., 1] = illumination_sv(hsv[..., 1], a, b, c)
    out[..., 2] = illumination_sv(hsv[..., 2], a, b, c)
    return hsv2bgr_u8(out, vec)


# --------------------------------------------------------------------------------------
# a-10 augmentation.object_size / fg_center         (reference augmentation.py:10-21)
# --------------------------------------------------------------------------------------

def object_size(alpha):
    return np.sqrt(np.count_nonzero(alpha != 0.))


def fg_center(alpha):
    nz = np.where(alpha != 0.)
    return int(np.mean(nz[1])), int(np.mean(nz[0]))


# --------------------------------------------------------------------------------------
# a-11 augmentation.augment                         (reference augmentation.py:102-135)
# --------------------------------------------------------------------------------------

def augment_params(h, w, alpha, rng=None, n=5):
    """The RNG contract of augment(): draws in the reference's order (global np.random by
    default) and returns every parameter the warps need."""
    rng = np.random if rng is None else rng
    fg_size = object_size(alpha)
    tu_bg = int(rng.uniform(-w * 0.05, w * 0.05))
    tv_bg = int(rng.uniform(-h * 0.05, h * 0.05))
    scale_bg = rng.uniform(1., 1.15)
    params_bg = ((tu_bg, tv_bg), 0., scale_bg, (w // 2, h // 2))
    grid, def_grid = deform_grid(h, w, n, rng)
    tu_fg = int(rng.uniform(-fg_size * 0.05, fg_size * 0.05))
    tv_fg = int(rng.uniform(-fg_size * 0.05, fg_size * 0.05))
    rot_fg = rng.uniform(-10, 10)
    scale_fg = rng.uniform(1., 1.15)
    params_fg = ((tu_fg, tv_fg), rot_fg, scale_fg, fg_center(alpha))
    a = rng.uniform(0.95, 1.05)
    b = rng.uniform(0.7, 1.3)
    c = rng.uniform(-0.07, 0.07)
    return params_bg, params_fg, (grid, def_grid), (a, b, c)


def augment(fg, bg, alpha, rng=None, vec=HSV_VEC):
    h, w = fg.shape[:2]
    params_bg, params_fg, grids, (a, b, c) = augment_params(h, w, alpha, rng)
    new_bg = warp_image(bg, params_bg)
    new_fg = warp_image(fg, params_fg, thin=grids)
    new_alpha = warp_image(alpha, params_fg, thin=grids)
    return change_illumination(new_fg, a, b, c, vec), change_illumination(new_bg, a, b, c, vec), new_alpha


# --------------------------------------------------------------------------------------
# SURVEY 8(d) "C4 oracle pipeline": flow warp + mask + TPS + composite
# --------------------------------------------------------------------------------------

def pipeline_c2(fg_bgra, flow_b, flow_f):
    """warp_bgr + warp_img + correct_alpha on a BGRA uint8 frame (alpha = A/255)."""
    alpha, bgr = split_fg(fg_bgra)
    b = warp_bgr(bgr, flow_b)
    a = warp_img(alpha, flow_b)
    a = correct_alpha(flow_b, flow_f, a)
    return b, a


def pipeline_c4(fg_bgra, flow_b, flow_f, grids, bg):
    h, w = fg_bgra.shape[:2]
    b, a = pipeline_c2(fg_bgra, flow_b, flow_f)
    ident = ((0, 0), 0., 1., (w // 2, h // 2))
    b2 = warp_image(b, ident, thin=grids)
    a2 = warp_image(a, ident, thin=grids)
    return create_composite_image(b2, bg, a2), a2


def pipeline_c3(fg_bgra, grids, bg):
    """TPS (no flow) + composite."""
    h, w = fg_bgra.shape[:2]
    alpha, bgr = split_fg(fg_bgra)
    ident = ((0, 0), 0., 1., (w // 2, h // 2))
    b2 = warp_image(np.ascontiguousarray(bgr), ident, thin=grids)
    a2 = warp_image(alpha, ident, thin=grids)
    return create_composite_image(b2, bg, a2), a2


# --------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY 8(d) (numpy only, so they can be generated on the GPU box)
# --------------------------------------------------------------------------------------

def _smooth_noise(rng, h, w, cell=32):
    """Smooth random field in ~N(0,1): bilinear up-sampling of a coarse normal grid."""
    gh, gw = h // cell + 2, w // cell + 2
    g = rng.standard_normal((gh, gw))
    yi = np.linspace(0, gh - 1, h)
    xi = np.linspace(0, gw - 1, w)
    y0 = np.minimum(np.floor(yi).astype(int), gh - 2)
    x0 = np.minimum(np.floor(xi).astype(int), gw - 2)
    fy = (yi - y0)[:, None]
    fx = (xi - x0)[None, :]
    return (g[y0][:, x0] * (1 - fy) * (1 - fx) + g[y0][:, x0 + 1] * (1 - fy) * fx +
            g[y0 + 1][:, x0] * fy * (1 - fx) + g[y0 + 1][:, x0 + 1] * fy * fx)


def synth_frame(seed, h, w):
    """BGRA uint8 frame: white-noise colour (worst case for interpolation parity) and a
    blobby alpha with ~1/3 zeros, ~1/3 ones, rest fractional."""
    rng = np.random.default_rng(seed)
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., :3] = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    out[..., 3] = np.clip(128 + 384 * _smooth_noise(rng, h, w), 0, 255).astype(np.uint8)
    return out


def synth_flows(seed, h, w):
    """(backward, forward) float32 flows: smooth ~8 px*(w/1920) backward flow; forward =
    -backward + 0.5 px noise, plus a rectangle (10 % of the area) offset by +25 px so the
    15 px consistency test fires."""
    rng = np.random.default_rng(seed + 7919)
    amp = 8.0 * (w / 1920.)
    back = np.stack([_smooth_noise(rng, h, w), _smooth_noise(rng, h, w)], -1) * amp
    fwd = -back + 0.5 * np.stack([_smooth_noise(rng, h, w), _smooth_noise(rng, h, w)], -1)
    rh, rw = int(h * 0.316), int(w * 0.316)
    r0, c0 = int(rng.integers(0, h - rh + 1)), int(rng.integers(0, w - rw + 1))
    fwd[r0:r0 + rh, c0:c0 + rw] += 25.0
    return back.astype(np.float32), fwd.astype(np.float32)


def synth_grids(seed, h, w, n=5):
    rng = np.random.RandomState(seed)
    return deform_grid(h, w, n, rng)


def synth_background(seed, h, w):
    rng = np.random.default_rng(seed + 104729)
    base = np.stack([_smooth_noise(rng, h, w, 16) for _ in range(3)], -1)
    return np.clip(128 + 64 * base + rng.integers(-20, 21, size=(h, w, 3)), 0, 255).astype(np.uint8)

"""Drop-in replacement for the reference ``augmentation`` module (reference
augmentation.py:10-166): random TPS + affine warps of foreground / alpha / background and
HSV illumination jitter, with the reference's global-``np.random`` draw order."""
import os

import numpy as np
import torch

from . import _native as N
from . import pipeline as P
from . import reader, tps


def _stats(alpha):
    a, _ = N.to_device(alpha)
    if a.dtype not in (torch.uint8, torch.float32, torch.float64):
        a = a.to(torch.float64)
    cnt, si, sj = (int(v) for v in P.alpha_stats(a).cpu())
    return cnt, si, sj


def object_size(alpha):
    """typical side of foreground image - reference augmentation.py:10-14."""
    return np.sqrt(_stats(alpha)[0])


def fg_center(alpha):
    """barycenter (x, y) of the foreground - reference augmentation.py:17-21."""
    cnt, si, sj = _stats(alpha)
    if cnt == 0:
        raise ValueError("cannot convert float NaN to integer")       # int(np.mean([]))
    return int(sj / cnt), int(si / cnt)


def deform_grid(h, w, n=5):
    """regular n x n grid and a randomly perturbed copy - reference augmentation.py:24-41.
    Draws from the global np.random stream: x first (if the point is not on a vertical edge),
    then y (if not on a horizontal edge), row-major over the grid."""
    bound = min(w, h) * 0.05
    rows = (h / (n - 1)) * np.arange(n)
    cols = (w / (n - 1)) * np.arange(n)
    grid = np.transpose([np.repeat(rows, n), np.tile(cols, n)])
    moved = grid.copy()
    # one vector draw = the reference's scalar draws in the same order (legacy RandomState.uniform consumes
    # one double per value either way): per point, x if 0 < x < w, then y if 0 < y < h
    need = np.stack([(grid[:, 1] > 0.) & (grid[:, 1] < w), (grid[:, 0] > 0.) & (grid[:, 0] < h)], axis=1)   # (n*n, {x, y})
    draws = np.random.uniform(-bound, bound, size=int(need.sum()))
    flat = np.zeros(need.shape)
    flat[need] = draws                                               # row-major over (point, {x, y}): the reference's order
    moved[:, 1] += flat[:, 0]
    moved[:, 0] += flat[:, 1]
    return grid, moved


def _rotation_matrix(center, angle, scale):
    """cv2.getRotationMatrix2D in float64."""
    ang = angle * (np.pi / 180.)
    al, be = np.cos(ang) * scale, np.sin(ang) * scale
    cx, cy = float(center[0]), float(center[1])
    return np.array([[al, be, (1 - al) * cx - be * cy], [-be, al, be * cx + (1 - al) * cy]])


def warp_image(img, params, thin=None):
    """warp image according to given parameters - reference augmentation.py:44-63:
    optional TPS deformation, integer translation, then rotation/scale about ``center``."""
    (tu, tv), rot, scale, center = params
    src, kind = N.to_device(img)
    h, w = src.shape[:2]
    if thin is not None:
        grid, def_grid = thin
        plan = P.get_plan((0, 0, h, w), 2, src.device)
        coarse = tps._coarse_for(grid, def_grid, plan)
        if src.dim() == 3 and src.shape[2] != 3:
            raise RuntimeError("invalid shape for coordinate array")
        src = P.tps_warp(src, coarse, plan)                           # (h+1, w+1[, 3])
    mt = np.float32([[1, 0, tu], [0, 1, tv]])
    translated = P.warp_affine(src, mt, (w, h))
    rotated = P.warp_affine(translated, _rotation_matrix(center, rot, scale), (w, h))
    return N.from_device(rotated, kind)


def identity(m, n):
    """array s.t. arr[i, j] = [i+1, j+1] - reference augmentation.py:66-70."""
    v1, v2 = np.arange(1, n + 1), np.arange(1, m + 1)
    return np.transpose([np.repeat(v2, n), np.tile(v1, m)]).reshape(m, n, 2)


def synthetize_flow(fg_params, bg_params, grids, warped_alpha):
    """reference augmentation.py:73-85.  Kept for API completeness: like the reference it
    fails (RuntimeError) because a 2-channel image cannot go through the TPS path."""
    h, w = warped_alpha.shape[:2]
    id_fg, id_bg = identity(h, w), identity(h, w)
    w_fg = warp_image(id_fg, fg_params, thin=grids)
    w_bg = warp_image(id_bg, bg_params)
    bi_alpha = np.zeros((h, w, 2), dtype=float)
    bi_alpha[:, :, 0] = warped_alpha
    bi_alpha[:, :, 1] = warped_alpha
    return np.multiply(bi_alpha, w_fg - id_fg) + np.multiply(1. - bi_alpha, w_bg - id_bg)


def illumination_lut(a, b, c):
    """S/V transfer table of change_illumination: the reference expression
    (augmentation.py:91-98) evaluated by numpy on the 256 possible uint8 inputs."""
    x = np.arange(256, dtype=np.uint8)
    new = np.clip(a * np.power(x / 255., b) + c, 0., 1.)
    return (255. * new).astype(np.uint8)


def change_illumination(bgr, a, b, c):
    """randomly changes illumination via [H]SV transformation - reference augmentation.py:88-99."""
    src, kind = N.to_device(bgr)
    if src.dtype != torch.uint8:
        raise TypeError("change_illumination expects a uint8 BGR image")
    return N.from_device(P.illumination(src, illumination_lut(a, b, c)), kind)


def augment(fg, bg, alpha):
    """randomly modify input image to create synthetic data - reference augmentation.py:102-135.
    Same 40 global np.random draws, in the same order, as the reference."""
    bound_translate, bound_rotate, bound_scale = 0.05, 10, 0.15
    h, w = fg.shape[:2]
    a_dev, _ = N.to_device(alpha)
    fg_size = object_size(a_dev)
    tu_bg = int(np.random.uniform(-w * bound_translate, w * bound_translate))
    tv_bg = int(np.random.uniform(-h * bound_translate, h * bound_translate))
    scale_bg = np.random.uniform(1., 1. + bound_scale)
    new_bg = warp_image(bg, ((tu_bg, tv_bg), 0., scale_bg, (w // 2, h // 2)))
    grid, def_grid = deform_grid(h, w)
    tu_fg = int(np.random.uniform(-fg_size * bound_translate, fg_size * bound_translate))
    tv_fg = int(np.random.uniform(-fg_size * bound_translate, fg_size * bound_translate))
    rot_fg = np.random.uniform(-bound_rotate, bound_rotate)
    scale_fg = np.random.uniform(1., 1. + bound_scale)
    params_fg = (tu_fg, tv_fg), rot_fg, scale_fg, fg_center(a_dev)
    new_fg = warp_image(fg, params_fg, thin=(grid, def_grid))
    new_alpha = warp_image(alpha, params_fg, thin=(grid, def_grid))
    a = np.random.uniform(0.95, 1.05)
    b = np.random.uniform(0.7, 1.3)
    c = np.random.uniform(-0.07, 0.07)
    return change_illumination(new_fg, a, b, c), change_illumination(new_bg, a, b, c), new_alpha


AUG_PARAMS = np.dtype([("M", "<f8", (6,)), ("tu", "<i4"), ("tv", "<i4")])       # csrc/vm_affine.cu VmAugParams


def _augment_plan(stats, h, w, before_frame=None):
    """The reference's random draws for every frame of a clip, frame by frame, in its order (40 per frame,
    augmentation.py:102-135) -> (background params, foreground params, S/V tables, TPS grids).
    ``stats``: (n, 3) {count(alpha != 0), sum(rows), sum(cols)}; ``before_frame(k)`` runs ahead of frame k's
    draws (augmentation() draws its background index there, augmentation.py:158)."""
    n = len(stats)
    bt, br, bs = 0.05, 10, 0.15
    par_bg = np.zeros(n, dtype=AUG_PARAMS); par_fg = np.zeros(n, dtype=AUG_PARAMS)
    luts = np.zeros((n, 256), dtype=np.uint8)
    grids = []
    if before_frame is None and n > 1:
        return _augment_plan_batched(stats, h, w)
    for k in range(n):
        if before_frame is not None:
            before_frame(k)
        cnt, si, sj = (int(v) for v in stats[k])
        fg_size = np.sqrt(cnt)
        tu_bg = int(np.random.uniform(-w * bt, w * bt)); tv_bg = int(np.random.uniform(-h * bt, h * bt))
        scale_bg = np.random.uniform(1., 1. + bs)
        grids.append(deform_grid(h, w))
        tu_fg = int(np.random.uniform(-fg_size * bt, fg_size * bt)); tv_fg = int(np.random.uniform(-fg_size * bt, fg_size * bt))
        rot_fg = np.random.uniform(-br, br)
        scale_fg = np.random.uniform(1., 1. + bs)
        if cnt == 0:
            raise ValueError("cannot convert float NaN to integer")       # int(np.mean([])) in fg_center
        center = (int(sj / cnt), int(si / cnt))
        a = np.random.uniform(0.95, 1.05); b = np.random.uniform(0.7, 1.3); c = np.random.uniform(-0.07, 0.07)
        par_bg[k] = (_rotation_matrix((w // 2, h // 2), 0., scale_bg).reshape(6), tu_bg, tv_bg)
        par_fg[k] = (_rotation_matrix(center, rot_fg, scale_fg).reshape(6), tu_fg, tv_fg)
        luts[k] = illumination_lut(a, b, c)
    return par_bg, par_fg, luts, grids


def _augment_plan_batched(stats, h, w, n_grid=5):
    """_augment_plan for a whole clip with ONE call into the generator: the legacy ``RandomState.uniform(lo, hi)``
    is ``lo + (hi - lo) * random_sample()``, one double per draw, so the (n, 40) block of ``random_sample`` holds
    exactly the doubles the frame-by-frame calls would consume, in their order, and the same expression gives the
    same values (checked bit for bit against the sequential plan in tests/test_host_cpu.py)."""
    n = len(stats)
    bt, br, bs = 0.05, 10, 0.15
    bound = min(w, h) * 0.05
    rows = (h / (n_grid - 1)) * np.arange(n_grid)
    cols = (w / (n_grid - 1)) * np.arange(n_grid)
    grid = np.transpose([np.repeat(rows, n_grid), np.tile(cols, n_grid)])
    need = np.stack([(grid[:, 1] > 0.) & (grid[:, 1] < w), (grid[:, 0] > 0.) & (grid[:, 0] < h)], axis=1)
    ng = int(need.sum())
    per = 3 + ng + 7
    u = np.random.random_sample((n, per))
    uni = lambda lo, hi, x: lo + (hi - lo) * x
    cnt = stats[:, 0].astype(np.int64)
    if (cnt == 0).any():
        raise ValueError("cannot convert float NaN to integer")           # int(np.mean([])) in fg_center
    fg_size = np.sqrt(cnt)
    tu_bg = np.trunc(uni(-w * bt, w * bt, u[:, 0])).astype(np.int64)
    tv_bg = np.trunc(uni(-h * bt, h * bt, u[:, 1])).astype(np.int64)
    scale_bg = uni(1., 1. + bs, u[:, 2])
    draws = uni(-bound, bound, u[:, 3:3 + ng])
    o = 3 + ng
    tu_fg = np.trunc(uni(-fg_size * bt, fg_size * bt, u[:, o])).astype(np.int64)
    tv_fg = np.trunc(uni(-fg_size * bt, fg_size * bt, u[:, o + 1])).astype(np.int64)
    rot_fg = uni(-br, br, u[:, o + 2])
    scale_fg = uni(1., 1. + bs, u[:, o + 3])
    a, b, c = uni(0.95, 1.05, u[:, o + 4]), uni(0.7, 1.3, u[:, o + 5]), uni(-0.07, 0.07, u[:, o + 6])
    par_bg = np.zeros(n, dtype=AUG_PARAMS); par_fg = np.zeros(n, dtype=AUG_PARAMS)
    luts = np.zeros((n, 256), dtype=np.uint8)
    grids = []
    for k in range(n):
        flat = np.zeros(need.shape)
        flat[need] = draws[k]
        moved = grid.copy()
        moved[:, 1] += flat[:, 0]
        moved[:, 0] += flat[:, 1]
        grids.append((grid, moved))
        center = (int(int(stats[k, 2]) / int(cnt[k])), int(int(stats[k, 1]) / int(cnt[k])))
        par_bg[k] = (_rotation_matrix((w // 2, h // 2), 0., scale_bg[k]).reshape(6), tu_bg[k], tv_bg[k])
        par_fg[k] = (_rotation_matrix(center, rot_fg[k], scale_fg[k]).reshape(6), tu_fg[k], tv_fg[k])
        luts[k] = illumination_lut(a[k], b[k], c[k])
    return par_bg, par_fg, luts, grids


def _alpha_stats_clip(fg_d):
    """object_size / fg_center sums of every BGRA frame: one kernel, one host read."""
    n, h, w = fg_d.shape[:3]
    stats = torch.zeros((n, 3), dtype=torch.int64, device=fg_d.device)
    N.check(N.load().vm_alpha_stats_bgra(N.ptr(fg_d), n, h, w, N.ptr(stats), N.stream_ptr()))
    return stats.cpu().numpy()


def _augment_run(fg_d, bg_d, plan, alpha_dtype=torch.float32, status=None, pool=None):
    """Device stages of augment() for a clip whose random parameters are known: host TPS solve (system built
    from the deformed grid, tps.py:51), spline, TPS resampling, fused affine passes + illumination.
    ``alpha_dtype`` float64 carries alpha through both stages in float64 with scipy's / OpenCV's operation
    order (what the reference returns); float32 halves the intermediate traffic."""
    lib = N.load()
    par_bg, par_fg, luts, grids = plan
    n, h, w = fg_d.shape[:3]
    dev = fg_d.device
    new_fg = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    new_bg = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    wide = alpha_dtype == torch.float64
    new_alpha = torch.empty((n, h, w), dtype=torch.float64 if wide else torch.float32, device=dev)
    a64 = torch.empty((n, h + 1, w + 1), dtype=torch.float64, device=dev) if wide else None
    tplan = P.get_plan((0, 0, h, w), 2, dev)
    ctrl_h, coef_h = P.solve_grids_host(grids, pool=pool)
    flat = lambda arr: np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    # one asynchronous upload per call (pinned ring): nothing here waits for the kernels of the previous call
    ctrl, coef, par_bg_d, par_fg_d, luts_d = N.upload_many([ctrl_h, coef_h, flat(par_bg), flat(par_fg), flat(luts)], dev)
    T = torch.empty((n, tplan.nx, tplan.ny, 2), dtype=torch.float64, device=dev)
    counter = torch.zeros(64, dtype=torch.int32, device=dev)
    inter = torch.empty((n, h + 1, w + 1, 2), dtype=torch.int32, device=dev)
    Np = ctrl.shape[1]
    N.check(lib.vm_tps_coarse_packed(N.ptr(ctrl), N.ptr(coef), n, Np, tplan.nx, tplan.ny, tplan.step_x, tplan.step_y,
                                     N.ptr(T), N.ptr(counter), N.stream_ptr()))
    N.check(lib.vm_aug_tps(N.ptr(fg_d), N.ptr(T), tplan.nx, tplan.ny, N.ptr(tplan.rows), N.ptr(tplan.cols), n, h, w,
                           N.ptr(inter), N.ptr(a64), N.ptr(status), N.stream_ptr()))
    N.check(lib.vm_aug_affine(1, N.ptr(inter), N.ptr(a64), N.ptr(par_fg_d), N.ptr(luts_d), n, h, w, N.ptr(new_fg),
                              None if wide else N.ptr(new_alpha), N.ptr(new_alpha) if wide else None, N.hsv_vec(), N.stream_ptr()))
    N.check(lib.vm_aug_affine(0, N.ptr(bg_d), None, N.ptr(par_bg_d), N.ptr(luts_d), n, h, w, N.ptr(new_bg), None, None,
                              N.hsv_vec(), N.stream_ptr()))
    return new_fg, new_bg, new_alpha


def alpha_stats(fg_bgra):
    """(n, 3) int64 {count(alpha != 0), sum(rows), sum(cols)} of a BGRA clip - what object_size / fg_center need
    (augmentation.py:10-21).  Pass it to ``augment_clip(..., stats=...)`` when the same foregrounds are augmented
    repeatedly: the call then has no device-to-host synchronisation and its host work overlaps the kernels of the
    previous call."""
    fg_d, _ = N.to_device(fg_bgra)
    return _alpha_stats_clip(fg_d.contiguous())


def augment_clip(fg_bgra, bg, alpha_dtype=torch.float32, stats=None, status=None, pool=None):
    """augment() for a whole clip in a handful of launches (BASELINE config 5).

    ``fg_bgra`` (n, H, W, 4) uint8 BGRA with alpha = A/255 (what reader.read_fg_img returns, reference
    reader.py:16-17), ``bg`` (n, H, W, 3) uint8; NumPy arrays or CUDA tensors.  Equivalent to
    ``[augment(fg[k, ..., :3], bg[k], fg[k, ..., 3] / 255.) for k in range(n)]``: the global np.random
    stream is consumed in exactly that order (40 draws per frame, reference augmentation.py:102-135), the
    TPS systems are solved on the host with numpy's pinv (reference tps.py:119).  Returns
    (new_fg (n,H,W,3) uint8, new_bg (n,H,W,3) uint8, new_alpha (n,H,W) float32) of the input kind;
    ``alpha_dtype=torch.float64`` returns the alpha in float64 as the reference does (same operation order);
    ``stats=alpha_stats(fg_bgra)`` skips the per-call alpha reduction and its host synchronisation;
    ``status`` (a device int32[8] block, ``_native.new_status()``) receives the count of TPS samples that fell
    outside the source (word 3) - nothing is counted when it is None; ``pool`` (a ``pipeline.SolverPool``) spreads the
    per-frame ``np.linalg.pinv`` solves over host cores (same numpy call, bit-identical coefficients)."""
    fg_d, kind = N.to_device(fg_bgra)
    bg_d, _ = N.to_device(bg)
    assert fg_d.dtype == torch.uint8 and fg_d.dim() == 4 and fg_d.shape[3] == 4, "fg must be (n, H, W, 4) uint8 BGRA"
    n, h, w = fg_d.shape[:3]
    assert bg_d.dtype == torch.uint8 and tuple(bg_d.shape) == (n, h, w, 3), "bg must be (n, H, W, 3) uint8"
    fg_d, bg_d = fg_d.contiguous(), bg_d.contiguous()
    if n == 0:
        dev = fg_d.device
        empty = (torch.empty((0, h, w, 3), dtype=torch.uint8, device=dev), torch.empty((0, h, w, 3), dtype=torch.uint8, device=dev),
                 torch.empty((0, h, w), dtype=alpha_dtype, device=dev))
        return tuple(N.from_device(t, kind) for t in empty)
    plan = _augment_plan(_alpha_stats_clip(fg_d) if stats is None else np.asarray(stats), h, w)
    return tuple(N.from_device(t, kind) for t in _augment_run(fg_d, bg_d, plan, alpha_dtype, status, pool))


#: variants written per foreground by augmentation() (reference augmentation.py:140) and how many of them go
#: through the device stages at once
N_VARIANTS = 50
VARIANT_BATCH = 10


def augmentation(dim_dataset, voc_dataset, sig_dataset):
    """create synthetic data for video matting (DIM mattes over VOC backgrounds) - reference
    augmentation.py:138-166; disk layout, file names and the np.random draw order (one background index,
    then augment()'s 40 draws, per variant) follow the reference.  The 50 variants of a foreground are
    planned first, then run through ``augment_clip``'s device stages ten at a time while host threads read
    the backgrounds and write the PNGs.  The alpha bytes are ``(255. * alpha).astype(uint8)`` on the float64
    alpha, as in the reference.  Where alpha is locally constant that product sits within an ulp of an integer
    (255 * 0.9999999999999999 truncates to 254), so those bytes follow the last bits of the interpolation
    weights: they agree with the per-variant ``augment()`` loop to +-1 level (about 1 % of the bytes differ)."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    n = N_VARIANTS
    paths = [os.path.join(dim_dataset, 'fg', folder, f)
             for folder in ('DIM_TEST', 'DIM_TRAIN')
             for f in os.listdir(os.path.join(dim_dataset, 'fg', folder))]
    voc_list = [os.path.join(voc_dataset, f) for f in os.listdir(voc_dataset)]
    dst_fg = os.path.join(sig_dataset, 'fg', 'augmented')
    dst_bg = os.path.join(sig_dataset, 'bg', 'augmented')
    ref_lut = (255. * (np.arange(256) / 255.)).astype(np.uint8)         # (255. * alpha).astype(uint8) of alpha = A / 255.

    def load_bg(p, h, w):
        bg = cv2.imread(p)
        return cv2.resize(bg, dsize=(w, h), interpolation=cv2.INTER_LINEAR)  # augmentation.py:159-160 (host thread; the
        #                                                                       device twin is reader.resize_background)

    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as pool:
        pending = []
        for k, path in enumerate(paths):
            img = reader.read_bgra(path)
            h, w = img.shape[:2]
            name = os.path.basename(path).split('.')[0]
            print('Processing image {} ({}/{})'.format(name, k + 1, len(paths)))
            ref = img.copy()
            ref[:, :, 3] = ref_lut[img[:, :, 3]]
            pending.append(pool.submit(cv2.imwrite, os.path.join(dst_fg, '{}_fg_ref.png'.format(name)), ref))
            fg_d = torch.from_numpy(np.ascontiguousarray(img[None])).cuda()
            stats = np.repeat(_alpha_stats_clip(fg_d), n, axis=0)
            picks = []
            plan = _augment_plan(stats, h, w, before_frame=lambda i: picks.append(voc_list[np.random.randint(len(voc_list))]))
            # backgrounds are decoded / resized at most two batches ahead of the device stages (bounded host memory:
            # 50 resized backgrounds of a 3000 x 2000 foreground would be 0.9 GB)
            bg_jobs = {}

            def prefetch(upto):
                for i in range(len(bg_jobs), min(upto, n)):
                    bg_jobs[i] = pool.submit(load_bg, picks[i], h, w)

            prefetch(2 * VARIANT_BATCH)
            for lo in range(0, n, VARIANT_BATCH):
                hi = min(lo + VARIANT_BATCH, n)
                prefetch(hi + 2 * VARIANT_BATCH)
                bgs = [bg_jobs.pop(i).result() for i in range(lo, hi)]
                bg_d = torch.from_numpy(np.stack(bgs)).cuda()
                sub = (plan[0][lo:hi], plan[1][lo:hi], plan[2][lo:hi], plan[3][lo:hi])
                nfg, nbg, nal = _augment_run(fg_d.expand(hi - lo, h, w, 4).contiguous(), bg_d, sub, torch.float64)
                na8 = (255. * nal).to(torch.uint8)                      # float64 product, truncation (augmentation.py:162)
                out_fg = torch.cat((nfg, na8[..., None]), dim=3).cpu().numpy()
                out_bg = nbg.cpu().numpy()
                for i in range(lo, hi):
                    pending.append(pool.submit(cv2.imwrite, os.path.join(dst_bg, '{}_bg_ref_{:04d}.png'.format(name, i)), bgs[i - lo]))
                    pending.append(pool.submit(cv2.imwrite, os.path.join(dst_bg, '{}_bg_{:04d}.png'.format(name, i)), out_bg[i - lo]))
                    pending.append(pool.submit(cv2.imwrite, os.path.join(dst_fg, '{}_fg_{:04d}.png'.format(name, i)), out_fg[i - lo]))
            pending = [f for f in pending if not f.done() or f.result() is False]
        for f in pending:
            if f.result() is False:
                raise IOError("cv2.imwrite failed")

"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, cv2, scipy):

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs on small seeded inputs are
committed as fixtures; every parity test (oracle on CPU, CUDA path on GPU) checks against
them.  The import shim follows SURVEY.md Appendix A.0: restore the removed numpy aliases,
stub ``progressbar`` and the GUI calls of headless cv2.  Nothing here modifies or copies
reference sources.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VM_REFERENCE_DIR", "/root/reference")


def import_reference():
    import cv2
    np.float = float
    np.int = int
    pb = types.ModuleType("progressbar")
    pb.progressbar = lambda it, *a, **k: it
    sys.modules["progressbar"] = pb
    cv2.imshow = lambda *a, **k: None
    cv2.waitKey = lambda *a, **k: 27
    sys.path.insert(0, REF)
    import reader, flow, tps, augmentation  # noqa: E401
    return reader, flow, tps, augmentation


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vm_oracle as O
    import cv2
    import contextlib
    import io
    reader, flow, tps, augmentation = import_reference()
    quiet = contextlib.redirect_stdout(io.StringIO())
    out = {}

    # ---- flow warp / mask on synthetic frames (odd and even sizes) ------------------
    for tag, (h, w), seed in (("a", (61, 81), 11), ("b", (96, 128), 12), ("c", (135, 240), 13)):
        fg = O.synth_frame(seed, h, w)
        fb, ff = O.synth_flows(seed, h, w)
        if tag == "a":   # stress: large displacements, out-of-frame, but no IndexError
            rng = np.random.default_rng(5)
            fb = (fb + rng.normal(0, 6, fb.shape)).astype(np.float32)
        alpha, bgr = fg[..., 3] / 255., np.ascontiguousarray(fg[..., :3])
        wa = flow.warp_img(alpha, fb)
        wa8 = flow.warp_img(np.ascontiguousarray(fg[..., 3]), fb)
        wa32 = flow.warp_img(alpha.astype(np.float32), fb)
        wb = flow.warp_bgr(bgr, fb)
        with quiet:
            ca = flow.correct_alpha(fb, ff, wa.copy())
        grids = O.synth_grids(seed, h, w, 5)
        ident = ((0, 0), 0., 1., (w // 2, h // 2))
        b2 = augmentation.warp_image(wb, ident, thin=grids)
        a2 = augmentation.warp_image(ca, ident, thin=grids)
        bg = O.synth_background(seed, h, w)
        cmp4 = reader.create_composite_image(b2, bg, a2)
        out.update({f"flow_{tag}_fg": fg, f"flow_{tag}_fb": fb, f"flow_{tag}_ff": ff,
                    f"flow_{tag}_warp_alpha": wa, f"flow_{tag}_warp_alpha_u8": wa8,
                    f"flow_{tag}_warp_alpha_f32": wa32, f"flow_{tag}_warp_bgr": wb,
                    f"flow_{tag}_corrected": ca,
                    f"c4_{tag}_grid": grids[0], f"c4_{tag}_defgrid": grids[1], f"c4_{tag}_bg": bg,
                    f"c4_{tag}_bgr": b2, f"c4_{tag}_alpha": a2, f"c4_{tag}_cmp": cmp4})

    # ---- remap special values --------------------------------------------------------
    h, w = 32, 48
    fg = O.synth_frame(21, h, w)
    fb = np.zeros((h, w, 2), np.float32)
    rng = np.random.default_rng(3)
    fb[...] = rng.normal(0, 40, fb.shape)
    fb[3, 5] = (np.nan, 0); fb[4, 6] = (0, np.inf); fb[5, 7] = (-np.inf, 1); fb[6, 8] = (3e9, 0)
    fb[7, 9] = (1e5, -1e5); fb[8, 10] = (0.015625, 0.046875); fb[9, 11] = (0.484375, 0.515625)
    out.update({"special_fg": fg, "special_fb": fb,
                "special_warp_bgr": flow.warp_bgr(np.ascontiguousarray(fg[..., :3]), fb),
                "special_warp_alpha": flow.warp_img(fg[..., 3] / 255., fb)})

    # ---- TPS transform + warp_images -------------------------------------------------
    for tag, (h, w), n, seed in (("a", (61, 81), 5, 31), ("b", (64, 96), 4, 32), ("c", (120, 90), 5, 33)):
        grid, dgrid = O.synth_grids(seed, h, w, n)
        tr = tps._make_inverse_warp(grid, dgrid, (0, 0, h, w), 2)
        fg = O.synth_frame(seed, h, w)
        alpha = fg[..., 3] / 255.
        res = tps.warp_images(grid, dgrid, [fg[..., 0], fg[..., 1], fg[..., 2], alpha], (0, 0, h, w),
                              interpolation_order=1, approximate_grid=2)
        out.update({f"tps_{tag}_grid": grid, f"tps_{tag}_defgrid": dgrid, f"tps_{tag}_fg": fg,
                    f"tps_{tag}_t0": tr[0], f"tps_{tag}_t1": tr[1],
                    f"tps_{tag}_b": res[0], f"tps_{tag}_g": res[1], f"tps_{tag}_r": res[2],
                    f"tps_{tag}_alpha": res[3]})

    # ---- deform_grid RNG contract ---------------------------------------------------
    np.random.seed(1234)
    g, d = augmentation.deform_grid(108, 192)
    np.random.seed(1234)
    g4, d4 = augmentation.deform_grid(64, 64, n=4)
    out.update({"grid_108x192_n5": g, "defgrid_108x192_n5": d, "grid_64x64_n4": g4, "defgrid_64x64_n4": d4})

    # ---- warp_image (affine only and TPS + affine) -----------------------------------
    h, w = 75, 110
    fg = O.synth_frame(41, h, w)
    bgr = np.ascontiguousarray(fg[..., :3]); alpha = fg[..., 3] / 255.
    grids = O.synth_grids(41, h, w, 5)
    p1 = ((3, -2), 7.5, 1.1, (52, 31))
    p2 = ((-4, 5), 0., 1.07, (w // 2, h // 2))
    out.update({"wi_fg": fg, "wi_grid": grids[0], "wi_defgrid": grids[1],
                "wi_p1_bgr": augmentation.warp_image(bgr, p1), "wi_p1_alpha": augmentation.warp_image(alpha, p1),
                "wi_p2_bgr": augmentation.warp_image(bgr, p2),
                "wi_p1_thin_bgr": augmentation.warp_image(bgr, p1, thin=grids),
                "wi_p1_thin_alpha": augmentation.warp_image(alpha, p1, thin=grids)})

    # ---- change_illumination, object_size, fg_center, augment ------------------------
    ci = augmentation.change_illumination(bgr, 1.03, 0.8, -0.02)
    out.update({"ci_bgr": bgr, "ci_out": ci, "ci_hsv": cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV),
                "stats_size": np.float64(augmentation.object_size(alpha)),
                "stats_center": np.array(augmentation.fg_center(alpha))})
    bg = O.synth_background(41, h, w)
    np.random.seed(77)
    nfg, nbg, nal = augmentation.augment(bgr, bg, alpha)
    out.update({"aug_bg": bg, "aug_fg_out": nfg, "aug_bg_out": nbg, "aug_alpha_out": nal})

    # ---- composite + uint16 read quirk ----------------------------------------------
    cmp_ = reader.create_composite_image(bgr, bg, alpha)
    v16 = np.array([0, 1, 254, 255, 256, 510, 511, 512, 767, 768, 32767, 32768, 65279, 65280,
                    65534, 65535], dtype=np.uint16)
    q = (((v16 + 1) / 256.) - 1).astype(np.uint8)
    out.update({"cmp_out": cmp_, "u16_in": v16, "u16_out": q})

    # ---- reference-owned golden vectors: cmp1.png / cmp2.png on a crop ---------------
    # (cropped so that the fixture stays small; the full-frame identity is asserted in
    #  make_golden itself)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        for k, name in ((1, "in0062.png"), (2, "in0063.png")):
            fgi, bgi, cmpi, ali = reader.load_test_image(name, "sea.jpg")
            gold = cv2.imread(os.path.join("test_data", f"cmp{k}.png"))
            assert np.array_equal(np.rint(cmpi).astype(np.uint8), gold), "cmp golden identity failed"
            raw = cv2.imread(os.path.join("test_data", name), cv2.IMREAD_UNCHANGED)
            sl = (slice(150, 278), slice(520, 712))
            out.update({f"cmp{k}_raw16": raw[sl], f"cmp{k}_bg": bgi[sl], f"cmp{k}_gold": gold[sl]})
    finally:
        os.chdir(cwd)

    # ---- .flo -----------------------------------------------------------------------
    fb, _ = O.synth_flows(5, 17, 23)
    p = os.path.join(HERE, "tiny.flo")
    O.write_flo(p, fb)
    assert np.array_equal(reader.read_flow(p), fb)
    out["flo_tiny"] = fb

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    sz = os.path.getsize(os.path.join(HERE, "golden.npz"))
    print(f"wrote golden.npz: {len(out)} arrays, {sz / 1e6:.2f} MB")


if __name__ == "__main__":
    main()

"""CPU tests: the oracle (oracle/vm_oracle.py) against the golden vectors produced by the
unmodified reference, and against cv2 / scipy where those are importable.

Bars: uint8 outputs, masks, .flo parse bit-exact; float alpha / composites
|got-ref| <= 1e-5*|ref| + 1e-6; change_illumination exact (round 2: row f2).
"""
import os
import sys

import numpy as np
import pytest

import vm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(a, b, rtol=1e-5, atol=1e-6):
    return np.allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_flow_warp_and_mask(golden, tag):
    fg, fb, ff = golden[f"flow_{tag}_fg"], golden[f"flow_{tag}_fb"], golden[f"flow_{tag}_ff"]
    alpha, bgr = fg[..., 3] / 255., np.ascontiguousarray(fg[..., :3])
    assert np.array_equal(O.warp_bgr(bgr, fb), golden[f"flow_{tag}_warp_bgr"])
    assert np.array_equal(O.warp_img(np.ascontiguousarray(fg[..., 3]), fb), golden[f"flow_{tag}_warp_alpha_u8"])
    wa = O.warp_img(alpha, fb)
    assert np.array_equal(wa, golden[f"flow_{tag}_warp_alpha"])          # float64 path is bit-equal
    assert np.array_equal(O.warp_img(alpha.astype(np.float32), fb), golden[f"flow_{tag}_warp_alpha_f32"])
    ca = O.correct_alpha(fb, ff, wa.copy())
    assert np.array_equal(ca, golden[f"flow_{tag}_corrected"])
    assert (ca != wa).any(), "fixture must exercise the occlusion mask"


def test_correct_alpha_loop_matches_vectorised(golden):
    fb, ff = golden["flow_a_fb"][:24, :30], golden["flow_a_ff"][:24, :30]
    a = np.random.default_rng(0).random((24, 30))
    assert np.array_equal(O.correct_alpha_loop(fb, ff, a.copy()), O.correct_alpha(fb, ff, a.copy()))


def test_remap_special_values(golden):
    fg, fb = golden["special_fg"], golden["special_fb"]
    assert np.array_equal(O.warp_bgr(np.ascontiguousarray(fg[..., :3]), fb), golden["special_warp_bgr"])
    assert np.array_equal(O.warp_img(fg[..., 3] / 255., fb), golden["special_warp_alpha"])


@pytest.mark.parametrize("tag,n", [("a", 5), ("b", 4), ("c", 5)])
def test_tps_transform_and_warp(golden, tag, n):
    grid, dgrid, fg = golden[f"tps_{tag}_grid"], golden[f"tps_{tag}_defgrid"], golden[f"tps_{tag}_fg"]
    h, w = fg.shape[:2]
    (t0, t1), _, _ = O.tps_inverse_transform(grid, dgrid, h, w)
    assert t0.shape == (h + 1, w + 1)
    assert np.array_equal(t0, golden[f"tps_{tag}_t0"]) and np.array_equal(t1, golden[f"tps_{tag}_t1"])
    res = O.tps_warp_images(grid, dgrid, [fg[..., 0], fg[..., 1], fg[..., 2], fg[..., 3] / 255.], (0, 0, h, w))
    for got, key in zip(res, "bgr"):
        assert np.array_equal(got, golden[f"tps_{tag}_{key}"])
    assert np.array_equal(res[3], golden[f"tps_{tag}_alpha"])


def test_deform_grid_rng_contract(golden):
    np.random.seed(1234)
    g, d = O.deform_grid(108, 192)
    assert np.array_equal(g, golden["grid_108x192_n5"]) and np.array_equal(d, golden["defgrid_108x192_n5"])
    np.random.seed(1234)
    g, d = O.deform_grid(64, 64, n=4)
    assert np.array_equal(g, golden["grid_64x64_n4"]) and np.array_equal(d, golden["defgrid_64x64_n4"])


def test_warp_image_chain(golden):
    fg = golden["wi_fg"]
    h, w = fg.shape[:2]
    bgr, alpha = np.ascontiguousarray(fg[..., :3]), fg[..., 3] / 255.
    grids = (golden["wi_grid"], golden["wi_defgrid"])
    p1 = ((3, -2), 7.5, 1.1, (52, 31))
    p2 = ((-4, 5), 0., 1.07, (w // 2, h // 2))
    assert np.array_equal(O.warp_image(bgr, p1), golden["wi_p1_bgr"])
    assert np.array_equal(O.warp_image(alpha, p1), golden["wi_p1_alpha"])
    assert np.array_equal(O.warp_image(bgr, p2), golden["wi_p2_bgr"])
    assert np.array_equal(O.warp_image(bgr, p1, thin=grids), golden["wi_p1_thin_bgr"])
    assert np.array_equal(O.warp_image(alpha, p1, thin=grids), golden["wi_p1_thin_alpha"])


def test_illumination_and_stats(golden):
    bgr = golden["ci_bgr"]
    assert np.array_equal(O.bgr2hsv_u8(bgr), golden["ci_hsv"])
    # HSV2BGR is bit-exact since round 2 (truncating SIMD body, rounding row tail; fixtures made with AVX2: 32 px per step)
    assert np.array_equal(O.change_illumination(bgr, 1.03, 0.8, -0.02, vec=32), golden["ci_out"])
    alpha = golden["wi_fg"][..., 3] / 255.
    assert O.object_size(alpha) == float(golden["stats_size"])
    assert tuple(O.fg_center(alpha)) == tuple(golden["stats_center"])


def test_augment_end_to_end(golden):
    fg = golden["wi_fg"]
    bgr, alpha = np.ascontiguousarray(fg[..., :3]), fg[..., 3] / 255.
    np.random.seed(77)
    nfg, nbg, nal = O.augment(bgr, golden["aug_bg"], alpha)
    assert np.array_equal(nal, golden["aug_alpha_out"])
    assert np.array_equal(nfg, golden["aug_fg_out"]) and np.array_equal(nbg, golden["aug_bg_out"])


def test_composite_and_uint16_quirk(golden):
    fg = golden["wi_fg"]
    bgr, alpha = np.ascontiguousarray(fg[..., :3]), fg[..., 3] / 255.
    assert np.array_equal(O.create_composite_image(bgr, golden["aug_bg"], alpha), golden["cmp_out"])
    assert np.array_equal(O.fg_from_uint16(golden["u16_in"]), golden["u16_out"])


@pytest.mark.parametrize("k", [1, 2])
def test_reference_owned_golden_cmp_png(golden, k):
    """cmp1.png / cmp2.png (reference test_data) == rint(composite(read_fg_img(in006x), bg))."""
    img8 = O.fg_from_uint16(golden[f"cmp{k}_raw16"])
    alpha, bgr = O.split_fg(img8)
    cmp_ = O.create_composite_image(bgr, golden[f"cmp{k}_bg"], alpha)
    assert np.array_equal(np.rint(cmp_).astype(np.uint8), golden[f"cmp{k}_gold"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_c4_pipeline(golden, tag):
    fg, fb, ff = golden[f"flow_{tag}_fg"], golden[f"flow_{tag}_fb"], golden[f"flow_{tag}_ff"]
    grids = (golden[f"c4_{tag}_grid"], golden[f"c4_{tag}_defgrid"])
    cmp_, a2 = O.pipeline_c4(fg, fb, ff, grids, golden[f"c4_{tag}_bg"])
    assert np.array_equal(a2, golden[f"c4_{tag}_alpha"])
    assert np.array_equal(cmp_, golden[f"c4_{tag}_cmp"])


def test_flo_roundtrip(tmp_path, golden):
    with open(os.path.join(ROOT, "tests", "golden", "tiny.flo"), "rb") as f:
        buf = f.read()
    flow, ok = O.parse_flo(buf)
    assert ok and np.array_equal(flow, golden["flo_tiny"])
    bad = bytearray(buf); bad[0] ^= 0xFF
    flow2, ok2 = O.parse_flo(bytes(bad))
    assert not ok2 and np.array_equal(flow2, flow)                     # bad magic: data still returned
    with pytest.raises(ValueError):
        O.parse_flo(buf[:-8])                                          # truncated payload


# ---- differential tests against the third-party kernels, where importable -------------

def test_remap_vs_cv2_random():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (h, w) in ((37, 53), (64, 64)):
        src8 = rng.integers(0, 256, (h, w), dtype=np.uint8)
        src64 = rng.random((h, w))
        for sigma in (2, 50, 1e5):
            flow = rng.normal(0, sigma, (h, w, 2)).astype(np.float32)
            mx, my = O.flow_map(flow)
            m = np.stack([mx, my], -1)
            assert np.array_equal(O.warp_img(src8, flow), cv2.remap(src8, m, None, cv2.INTER_LINEAR))
            assert np.array_equal(O.warp_img(src64, flow), cv2.remap(src64, m, None, cv2.INTER_LINEAR))


def test_warp_affine_vs_cv2_random():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    src8 = rng.integers(0, 256, (45, 70, 3), dtype=np.uint8)
    src64 = rng.random((46, 71))
    for _ in range(6):
        M = cv2.getRotationMatrix2D((rng.uniform(0, 70), rng.uniform(0, 45)), rng.uniform(-10, 10),
                                    rng.uniform(1, 1.15))
        assert np.array_equal(O.rotation_matrix_2d((0, 0), 0, 1), cv2.getRotationMatrix2D((0, 0), 0, 1))
        assert np.array_equal(O.warp_affine(src8, M, (70, 45)), cv2.warpAffine(src8, M, (70, 45)))
        assert np.array_equal(O.warp_affine(src64, M, (70, 45)), cv2.warpAffine(src64, M, (70, 45)))


def test_map_coordinates_vs_scipy_random():
    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(2)
    img8 = rng.integers(0, 256, (33, 47), dtype=np.uint8)
    img64 = rng.random((33, 47))
    t0 = rng.uniform(-2, 35, (40, 50)); t1 = rng.uniform(-2, 49, (40, 50))
    t0[0, 0], t1[0, 0] = 0.0, 46.0
    t0[0, 1], t1[0, 1] = 32.0, 0.5
    t0[0, 2], t1[0, 2] = np.nextafter(32.0, 64), 3.0
    t0[0, 3], t1[0, 3] = 2.5, 7.5
    assert np.array_equal(O.map_coordinates_linear(img8, t0, t1), ndi.map_coordinates(img8, [t0, t1], order=1))
    assert np.array_equal(O.map_coordinates_linear(img64, t0, t1), ndi.map_coordinates(img64, [t0, t1], order=1))


def test_bgr2hsv_vs_cv2_exhaustive_slice():
    cv2 = pytest.importorskip("cv2")
    v = np.arange(256, dtype=np.uint8)
    for b in (0, 1, 77, 128, 254, 255):
        g, r = np.meshgrid(v, v)
        bgr = np.stack([np.full_like(g, b), g, r], -1)
        assert np.array_equal(O.bgr2hsv_u8(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV))


# ---- BASELINE config 1 on the reference's own test images (tests/golden/make_c1_golden.py) ----------------

@pytest.fixture(scope="module")
def c1():
    with np.load(os.path.join(ROOT, "tests", "golden", "c1_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def test_config1_window_oracle(c1):
    """in0063 warped onto in0062 with DIS stand-in flows, consistency mask, composite onto sea.jpg: the oracle
    against the unmodified reference on a window of the real frames."""
    fg = c1["fg63_bgra"]
    alpha, bgr = fg[..., 3] / 255., np.ascontiguousarray(fg[..., :3])
    wa = O.warp_img(alpha, c1["backward"])
    assert np.array_equal(wa, c1["warp_alpha"])
    assert np.array_equal(O.warp_bgr(bgr, c1["backward"]), c1["warp_bgr"])
    ca = O.correct_alpha(c1["backward"], c1["forward"], wa.copy())
    assert np.array_equal(ca, c1["corrected"]) and int((ca != wa).sum()) > 0
    assert np.array_equal(O.create_composite_image(c1["warp_bgr"], c1["bg"], ca), c1["composite"])
    assert c1["mae"][1] < 0.2 * c1["mae"][0]          # recorded at native resolution by the generator


# ---- BASELINE config 1 at NATIVE resolution, 500 x 1200 (tests/golden/make_c1_full_golden.py) -----------------

def load_c1_full():
    """Fixture of make_c1_full_golden.py -> dict with the decoded inputs added (decoded with cv2 only:
    the product / oracle readers are what the tests check)."""
    import hashlib
    with np.load(os.path.join(ROOT, "tests", "golden", "c1_full_golden.npz")) as z:
        c = {k: z[k] for k in z.files}
    c["backward"] = c["backward_q16"].astype(np.float32) / 16.0
    c["forward"] = c["forward_q16"].astype(np.float32) / 16.0
    c["sha"] = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    return c


def test_config1_native_oracle():
    """The oracle against the unmodified reference on the whole 500 x 1200 frames: bit-exact outputs by
    digest, the composite on the stored sub-grid."""
    import cv2
    c = load_c1_full()
    sha = c["sha"]
    img16 = cv2.imdecode(c["in0063_png"], cv2.IMREAD_UNCHANGED)
    assert img16.dtype == np.uint16 and img16.shape == (500, 1200, 4)
    alpha, bgr = O.split_fg(O.fg_from_uint16(img16))
    bgr = np.ascontiguousarray(bgr)
    assert np.array_equal(sha(bgr), c["sha_fg63"]) and np.array_equal(sha(alpha), c["sha_alpha63"])
    jpg = cv2.imdecode(c["sea_jpg"], cv2.IMREAD_COLOR)
    bg = cv2.resize(jpg, dsize=(1200, 500), interpolation=cv2.INTER_LINEAR)          # host call of the reference (reader.py:41)
    assert np.array_equal(sha(bg), c["sha_bg"])
    assert np.array_equal(O.resize_linear_u8(jpg, (1200, 500)), bg)                   # row f2: the uint8 resize restated
    wa = O.warp_img(alpha, c["backward"])
    assert np.array_equal(sha(wa), c["sha_warp_alpha"])
    wb = O.warp_bgr(bgr, c["backward"])
    assert np.array_equal(sha(wb), c["sha_warp_bgr"]) and int(wb.astype(np.int64).sum()) == int(c["sum_warp_bgr"][0])
    ca = O.correct_alpha(c["backward"], c["forward"], wa.copy())
    assert int((ca != wa).sum()) == int(c["n_masked"][0])
    assert np.array_equal(sha(ca), c["sha_corrected"])
    cmp_ = O.create_composite_image(wb, bg, ca)
    assert np.array_equal(cmp_[::3, ::3].astype(np.float32), c["composite_sub"])
    assert cmp_.sum() == c["sum_composite"][0]


def test_tps_order0_oracle_vs_scipy_and_reference():
    """interpolation_order=0 (tps.py:22,34): the oracle's nearest-neighbour rule against scipy itself on
    knife-edge coordinates, and against the unmodified reference's tps.warp_images when it is importable."""
    from scipy import ndimage
    rng = np.random.default_rng(3)
    img8 = rng.integers(0, 256, (23, 31), dtype=np.uint8)
    img64 = rng.random((23, 31))
    t0 = rng.uniform(-2, 25, (40, 50))
    t1 = rng.uniform(-2, 33, (40, 50))
    t0[:6, :8] = np.array([-0.5, -1e-9, 0.0, 0.5, 1.5, 2.5, 22.0, 22.0000001])[None, :]
    t1[:6, :8] = np.array([0.5, 1.5, 2.5, 29.5, 30.0, 30.0000001])[:, None]
    for img in (img8, img64):
        ref = ndimage.map_coordinates(img, [t0, t1], order=0)
        assert np.array_equal(O.map_coordinates_nearest(img, t0, t1), ref)
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refshim
    if refshim.reference_dir() is None:
        pytest.skip("reference modules not available")
    tps = refshim.load(("tps",))["tps"]
    h, w = 61, 83
    grid, dgrid = O.synth_grids(9, h, w, 4)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    for order in (0, 1):
        ref = tps.warp_images(grid, dgrid, [img, img / 255.], (0, 0, h, w), interpolation_order=order)
        got = O.tps_warp_images(grid, dgrid, [img, img / 255.], (0, 0, h, w), interpolation_order=order)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])


# ---- row f2: cv2's uint8 HSV2BGR and uint8 INTER_LINEAR resize, restated bit for bit -------------------------------

def test_hsv2bgr_exhaustive_vs_cv2():
    """All 180 x 256 x 256 HSV inputs through cv2's SIMD body (rows without a tail) and through its scalar tail
    (one-pixel rows) against the oracle's two roundings; then the row rule on images with a tail."""
    import cv2
    vec = O.probe_hsv_vec()
    assert vec in (16, 32, 64)
    H, S, V = np.meshgrid(np.arange(180, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    hsv = np.stack([H, S, V], -1).reshape(2880, 4096, 3)                  # 4096 % 64 == 0: no tail
    assert np.array_equal(cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR), O.hsv2bgr_u8(hsv, vec))
    col = hsv.reshape(-1, 1, 3)[::7]                                         # one-pixel rows: tail only
    assert np.array_equal(cv2.cvtColor(np.ascontiguousarray(col), cv2.COLOR_HSV2BGR), O.hsv2bgr_u8(col, vec))
    rng = np.random.default_rng(4)
    for w in (1, 31, 33, 70, 127, 200):
        img = np.stack([rng.integers(0, 180, (9, w)), rng.integers(0, 256, (9, w)), rng.integers(0, 256, (9, w))], -1).astype(np.uint8)
        assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_HSV2BGR), O.hsv2bgr_u8(img, vec)), w
        bgr = rng.integers(0, 256, (9, w, 3), dtype=np.uint8)
        hs = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
        assert np.array_equal(hs, O.bgr2hsv_u8(bgr))


def test_change_illumination_exact_vs_reference():
    """augmentation.change_illumination of the unmodified reference == the oracle, bit for bit, incl. widths with a tail"""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refshim
    if refshim.reference_dir() is None:
        pytest.skip("reference modules not available")
    aug = refshim.load(("augmentation",))["augmentation"]
    vec = O.probe_hsv_vec()
    rng = np.random.default_rng(8)
    for (h, w), (a, b, c) in (((37, 53), (1.03, 0.8, -0.02)), ((40, 64), (0.95, 1.3, 0.07)), ((21, 100), (1.05, 0.7, -0.07))):
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(aug.change_illumination(bgr, a, b, c), O.change_illumination(bgr, a, b, c, vec))


def test_resize_u8_vs_cv2():
    import cv2
    rng = np.random.default_rng(0)
    shapes = [(300, 400, 500, 1200), (333, 517, 500, 1200), (375, 500, 1080, 1920), (1080, 1920, 512, 512), (100, 100, 37, 53),
              (64, 64, 128, 128), (128, 128, 64, 64), (128, 130, 64, 65), (5, 7, 50, 120), (1, 1, 4, 4), (2, 3, 1, 1), (17, 31, 16, 30),
              (200, 200, 199, 201), (50, 120, 50, 120)]
    for sh, sw, dh, dw in shapes:
        for cn in (1, 3):
            src = rng.integers(0, 256, (sh, sw, cn) if cn > 1 else (sh, sw), dtype=np.uint8)
            assert np.array_equal(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), O.resize_linear_u8(src, (dw, dh))), (sh, sw, dh, dw, cn)

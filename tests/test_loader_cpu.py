"""CPU tests of the batch-loader row: the oracle restatement against the golden
vectors produced by the unmodified reference ``loader.py`` and against cv2.resize, plus the
host-side crop planning of the product against the oracle's array slicing."""
import os
import sys

import numpy as np
import pytest

import vm_loader_oracle as LO
import vm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_loader_golden as MG  # noqa: E402

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def lgold():
    with np.load(os.path.join(ROOT, "tests", "golden", "loader_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def sample_files(lgold, tag, d):
    return MG.write_inputs(str(d), {k: lgold[f"{tag}_file_{k}"] for k in ("fg", "prev", "bg", "flo", "hw")})


def decode(paths):
    fg = cv2.imread(paths["fg"], cv2.IMREAD_UNCHANGED)
    prev = cv2.imread(paths["prev"], cv2.IMREAD_UNCHANGED)
    bg = cv2.imread(paths["bg"])
    tri = cv2.imread(paths["tri"], 0)
    flo, ok = O.parse_flo(open(paths["flo"], "rb").read())
    assert ok
    return fg, prev, bg, tri, flo


def close(got, ref, what):
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    err = np.abs(got - ref)
    tol = 1e-5 * np.abs(ref) + 1e-9
    assert np.all(err <= tol), f"{what}: max err {err.max():.3e}"


@pytest.mark.parametrize("shape,dsize", [
    ((480, 480, 3), (320, 320)), ((640, 640, 3), (320, 320)), ((320, 320, 3), (320, 320)),
    ((561, 998, 3), (320, 320)), ((77, 131), (320, 320)), ((640, 640), (320, 320)),
    ((333, 640, 3), (320, 320)), ((640, 1280, 3), (320, 640)), ((200, 100, 3), (96, 128)),
    ((641, 640, 1), (320, 320)), ((3, 2, 3), (16, 16)), ((2, 5), (7, 9)), ((480, 640, 3), (128, 96))])
def test_resize_model_matches_cv2(shape, dsize):
    rng = np.random.default_rng(sum(shape) + dsize[0])
    src = rng.uniform(0, 255, shape)
    ref = cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR)
    got = LO.resize_linear_f64(src, dsize)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-9


def test_get_padded_img_matches_reference_draw_order():
    # the reference draws rows before columns, one randint per axis (loader.py:15-34)
    rng = np.random.RandomState(5)
    img = np.arange(6 * 9 * 2, dtype=np.float64).reshape(6, 9, 2)
    out = LO.get_padded_img(img, 8, 4, rng)
    rng2 = np.random.RandomState(5)
    oi = rng2.randint(0, 8 - 6 + 1)
    ij = rng2.randint(0, 9 - 4 + 1)
    exp = np.zeros((8, 9, 2))
    exp[oi:oi + 6, 0:4] = img[:, ij:ij + 4]
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("tag", [c[0] for c in MG.CASES])
def test_oracle_matches_reference_golden(lgold, tag, tmp_path):
    lat = int(lgold["lattice"])
    fg, prev, bg, tri, flo = decode(sample_files(lgold, tag, tmp_path))
    w_in, h_in, seed = (int(v) for v in lgold[f"{tag}_meta"])
    input_size = (w_in, h_in)
    res = LO.video_sample(fg, bg, prev, flo, input_size, np.random.RandomState(seed))
    for name, a in zip(("cmp", "bg", "label", "warped", "fg"), res):
        close(a[::lat, ::lat], lgold[f"{tag}_video_{name}"], f"{tag} video {name}")
    shapes = [a.shape + (0,) * (3 - a.ndim) for a in res]
    assert np.array_equal(np.array(shapes), lgold[f"{tag}_video_shapes"])
    res = LO.simple_sample(fg, bg, input_size, np.random.RandomState(seed + 100))
    for name, a in zip(("cmp", "bg", "label", "fg"), res):
        close(a[::lat, ::lat], lgold[f"{tag}_simple_{name}"], f"{tag} simple {name}")
    res = LO.trimap_sample(fg, tri, bg, input_size, np.random.RandomState(seed + 200))
    for name, a in zip(("inp", "label", "fg"), res):
        close(a[::lat, ::lat], lgold[f"{tag}_trimap_{name}"], f"{tag} trimap {name}")


def test_psnr(lgold):
    assert abs(LO.psnr(lgold["psnr_a"], lgold["psnr_b"]) - float(lgold["psnr_3"])) < 1e-9
    assert abs(LO.psnr(lgold["psnr_a"][:, :, 0], lgold["psnr_b"][:, :, 0]) - float(lgold["psnr_1"])) < 1e-9


# ---- host-side planning of the product (no GPU needed) ------------------------------------------

def view_window(img, v):
    """NumPy emulation of what csrc/vm_loader.cu reads through a view descriptor."""
    out = np.zeros((int(v["win_h"]), int(v["win_w"])) + img.shape[2:], dtype=np.float64)
    for r in range(out.shape[0]):
        cr = int(v["wi"]) + r
        if not (int(v["vi0"]) <= cr < int(v["vi1"])):
            continue
        c0, c1 = max(int(v["vj0"]), int(v["wj"])), min(int(v["vj1"]), int(v["wj"]) + out.shape[1])
        if c1 > c0:
            src_r = cr - int(v["vi0"]) + int(v["si"])
            src_c = c0 - int(v["vj0"]) + int(v["sj"])
            out[r, c0 - int(v["wj"]):c1 - int(v["wj"])] = img[src_r, src_c:src_c + (c1 - c0)]
    return out


@pytest.mark.parametrize("fg_hw,bg_hw,input_size,seed", [
    ((400, 520), (300, 410), (320, 320), 0), ((250, 700), (640, 640), (320, 320), 1),
    ((700, 250), (333, 222), (160, 160), 2), ((100, 90), (50, 60), (64, 48), 3),
    ((660, 650), (320, 320), (320, 320), 4), ((640, 640), (640, 640), (320, 320), 5),
    ((480, 900), (100, 100), (320, 320), 6), ((320, 320), (320, 320), (320, 320), 7)])
def test_plan_matches_oracle_crops(vm, fg_hw, bg_hw, input_size, seed):
    rng = np.random.default_rng(seed)
    fg = rng.integers(0, 256, size=fg_hw + (4,), dtype=np.uint8)
    bg = rng.integers(0, 256, size=bg_hw + (3,), dtype=np.uint8)
    for rep in range(4):
        np.random.seed(1000 * seed + rep)
        fgv, bgv = vm.loader._plan_sample(fg_hw[0], fg_hw[1], bg_hw[0], bg_hw[1], input_size)
        state_after = np.random.get_state()[1][:8].copy(), np.random.get_state()[2]
        ref = LO.simple_sample(fg, bg, input_size, np.random.RandomState(1000 * seed + rep))
        # same number of draws consumed as the reference sequence
        r2 = np.random.RandomState(1000 * seed + rep)
        LO.simple_sample(fg, bg, input_size, r2)
        assert np.array_equal(state_after[0], r2.get_state()[1][:8]) and state_after[1] == r2.get_state()[2]
        got_fg = LO.resize_linear_f64(view_window(fg[..., :3], fgv), input_size)
        got_bg = LO.resize_linear_f64(view_window(bg, bgv), input_size) - LO.VGG_MEAN
        assert np.array_equal(got_fg, ref[3])
        assert np.array_equal(got_bg, ref[1])
        assert int(fgv["mode"]) == int(fgv["win_h"] == 2 * input_size[1] and fgv["win_w"] == 2 * input_size[0])
        # staging copies only the rectangle a view can touch; the re-based view reads the same window
        for img, v in ((fg, fgv), (bg, bgv)):
            (r0, r1, c0, c1), v2 = vm.loader._touched(v, img.shape[0], img.shape[1])
            assert 0 <= r0 < r1 <= img.shape[0] and 0 <= c0 < c1 <= img.shape[1]
            assert np.array_equal(view_window(img[r0:r1, c0:c1], v2), view_window(img, v))


def test_product_get_padded_img_matches_oracle(vm):
    img = np.random.default_rng(0).uniform(0, 1, (37, 91, 5))
    for k, (ch, cw) in enumerate([(64, 64), (20, 120), (64, 50), (37, 91), (10, 10)]):
        np.random.seed(k)
        got = vm.loader.get_padded_img(img, ch, cw)
        assert np.array_equal(got, LO.get_padded_img(img, ch, cw, np.random.RandomState(k)))


def test_descriptor_layout_matches_header(vm):
    import re
    hdr = open(os.path.join(ROOT, "include", "vm_b200.h")).read()
    body = re.search(r"typedef struct \{([^}]*)\} vm_loader_view;", hdr).group(1)
    names = re.findall(r"\b(\w+)\s*[,;]", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == list(vm.loader.VIEW_DTYPE.names)
    body = re.search(r"typedef struct \{([^}]*)\} vm_loader_sample;", hdr).group(1)
    names = re.findall(r"\*?(\w+)\s*[,;]", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == list(vm.loader.SAMPLE_DTYPE.names)


def test_trimap_oracle_matches_reference(lgold):
    for tag in ("t0", "t1"):
        m = lgold[f"{tag}_matte_u8"] / 255.
        assert np.array_equal(LO.trimap_from_matte_loop(m), lgold[f"{tag}_trimap"])
        assert np.array_equal(LO.trimap_from_matte(m), lgold[f"{tag}_trimap"])
    rng = np.random.default_rng(0)
    m = rng.choice([0., 1., 0.5], size=(23, 31), p=[0.45, 0.45, 0.1])
    assert np.array_equal(LO.trimap_from_matte(m), LO.trimap_from_matte_loop(m))


def test_u16_quirk_closed_form_is_the_reference_expression():
    # csrc/vm_loader.cu: vm_u16_quirk; reference reader.py:13-15 evaluated literally in numpy
    v = np.arange(65536, dtype=np.uint32)
    q = (v + 1) & 0xFFFF
    closed = np.where(q == 0, 255, np.where(q < 256, 0, (q >> 8) - 1)).astype(np.uint8)
    assert np.array_equal(closed, O.fg_from_uint16(v.astype(np.uint16)))
    v16 = v.astype(np.uint16)
    with np.errstate(all="ignore"):
        literal = (((v16 + np.uint16(1)) / 256.) - 1)
    assert np.array_equal(closed, (np.trunc(literal).astype(np.int64) & 0xFF).astype(np.uint8))


def test_plan_random_geometries_stay_in_bounds_and_match_oracle(vm):
    """Descriptor-driven reads must never leave the staged rectangles: 150 random image / background / output
    sizes (padding on either axis, all crop types, the 2x area branch) through the host planner, the window
    emulation and the oracle."""
    rng = np.random.default_rng(2025)
    for trial in range(150):
        fh, fw = int(rng.integers(2, 700)), int(rng.integers(2, 700))
        bh, bw = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        size = (int(rng.choice([32, 48, 160, 240, 320])), int(rng.choice([32, 48, 160, 240, 320])))
        fg = rng.integers(0, 256, size=(fh, fw, 4), dtype=np.uint8)
        bg = rng.integers(0, 256, size=(bh, bw, 3), dtype=np.uint8)
        np.random.seed(trial)
        fgv, bgv = vm.loader._plan_sample(fh, fw, bh, bw, size)
        for img, v in ((fg, fgv), (bg, bgv)):
            (r0, r1, c0, c1), v2 = vm.loader._touched(v, img.shape[0], img.shape[1])
            assert 0 <= r0 < r1 <= img.shape[0] and 0 <= c0 < c1 <= img.shape[1]
            # every canvas cell of the window that holds image data maps inside the staged rectangle
            rows = np.arange(int(v2["win_h"])) + int(v2["wi"])
            cols = np.arange(int(v2["win_w"])) + int(v2["wj"])
            rr = rows[(rows >= int(v2["vi0"])) & (rows < int(v2["vi1"]))] - int(v2["vi0"]) + int(v2["si"])
            cc = cols[(cols >= int(v2["vj0"])) & (cols < int(v2["vj1"]))] - int(v2["vj0"]) + int(v2["sj"])
            if rr.size and cc.size:                       # the kernel reads only where row AND column hold image data
                assert rr.min() >= 0 and rr.max() < r1 - r0 and cc.min() >= 0 and cc.max() < c1 - c0
        if trial % 5 == 0:
            ref = LO.simple_sample(fg, bg, size, np.random.RandomState(trial))
            assert np.array_equal(LO.resize_linear_f64(view_window(fg[..., :3], fgv), size), ref[3])
            assert np.array_equal(LO.resize_linear_f64(view_window(bg, bgv), size) - LO.VGG_MEAN, ref[1])

"""GPU parity tests: the CUDA path (through the ctypes C ABI) against the golden vectors of
the unmodified reference and against the oracle on seeded inputs.

Bars (BASELINE.json north_star / SURVEY 8d): .flo parse, uint8 images and the occlusion mask
bit-exact; float alpha and composites |got-ref| <= 1e-5*|ref| + atol (1e-6 alpha, 1e-5 on the
0..255 composites); change_illumination +-1 LSB.  TPS-resampled uint8 outputs are bit-exact
up to knife-edge samples (float64 transform differs from numpy's by ~1e-11 px), which are
counted and bounded, never masked.
"""
import os

import numpy as np
import pytest
import torch

import vm_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(got, ref, atol):
    return np.allclose(got, ref, rtol=RTOL, atol=atol)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ---------------------------------------------------------------------------- flow (a-1..a-3)

@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_dropin_flow_golden(vm, golden, tag, capsys):
    fg, fb, ff = golden[f"flow_{tag}_fg"], golden[f"flow_{tag}_fb"], golden[f"flow_{tag}_ff"]
    alpha, bgr = fg[..., 3] / 255., np.ascontiguousarray(fg[..., :3])
    wa = vm.flow.warp_img(alpha, fb)
    assert isinstance(wa, np.ndarray) and wa.dtype == np.float64
    assert np.array_equal(wa, golden[f"flow_{tag}_warp_alpha"])                 # float64: bit-equal
    assert np.array_equal(vm.flow.warp_img(alpha.astype(np.float32), fb), golden[f"flow_{tag}_warp_alpha_f32"])
    assert np.array_equal(vm.flow.warp_img(np.ascontiguousarray(fg[..., 3]), fb), golden[f"flow_{tag}_warp_alpha_u8"])
    assert np.array_equal(vm.flow.warp_bgr(bgr, fb), golden[f"flow_{tag}_warp_bgr"])
    a = wa.copy()
    r = vm.flow.correct_alpha(fb, ff, a)
    assert r is a, "correct_alpha must mutate and return its argument"
    assert np.array_equal(a, golden[f"flow_{tag}_corrected"])
    assert str(tuple(ff.shape)) in capsys.readouterr().out                      # flow.py:40 print


def test_flow_special_values(vm, golden):
    fg, fb = golden["special_fg"], golden["special_fb"]
    assert np.array_equal(vm.flow.warp_bgr(np.ascontiguousarray(fg[..., :3]), fb), golden["special_warp_bgr"])
    assert np.array_equal(vm.flow.warp_img(fg[..., 3] / 255., fb), golden["special_warp_alpha"])


def test_flow_errors(vm):
    h, w = 16, 20
    a = np.ones((h, w))
    fb = np.zeros((h, w, 2), np.float32)
    ff = np.zeros((h, w, 2), np.float32)
    with pytest.raises(AssertionError):
        vm.flow.warp_img(np.zeros((h, w, 3)), fb)                               # flow.py:11
    fb2 = fb.copy(); fb2[3, 4, 0] = -5.0 * w
    with pytest.raises(IndexError):
        vm.flow.correct_alpha(fb2, ff, a.copy())
    with pytest.raises(IndexError):
        O.correct_alpha(fb2, ff, a.copy())
    fb3 = fb.copy(); fb3[2, 2, 1] = np.nan
    with pytest.raises(ValueError):
        vm.flow.correct_alpha(fb3, ff, a.copy())
    # negative coordinates wrap Python-style instead of raising
    fb4 = fb.copy(); fb4[:, :2, 0] = -3.0
    ff4 = np.random.default_rng(0).normal(0, 9, (h, w, 2)).astype(np.float32)
    got = vm.flow.correct_alpha(fb4, ff4, a.copy())
    assert np.array_equal(got, O.correct_alpha(fb4, ff4, a.copy()))
    assert np.array_equal(got, O.correct_alpha_loop(fb4, ff4, a.copy()))


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (5, 1), (37, 53), (64, 130)])
def test_flow_ragged_sizes_vs_oracle(vm, shape):
    h, w = shape
    rng = np.random.default_rng(h * 100 + w)
    fg = O.synth_frame(h + w, max(h, 2), max(w, 2))[:h, :w]
    fb = rng.normal(0, 3, (h, w, 2)).astype(np.float32)
    ff = (-fb + rng.normal(0, 6, (h, w, 2))).astype(np.float32)
    fb[0, 0] = 0
    alpha, bgr = fg[..., 3] / 255., np.ascontiguousarray(fg[..., :3])
    assert np.array_equal(vm.flow.warp_bgr(bgr, fb), O.warp_bgr(bgr, fb))
    assert np.array_equal(vm.flow.warp_img(alpha, fb), O.warp_img(alpha, fb))
    try:
        ref = O.correct_alpha(fb, ff, alpha.copy())
    except IndexError:
        with pytest.raises(IndexError):
            vm.flow.correct_alpha(fb, ff, alpha.copy())
    else:
        assert np.array_equal(vm.flow.correct_alpha(fb, ff, alpha.copy()), ref)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fused_c2_golden(vm, golden, tag):
    fg, fb, ff = golden[f"flow_{tag}_fg"], golden[f"flow_{tag}_fb"], golden[f"flow_{tag}_ff"]
    bgr, alpha, st = vm.pipeline.flow_warp_mask(dev(fg[None]), dev(fb[None]), dev(ff[None]))
    assert np.array_equal(bgr[0].cpu().numpy(), golden[f"flow_{tag}_warp_bgr"])
    assert close(alpha[0].cpu().numpy(), golden[f"flow_{tag}_corrected"], 1e-6)
    assert int(st[0]) == 0 and int(st[1]) == 0
    # without the forward flow: plain warp
    _, alpha2, _ = vm.pipeline.flow_warp_mask(dev(fg[None]), dev(fb[None]), None)
    assert close(alpha2[0].cpu().numpy(), golden[f"flow_{tag}_warp_alpha"], 1e-6)


@pytest.mark.parametrize("shape", [(3, 5), (33, 61), (40, 128), (41, 132)])
def test_fused_c2_ragged_clip_vs_oracle(vm, shape):
    h, w = shape
    n = 3
    frames = np.stack([O.synth_frame(7 * k + h, max(h, 34), max(w, 34))[:h, :w] for k in range(n)])
    flows = [O.synth_flows(k + w, max(h, 34), max(w, 34)) for k in range(n)]
    fb = np.stack([f[0][:h, :w] for f in flows]); ff = np.stack([f[1][:h, :w] for f in flows])
    bgr, alpha, st = vm.pipeline.flow_warp_mask(dev(frames), dev(fb), dev(ff))
    bgr, alpha = bgr.cpu().numpy(), alpha.cpu().numpy()
    for k in range(n):
        rb, ra = O.pipeline_c2(frames[k], fb[k], ff[k])
        assert np.array_equal(bgr[k], rb)
        assert close(alpha[k], ra, 1e-6)


# ------------------------------------------------------------------------ composite (a-12/13)

def test_composite_golden(vm, golden):
    fg = golden["wi_fg"]
    bgr, alpha = np.ascontiguousarray(fg[..., :3]), fg[..., 3] / 255.
    got = vm.reader.create_composite_image(bgr, golden["aug_bg"], alpha)
    assert got.dtype == np.float64 and np.array_equal(got, golden["cmp_out"])


@pytest.mark.parametrize("k", [1, 2])
def test_reference_owned_cmp_png(vm, golden, k):
    img8 = O.fg_from_uint16(golden[f"cmp{k}_raw16"])
    alpha, bgr = O.split_fg(img8)
    cmp_ = vm.reader.create_composite_image(np.ascontiguousarray(bgr), golden[f"cmp{k}_bg"], alpha)
    assert np.array_equal(np.rint(cmp_).astype(np.uint8), golden[f"cmp{k}_gold"])


def test_read_flow_and_fg(vm, golden, tmp_path, capsys):
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    flow = vm.reader.read_flow(os.path.join(here, "golden", "tiny.flo"))
    assert flow.dtype == np.float32 and np.array_equal(flow, golden["flo_tiny"])
    raw = bytearray(open(os.path.join(here, "golden", "tiny.flo"), "rb").read())
    raw[0] ^= 0xFF
    p = tmp_path / "bad.flo"; p.write_bytes(bytes(raw))
    assert np.array_equal(vm.reader.read_flow(str(p)), flow)
    assert "ERROR: invalid key" in capsys.readouterr().out
    p2 = tmp_path / "short.flo"; p2.write_bytes(bytes(raw[:-8]))
    with pytest.raises(ValueError):
        vm.reader.read_flow(str(p2))
    cv2 = pytest.importorskip("cv2")
    p3 = str(tmp_path / "fg16.png")
    cv2.imwrite(p3, golden["cmp1_raw16"])
    alpha, bgr = vm.reader.read_fg_img(p3)
    img8 = O.fg_from_uint16(golden["cmp1_raw16"])
    assert np.array_equal(bgr, img8[..., :3]) and np.array_equal(alpha, img8[..., 3] / 255.)


# ----------------------------------------------------------------------------- TPS (a-5..a-7)

def knife_edges(got, ref):
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 1, "a TPS-resampled uint8 sample is off by more than a knife-edge flip"
    return int(np.count_nonzero(d))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_tps_golden(vm, golden, tag):
    grid, dgrid, fg = golden[f"tps_{tag}_grid"], golden[f"tps_{tag}_defgrid"], golden[f"tps_{tag}_fg"]
    h, w = fg.shape[:2]
    t = vm.tps._make_inverse_warp(grid, dgrid, (0, 0, h, w), 2)
    t0, t1 = t[0].cpu().numpy(), t[1].cpu().numpy()
    assert t0.shape == (h + 1, w + 1)
    err = max(np.abs(t0 - golden[f"tps_{tag}_t0"]).max(), np.abs(t1 - golden[f"tps_{tag}_t1"]).max())
    assert err < 5e-10, f"TPS transform differs from the reference by {err} px"
    res = vm.tps.warp_images(grid, dgrid, [fg[..., 0], fg[..., 1], fg[..., 2], fg[..., 3] / 255.], (0, 0, h, w))
    flips = sum(knife_edges(res[c], golden[f"tps_{tag}_{k}"]) for c, k in enumerate("bgr"))
    assert flips <= 1, f"{flips} knife-edge flips on a {h}x{w} frame"
    assert res[3].dtype == np.float64 and res[3].shape == (h + 1, w + 1)
    assert np.allclose(res[3], golden[f"tps_{tag}_alpha"], rtol=RTOL, atol=1e-6)


def test_map_coordinates_exact_vs_oracle(vm):
    rng = np.random.default_rng(2)
    img8 = rng.integers(0, 256, (33, 47), dtype=np.uint8)
    img64 = rng.random((33, 47))
    t0 = rng.uniform(-2, 35, (40, 50)); t1 = rng.uniform(-2, 49, (40, 50))
    t0[0, :4] = (0.0, 32.0, np.nextafter(32.0, 64), 2.5); t1[0, :4] = (46.0, 0.5, 3.0, 7.5)
    t0[1, :3] = (0.059983768309400186, 1e-300, 5.0); t1[1, :3] = (2.2471181058956478, 3.0, 1e-300)
    P = vm.pipeline
    assert np.array_equal(P.map_coordinates(dev(img8), dev(t0), dev(t1)).cpu().numpy(),
                          O.map_coordinates_linear(img8, t0, t1))
    assert np.array_equal(P.map_coordinates(dev(img64), dev(t0), dev(t1)).cpu().numpy(),
                          O.map_coordinates_linear(img64, t0, t1))


def test_tps_log_accuracy(vm):
    """Coarse-grid evaluation against the oracle's numpy evaluation on a 1080p-sized grid."""
    h, w = 1080, 1920
    grid, dgrid = O.synth_grids(3, h, w, 5)
    C = O.tps_coefficients(dgrid, grid)
    xs, ys, cx, cy = O.tps_coarse_axes(h, w)
    sl = slice(0, None, 7)
    X, Y = np.meshgrid(cx[sl], cy[sl], indexing="ij")
    ref0 = O.tps_eval(C[:, 0], dgrid, X, Y); ref1 = O.tps_eval(C[:, 1], dgrid, X, Y)
    P = vm.pipeline
    plan = P.get_plan((0, 0, h, w), 2)
    coarse = P.tps_coarse(dev(dgrid[None]), dev(C[None]), plan)[0].cpu().numpy()
    err = max(np.abs(coarse[0][sl, sl] - ref0).max(), np.abs(coarse[1][sl, sl] - ref1).max())
    assert err < 5e-10, f"coarse TPS transform off by {err} px at 1080p"


# ------------------------------------------------------------------- affine / augment (a-8..11)

def test_warp_image_golden(vm, golden):
    fg = golden["wi_fg"]
    h, w = fg.shape[:2]
    bgr, alpha = np.ascontiguousarray(fg[..., :3]), fg[..., 3] / 255.
    grids = (golden["wi_grid"], golden["wi_defgrid"])
    p1 = ((3, -2), 7.5, 1.1, (52, 31))
    p2 = ((-4, 5), 0., 1.07, (w // 2, h // 2))
    A = vm.augmentation
    assert np.array_equal(A.warp_image(bgr, p1), golden["wi_p1_bgr"])
    assert np.array_equal(A.warp_image(alpha, p1), golden["wi_p1_alpha"])       # float64 bit-equal
    assert np.array_equal(A.warp_image(bgr, p2), golden["wi_p2_bgr"])
    assert knife_edges(A.warp_image(bgr, p1, thin=grids), golden["wi_p1_thin_bgr"]) <= 1
    assert np.allclose(A.warp_image(alpha, p1, thin=grids), golden["wi_p1_thin_alpha"], rtol=RTOL, atol=1e-6)


def test_warp_affine_random_vs_oracle(vm):
    rng = np.random.default_rng(1)
    src8 = rng.integers(0, 256, (45, 70, 3), dtype=np.uint8)
    src64 = rng.random((46, 71))
    P = vm.pipeline
    for _ in range(6):
        M = O.rotation_matrix_2d((rng.uniform(0, 70), rng.uniform(0, 45)), rng.uniform(-30, 30), rng.uniform(0.7, 1.3))
        assert np.array_equal(P.warp_affine(dev(src8), M, (70, 45)).cpu().numpy(), O.warp_affine(src8, M, (70, 45)))
        assert np.array_equal(P.warp_affine(dev(src64), M, (64, 50)).cpu().numpy(), O.warp_affine(src64, M, (64, 50)))


def test_illumination_stats_golden(vm, golden):
    A = vm.augmentation
    got = A.change_illumination(golden["ci_bgr"], 1.03, 0.8, -0.02)
    vec = vm._native.hsv_vec()                       # the fixtures were made on an AVX2 host (32 pixels per SIMD step of cv2)
    assert np.array_equal(got, O.change_illumination(golden["ci_bgr"], 1.03, 0.8, -0.02, vec))
    if vec == 32:
        assert np.array_equal(got, golden["ci_out"])
    else:
        assert np.abs(got.astype(int) - golden["ci_out"].astype(int)).max() <= 1
    alpha = golden["wi_fg"][..., 3] / 255.
    assert A.object_size(alpha) == float(golden["stats_size"])
    assert tuple(A.fg_center(alpha)) == tuple(golden["stats_center"])
    with pytest.raises(ValueError):
        A.fg_center(np.zeros((4, 4)))


def test_augment_golden(vm, golden):
    fg = golden["wi_fg"]
    bgr, alpha = np.ascontiguousarray(fg[..., :3]), fg[..., 3] / 255.
    np.random.seed(77)
    nfg, nbg, nal = vm.augmentation.augment(bgr, golden["aug_bg"], alpha)
    after = np.random.uniform()
    np.random.seed(77)
    O.augment_params(bgr.shape[0], bgr.shape[1], alpha)
    assert after == np.random.uniform(), "augment must consume the reference's 40 RNG draws"
    assert nal.dtype == np.float64 and np.allclose(nal, golden["aug_alpha_out"], rtol=RTOL, atol=1e-6)
    if vm._native.hsv_vec() == 32:                    # same SIMD width as the host that made the fixtures: exact
        assert np.array_equal(nbg, golden["aug_bg_out"])
        # fg: bit-exact up to TPS knife-edge flips (<= 2 pixels), which the colour transform may amplify
        assert np.count_nonzero((nfg != golden["aug_fg_out"]).any(axis=2)) <= 2
    else:
        assert np.abs(nbg.astype(int) - golden["aug_bg_out"].astype(int)).max() <= 1
        assert np.count_nonzero(np.abs(nfg.astype(int) - golden["aug_fg_out"].astype(int)) > 1) <= 2


# ------------------------------------------------------------------------- fused C3 / C4

VARIANTS = [4, 5, 1]   # fused_variant: lean split pipeline (default), single-pass kernel, per-pixel gather fallback


@pytest.fixture(params=VARIANTS, ids=["lean", "fuse", "gather"])
def variant(request, vm):
    vm.pipeline.set_fused_variant(request.param)
    yield request.param
    vm.pipeline.set_fused_variant(vm.pipeline.DEFAULT_VARIANT)


@pytest.fixture
def lean(vm):
    """the lean split pipeline (fused_variant 4) for the tests of its own knobs"""
    vm.pipeline.set_fused_variant(4)
    yield
    vm.pipeline.set_fused_variant(vm.pipeline.DEFAULT_VARIANT)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fused_c4_golden(vm, golden, tag, variant):
    fg, fb, ff = golden[f"flow_{tag}_fg"], golden[f"flow_{tag}_fb"], golden[f"flow_{tag}_ff"]
    grids = (golden[f"c4_{tag}_grid"], golden[f"c4_{tag}_defgrid"])
    P = vm.pipeline
    ctrl, coef = P.solve_grids([grids])
    out, st = P.flow_tps_composite(dev(fg[None]), dev(fb[None]), dev(ff[None]), dev(golden[f"c4_{tag}_bg"][None]),
                                   ctrl, coef)
    out = out[0].cpu().numpy()
    assert close(out[..., 3], golden[f"c4_{tag}_alpha"], 1e-6)
    ref = golden[f"c4_{tag}_cmp"]
    bad = ~np.isclose(out[..., :3], ref, rtol=RTOL, atol=1e-5)
    assert np.count_nonzero(bad) <= 1, "composite differs beyond knife-edge flips"


@pytest.mark.parametrize("shape,n_ctrl", [((64, 64), 4), ((61, 83), 5), ((128, 96), 5), ((200, 333), 5), ((4, 6), 2)])
def test_fused_c3_c4_vs_oracle(vm, shape, n_ctrl, variant):
    h, w = shape
    n = 2
    P = vm.pipeline
    frames = [O.synth_frame(50 + k, h, w) for k in range(n)]
    flows = [O.synth_flows(50 + k, h, w) for k in range(n)]
    grids = [O.synth_grids(50 + k, h, w, n_ctrl) for k in range(n)]
    bgs = np.stack([O.synth_background(k, h, w) for k in range(n)])
    ctrl, coef = P.solve_grids(grids)
    fg_d = dev(np.stack(frames))
    out3, _ = P.tps_composite(fg_d, dev(bgs), ctrl, coef)
    out4, _ = P.flow_tps_composite(fg_d, dev(np.stack([f[0] for f in flows])), dev(np.stack([f[1] for f in flows])),
                                   dev(bgs), ctrl, coef)
    out3, out4 = out3.cpu().numpy(), out4.cpu().numpy()
    for k in range(n):
        r3, a3 = O.pipeline_c3(frames[k], grids[k], bgs[k])
        r4, a4 = O.pipeline_c4(frames[k], flows[k][0], flows[k][1], grids[k], bgs[k])
        for got, rc, ra in ((out3[k], r3, a3), (out4[k], r4, a4)):
            assert close(got[..., 3], ra, 1e-6)
            assert np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1


# -------------------------------------------------------- full-size, size-independent properties

def test_1080p_properties(vm):
    h, w, n = 1080, 1920, 2
    P = vm.pipeline
    frames = np.stack([O.synth_frame(900 + k, h, w) for k in range(n)])
    fg_d = dev(frames)
    bg = O.synth_background(9, h, w)
    # (1) integer-shift flow == shifted frame, exactly
    flow = np.zeros((n, h, w, 2), np.float32); flow[..., 0] = 3.0; flow[..., 1] = -2.0
    bgr, alpha, _ = P.flow_warp_mask(fg_d, dev(flow), None)
    bgr, alpha = bgr.cpu().numpy(), alpha.cpu().numpy()
    assert np.array_equal(bgr[:, 2:, :w - 3], frames[:, :h - 2, 3:, :3])
    assert (bgr[:, :2] == 0).all() and (bgr[:, :, w - 3:] == 0).all()
    assert np.allclose(alpha[:, 2:, :w - 3], frames[:, :h - 2, 3:, 3] / 255., rtol=RTOL, atol=1e-6)
    # (2) fused C2 == the generic per-function kernels (independent implementations)
    fb, ff = O.synth_flows(900, h, w)
    fbd, ffd = dev(fb[None]), dev(ff[None])
    bgr2, alpha2, st = P.flow_warp_mask(fg_d[:1], fbd, ffd)
    g_bgr = P.flow_warp(fg_d[0, :, :, :3].contiguous(), fbd[0])
    g_alpha = P.flow_warp((fg_d[0, :, :, 3].double() / 255.).contiguous(), fbd[0])
    mask, st2 = P.occlusion_mask(fbd[0], ffd[0])
    P.apply_mask(g_alpha, mask)
    assert torch.equal(bgr2[0], g_bgr)
    assert torch.allclose(alpha2[0].double(), g_alpha, rtol=RTOL, atol=1e-6)
    assert 0.02 < float(mask.float().mean()) < 0.5, "synthetic flows must trigger the 15 px test"
    # (3) zero flow + undeformed TPS grid + fused composite == plain composite
    grid, _ = O.synth_grids(0, h, w, 5)
    ctrl, coef = P.solve_grids([(grid, grid)])
    zero = torch.zeros((1, h, w, 2), dtype=torch.float32, device="cuda")
    out, _ = P.flow_tps_composite(fg_d[:1], zero, zero, dev(bg[None]), ctrl, coef)
    out3, _ = P.tps_composite(fg_d[:1], dev(bg[None]), ctrl, coef)
    ref = O.create_composite_image(frames[0, ..., :3], bg, frames[0, ..., 3] / 255.)
    inner = (slice(1, -1), slice(1, -1))       # border coords sit on the in/out knife edge
    for o in (out[0].cpu().numpy(), out3[0].cpu().numpy()):
        assert close(o[inner][..., :3], ref[inner], 1e-5)
        assert close(o[inner][..., 3], (frames[0, ..., 3] / 255.)[inner], 1e-6)
    # (4) a 1080p row band of the full C4 pipeline against the oracle (oracle on a crop is not
    # possible for TPS, so compare the whole frame but only one frame)
    grids = O.synth_grids(901, h, w, 5)
    ctrl, coef = P.solve_grids([grids])
    out4, _ = P.flow_tps_composite(fg_d[:1], fbd, ffd, dev(bg[None]), ctrl, coef)
    rc, ra = O.pipeline_c4(frames[0], fb, ff, grids, bg)
    got = out4[0].cpu().numpy()
    assert close(got[..., 3], ra, 1e-6)
    flips = np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5))
    assert flips <= 3, f"{flips} composite samples differ beyond tolerance at 1080p"


# ------------------------------------------------------------------ lean pipeline (default variant)

def _lean_case(h, w, n, n_ctrl, seed=300, stretch=1.0):
    frames = np.stack([O.synth_frame(seed + k, h, w) for k in range(n)])
    flows = [O.synth_flows(seed + k, h, w) for k in range(n)]
    grids = []
    for k in range(n):
        g, d = O.synth_grids(seed + k, h, w, n_ctrl)
        grids.append((g, g + (d - g) * stretch))
    bgs = np.stack([O.synth_background(k, h, w) for k in range(n)])
    return frames, np.stack([f[0] for f in flows]), np.stack([f[1] for f in flows]), grids, bgs


def test_lean_partition_independent(vm, lean):
    """The result must not depend on how the work is cut: frames per stage round, coarse rows per
    spline unit, resampling occupancy target, and whether a tile's source box is staged in shared
    memory (box capacity 64 entries forces every tile onto the global-gather path), and whether the resampling stage
    stages through tensor maps (k_lean_fine_tm, the default) or through one bulk copy per row (k_lean_fine)."""
    h, w, n = 200, 333 + 3, 5                       # w % 4 == 0 so that the box path is eligible
    P, Nt = vm.pipeline, vm._native
    frames, fb, ff, grids, bgs = _lean_case(h, w, n, 5)
    ctrl, coef = P.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    base, st0 = P.flow_tps_composite(*args)
    base3, _ = P.tps_composite(args[0], args[3], ctrl, coef)
    assert int(st0[4]) < 4, "most tiles of a mild grid must take the shared-memory box path"
    defaults = {"lean_chunk": 0, "lean_rb": 0, "lean_minb": 4, "lean_box_cap": 0, "lean_fine_rows": 8, "lean_sub": 0,
                "lean_b1_warps": 16, "lean_b1_dyr": 1, "flow_stage_layout": 0, "lean_floors": 1,
                "lean_b1_ctas": 0, "lean_tmap": 1}
    configs = [{"lean_tmap": 0}, {"lean_tmap": 0, "lean_box_cap": 64}, {"lean_box_cap": 512}, {"lean_tmap": 0, "lean_minb": 3},
               {"lean_chunk": 1}, {"lean_chunk": 2}, {"lean_chunk": 3}, {"lean_rb": 4}, {"lean_rb": 16}, {"lean_rb": 32},
               {"lean_sub": 2}, {"lean_minb": 2}, {"lean_minb": 3}, {"lean_fine_rows": 3}, {"lean_fine_rows": 5},
               {"lean_box_cap": 64}, {"lean_b1_dyr": 0, "lean_b1_warps": 24}, {"lean_b1_warps": 6, "lean_b1_ctas": 40},
               {"flow_stage_layout": 1}, {"lean_floors": 0}, {"lean_floors": 0, "lean_box_cap": 64},
               {"lean_chunk": 2, "lean_b1_warps": 8, "lean_box_cap": 64}]
    try:
        for cfg in configs:
            for key, v in cfg.items():
                Nt.set_option(key, v)
            out, st = P.flow_tps_composite(*args)
            out3, _ = P.tps_composite(args[0], args[3], ctrl, coef)
            torch.cuda.synchronize()
            assert torch.equal(out, base) and torch.equal(out3, base3), f"{cfg} changes the result"
            if cfg.get("lean_box_cap") in (64, 512):
                assert int(st[4]) > 0, "a small box capacity must push tiles onto the gather path"
            if cfg == {"lean_floors": 0}:
                assert int(st[4]) == int(st0[4]), "the packed floors must select the same tile boxes as T itself"
            for key in cfg:
                Nt.set_option(key, defaults[key])
    finally:
        for key, v in defaults.items():
            Nt.set_option(key, v)


@pytest.mark.parametrize("n_ctrl", [3, 4, 6, 8])
def test_lean_control_point_counts_vs_oracle(vm, n_ctrl, lean):
    """N = 16 and 25 use the unrolled spline kernels, every other count the run-time one (N <= 64)."""
    h, w, n = 96, 128, 2
    P = vm.pipeline
    frames, fb, ff, grids, bgs = _lean_case(h, w, n, n_ctrl, seed=410)
    ctrl, coef = P.solve_grids(grids)
    out, _ = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    out = out.cpu().numpy()
    for k in range(n):
        rc, ra = O.pipeline_c4(frames[k], fb[k], ff[k], grids[k], bgs[k])
        assert close(out[k][..., 3], ra, 1e-6)
        assert np.count_nonzero(~np.isclose(out[k][..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1


def test_lean_stretched_grid_vs_oracle(vm, lean):
    """30 % displacements: many tiles' source boxes exceed shared memory (gather path) and control points
    come close to coarse grid points (generic spline path); both must still match the oracle."""
    h, w = 256, 320
    P = vm.pipeline
    frames, fb, ff, grids, bgs = _lean_case(h, w, 1, 5, seed=77, stretch=6.0)
    ctrl, coef = P.solve_grids(grids)
    out, st = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    rc, ra = O.pipeline_c4(frames[0], fb[0], ff[0], grids[0], bgs[0])
    got = out[0].cpu().numpy()
    assert close(got[..., 3], ra, 1e-6)
    assert np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1
    assert int(st[5]) == 0


def test_lean_control_point_on_grid_point(vm, lean):
    """A control point that coincides with a coarse grid point (r = 0: U = 0, tps.py:78-82) and one a hair
    away from it (r^2 below the log table) take the generic spline path of their unit."""
    h, w = 128, 160
    P = vm.pipeline
    frames, fb, ff, grids, bgs = _lean_case(h, w, 1, 4, seed=55)
    g, d = grids[0]
    d = d.copy()
    sx, sy = h / float(h // 2 - 1), w / float(w // 2 - 1)          # coarse grid steps (tps.py:47)
    d[5] = (20 * sx, 31 * sy)                                       # exactly on coarse point (20, 31)
    d[6] = (33 * sx + 1e-3, 40 * sy - 2e-3)                         # 2e-3 px away from coarse point (33, 40)
    grids = [(g, d)]
    ctrl, coef = P.solve_grids(grids)
    out, _ = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    rc, ra = O.pipeline_c4(frames[0], fb[0], ff[0], grids[0], bgs[0])
    got = out[0].cpu().numpy()
    assert close(got[..., 3], ra, 1e-6)
    assert np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1


def test_lean_stage_timing_and_launch_count(vm, lean):
    import ctypes
    h, w, n = 128, 192, 3
    P, Nt = vm.pipeline, vm._native
    lib = Nt.load()
    frames, fb, ff, grids, bgs = _lean_case(h, w, n, 5)
    ctrl, coef = P.solve_grids(grids)
    before = lib.vm_lean_launch_count()
    Nt.set_option("lean_timing", 1)
    try:
        P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 4)()
        Nt.check(lib.vm_lean_stage_ms(ctypes.cast(buf, ctypes.c_void_p)))
    finally:
        Nt.set_option("lean_timing", 0)
    assert all(0.0 < v < 1000.0 for v in buf)
    assert lib.vm_lean_launch_count() - before == 4          # spline, boxes, flow stage, resampling


def test_4k_identity_property(vm):
    """4K (config 4 size): zero flow + undeformed grid through the fused C4 path == plain composite."""
    h, w = 2160, 3840
    P = vm.pipeline
    frame = O.synth_frame(990, h, w)
    bg = O.synth_background(4, h, w)
    grid, _ = O.synth_grids(0, h, w, 5)
    ctrl, coef = P.solve_grids([(grid, grid)])
    zero = torch.zeros((1, h, w, 2), dtype=torch.float32, device="cuda")
    out, _ = P.flow_tps_composite(dev(frame[None]), zero, zero, dev(bg[None]), ctrl, coef)
    ref = O.create_composite_image(frame[..., :3], bg, frame[..., 3] / 255.)
    inner = (slice(2, -2), slice(2, -2))
    o = out[0].cpu().numpy()
    assert close(o[inner][..., :3], ref[inner], 1e-5)
    assert close(o[inner][..., 3], (frame[..., 3] / 255.)[inner], 1e-6)


# ------------------------------------------------------------------------------- edge cases

def test_fused_empty_clip(vm):
    """n = 0 frames: every fused entry point returns an empty result and launches nothing."""
    h, w = 32, 48
    P = vm.pipeline
    e = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device="cuda")
    bgr, alpha, _ = P.flow_warp_mask(e(0, h, w, 4), e(0, h, w, 2, dt=torch.float32), e(0, h, w, 2, dt=torch.float32))
    assert bgr.shape == (0, h, w, 3) and alpha.shape == (0, h, w)
    ctrl, coef = e(0, 25, 2, dt=torch.float64), e(0, 28, 2, dt=torch.float64)
    out, _ = P.flow_tps_composite(e(0, h, w, 4), e(0, h, w, 2, dt=torch.float32), e(0, h, w, 2, dt=torch.float32),
                                  e(1, h, w, 3), ctrl, coef)
    out3, _ = P.tps_composite(e(0, h, w, 4), e(1, h, w, 3), ctrl, coef)
    assert out.shape == (0, h, w, 4) and out3.shape == (0, h, w, 4)
    torch.cuda.synchronize()


def test_fused_wild_flows(vm):
    """Flows of sigma 50 / 500 px (most taps outside the frame, wrapped / clamped mask look-ups) through the
    fused C2 and C4 paths against the oracle; then NaN / Inf / 3e9 components: cv2.remap reads 0 there, and
    correct_alpha would raise on the NaN / Inf ones - those pixels are counted in the status block."""
    h, w, n = 72, 100, 2
    P = vm.pipeline
    rng = np.random.default_rng(5)
    frames, _, _, grids, bgs = _lean_case(h, w, n, 5, seed=610)
    for sigma in (50.0, 500.0):
        fb = rng.normal(0, sigma, (n, h, w, 2)).astype(np.float32)
        ff = rng.normal(0, sigma, (n, h, w, 2)).astype(np.float32)
        ok = True
        try:
            refs = [O.pipeline_c4(frames[k], fb[k], ff[k], grids[k], bgs[k]) for k in range(n)]
        except IndexError:
            ok = False                                     # a look-up below -H / -W: the reference raises
        ctrl, coef = P.solve_grids(grids)
        bgr, alpha, st2 = P.flow_warp_mask(dev(frames), dev(fb), dev(ff))
        out, st4 = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
        assert (int(st2[0]) == 0) == ok and (int(st4[0]) == 0) == ok
        if ok:
            for k in range(n):
                rb, ra = O.pipeline_c2(frames[k], fb[k], ff[k])
                assert np.array_equal(bgr[k].cpu().numpy(), rb) and close(alpha[k].cpu().numpy(), ra, 1e-6)
                got = out[k].cpu().numpy()
                assert close(got[..., 3], refs[k][1], 1e-6)
                assert np.count_nonzero(~np.isclose(got[..., :3], refs[k][0], rtol=RTOL, atol=1e-5)) <= 1
    fb = rng.normal(0, 3, (n, h, w, 2)).astype(np.float32)
    ff = (-fb).copy()
    fb[0, 5, 7, 0] = np.nan; fb[0, 9, 11, 1] = np.inf; fb[1, 20, 30, 0] = -np.inf; fb[1, 40, 50, 1] = 3e9
    bgr, alpha, st = P.flow_warp_mask(dev(frames), dev(fb), dev(ff))
    assert int(st[1]) == 3 and int(st[0]) == 0, "3 NaN/Inf pixels (the reference raises ValueError); +3e9 only clamps"
    for (k, i, j) in ((0, 5, 7), (0, 9, 11), (1, 20, 30), (1, 40, 50)):
        assert bgr[k, i, j].sum().item() == 0 and alpha[k, i, j].item() == 0.0      # cv2.remap: NaN / Inf / huge -> 0
    clean = np.ones((n, h, w), bool)
    for (k, i, j) in ((0, 5, 7), (0, 9, 11), (1, 20, 30), (1, 40, 50)):
        clean[k, i, j] = False
    # the same through the C4 pipeline: counted, and the output stays finite
    ctrl, coef = P.solve_grids(grids)
    out, st4 = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    assert int(st4[1]) == 3 and int(st4[0]) == 0 and bool(torch.isfinite(out).all())
    # NaN in the forward flow where a pixel looks it up, and a look-up below -H: ValueError / IndexError
    fb2 = np.zeros((1, h, w, 2), np.float32); ff2 = np.zeros((1, h, w, 2), np.float32)
    ff2[0, 10, 10, 0] = np.nan
    fb2[0, 30, 30, 1] = -(h + 40.0)
    _, _, st = P.flow_warp_mask(dev(frames[:1]), dev(fb2), dev(ff2))
    assert int(st[1]) == 1 and int(st[0]) == 1
    fb_ok = np.where(np.isfinite(fb) & (np.abs(fb) < 1e9), fb, 0).astype(np.float32)
    for k in range(n):
        rb, ra = O.pipeline_c2(frames[k], fb_ok[k], ff[k])
        assert np.array_equal(bgr[k].cpu().numpy()[clean[k]], rb[clean[k]])


# ------------------------------------------------------------------ batched augmentation (config 5)

@pytest.mark.parametrize("shape", [(96, 128), (61, 83), (200, 333)])
def test_augment_clip_matches_sequential_augment(vm, shape):
    """augment_clip == [augment(frame) for frame in clip] with the same global np.random state: same draws in
    the same order, same uint8 images (up to knife-edge samples of the TPS stage, whose up-sampling order
    differs by ~1e-13 px from the drop-in kernel) and the same alpha within 1e-5."""
    h, w = shape
    n = 3
    A = vm.augmentation
    frames = np.stack([O.synth_frame(700 + k, h, w) for k in range(n)])
    bgs = np.stack([O.synth_background(20 + k, h, w) for k in range(n)])
    np.random.seed(2024)
    seq = [A.augment(np.ascontiguousarray(frames[k, ..., :3]), bgs[k], frames[k, ..., 3] / 255.) for k in range(n)]
    after_seq = np.random.uniform()
    np.random.seed(2024)
    nfg, nbg, nal = A.augment_clip(frames, bgs)
    assert np.random.uniform() == after_seq, "augment_clip must consume exactly the draws of n augment() calls"
    assert nfg.dtype == np.uint8 and nbg.dtype == np.uint8 and nal.dtype == np.float32
    for k in range(n):
        rfg, rbg, ral = seq[k]
        assert np.array_equal(nbg[k], rbg)
        assert np.count_nonzero(np.any(nfg[k] != rfg, axis=-1)) <= 2
        assert np.allclose(nal[k], ral, rtol=RTOL, atol=1e-6)
    np.random.seed(2024)
    again = A.augment_clip(frames, bgs, stats=A.alpha_stats(frames))          # precomputed statistics: same result
    assert all(np.array_equal(x, y) for x, y in zip(again, (nfg, nbg, nal)))
    # float64 alpha: the reference's dtype and operation order (scipy map_coordinates, then cv2.warpAffine on
    # float64) - equal to the drop-in's float64 alpha up to the ~1e-13 px difference of the transform
    np.random.seed(2024)
    nfg64, nbg64, nal64 = A.augment_clip(frames, bgs, alpha_dtype=torch.float64)
    assert nal64.dtype == np.float64 and np.array_equal(nfg64, nfg) and np.array_equal(nbg64, nbg)
    for k in range(n):
        assert np.abs(nal64[k] - seq[k][2]).max() <= 1e-9
        # (bit equality is out of reach even for saturated alpha: sum(w_k * 1.0) in scipy's order is 1 or 1 - 2^-53
        #  depending on the last bits of the weights, i.e. on the ~1e-13 px by which the two transforms differ)
        assert np.all(nal64[k][seq[k][2] == 0.] == 0.)


def test_augment_clip_golden(vm, golden):
    """One-frame clip against the golden vectors of the unmodified reference's augment()."""
    bgra, bg = golden["wi_fg"], golden["aug_bg"]                      # BGRA uint8: alpha = A/255 as in read_fg_img
    np.random.seed(77)
    nfg, nbg, nal = vm.augmentation.augment_clip(bgra[None], bg[None])
    assert np.allclose(nal[0], golden["aug_alpha_out"], rtol=RTOL, atol=1e-6)
    if vm._native.hsv_vec() == 32:
        assert np.array_equal(nbg[0], golden["aug_bg_out"])
        assert np.count_nonzero((nfg[0] != golden["aug_fg_out"]).any(axis=2)) <= 2
    else:
        assert np.abs(nbg[0].astype(int) - golden["aug_bg_out"].astype(int)).max() <= 1
        assert np.count_nonzero(np.abs(nfg[0].astype(int) - golden["aug_fg_out"].astype(int)) > 1) <= 2


def test_augmentation_writer_matches_sequential_augment(vm, tmp_path, capsys):
    """augmentation.augmentation (reference augmentation.py:138-166): same files and same np.random order as
    the per-variant loop of the reference, here emulated with the drop-in augment() (itself pinned by the
    reference golden vectors)."""
    import cv2
    A = vm.augmentation
    h, w = 52, 76
    dim, voc, sig = tmp_path / "DIM", tmp_path / "VOC", tmp_path / "SIG"
    for d in (dim / "fg" / "DIM_TEST", dim / "fg" / "DIM_TRAIN", voc, sig / "fg" / "augmented", sig / "bg" / "augmented"):
        os.makedirs(d)
    assert cv2.imwrite(str(dim / "fg" / "DIM_TEST" / "t0.png"), O.synth_frame(1, h, w))
    assert cv2.imwrite(str(dim / "fg" / "DIM_TRAIN" / "r0.png"), O.synth_frame(2, h, w))
    for k, shape in enumerate(((h, w), (40, 50), (90, 120))):
        assert cv2.imwrite(str(voc / f"v{k}.png"), O.synth_background(k, *shape))
    old = A.N_VARIANTS, A.VARIANT_BATCH
    A.N_VARIANTS, A.VARIANT_BATCH = 7, 3
    try:
        np.random.seed(31)
        A.augmentation(str(dim), str(voc), str(sig))
    finally:
        A.N_VARIANTS, A.VARIANT_BATCH = old
    assert "Processing image" in capsys.readouterr().out
    # sequential emulation with the same seed and the same directory listing order
    np.random.seed(31)
    paths = [str(dim / "fg" / f / n) for f in ("DIM_TEST", "DIM_TRAIN") for n in os.listdir(dim / "fg" / f)]
    voc_list = [str(voc / n) for n in os.listdir(voc)]
    n_alpha_off = 0
    for p in paths:
        alpha, fg = vm.reader.read_fg_img(p)
        name = os.path.basename(p).split(".")[0]
        ref = cv2.imread(str(sig / "fg" / "augmented" / f"{name}_fg_ref.png"), cv2.IMREAD_UNCHANGED)
        assert np.array_equal(ref[..., :3], fg) and np.array_equal(ref[..., 3], (255. * alpha).astype(np.uint8))
        for i in range(7):
            bg = cv2.imread(voc_list[np.random.randint(len(voc_list))])
            bg = cv2.resize(bg, dsize=(w, h), interpolation=cv2.INTER_LINEAR)
            nfg, nbg, nal = A.augment(np.ascontiguousarray(fg), bg, alpha)
            got_fg = cv2.imread(str(sig / "fg" / "augmented" / f"{name}_fg_{i:04d}.png"), cv2.IMREAD_UNCHANGED)
            assert np.array_equal(cv2.imread(str(sig / "bg" / "augmented" / f"{name}_bg_ref_{i:04d}.png")), bg)
            assert np.array_equal(cv2.imread(str(sig / "bg" / "augmented" / f"{name}_bg_{i:04d}.png")), nbg)
            assert np.array_equal(got_fg[..., :3], nfg)
            da = np.abs(got_fg[..., 3].astype(int) - (255. * nal).astype(np.uint8).astype(int))
            assert da.max() <= 1
            n_alpha_off += int(da.sum())
    # (255. * alpha).astype(uint8) truncates values that sit within an ulp of an integer wherever alpha is locally
    # constant (255 * 0.9999999999999999 -> 254), so these bytes follow the last bits of the interpolation weights;
    # they agree to +-1 level, and mostly exactly
    assert n_alpha_off <= 0.02 * 14 * h * w


# ------------------------------------------------------------------ BASELINE config 1 (real test images)

def test_config1_window_real_images(vm, capsys):
    """BASELINE config 1 on a window of the reference's own frames (tests/golden/make_c1_golden.py: in0063 warped
    onto in0062 with DIS stand-in flows, consistency mask, composite onto sea.jpg), through the drop-in functions
    and through the fused C2 entry point."""
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_golden.npz")) as z:
        c1 = {k: z[k] for k in z.files}
    fg = c1["fg63_bgra"]
    alpha, bgr = fg[..., 3] / 255., np.ascontiguousarray(fg[..., :3])
    wa = vm.flow.warp_img(alpha, c1["backward"])
    assert wa.dtype == np.float64 and np.array_equal(wa, c1["warp_alpha"])
    wb = vm.flow.warp_bgr(bgr, c1["backward"])
    assert np.array_equal(wb, c1["warp_bgr"])
    ca = vm.flow.correct_alpha(c1["backward"], c1["forward"], wa)
    assert ca is wa and np.array_equal(ca, c1["corrected"])
    assert str(tuple(c1["forward"].shape)) in capsys.readouterr().out
    cmp_ = vm.reader.create_composite_image(wb, c1["bg"], ca)
    assert close(cmp_, c1["composite"], 1e-5)
    # fused: one pass over the BGRA frame
    ob, oa, st = vm.pipeline.flow_warp_mask(dev(fg[None]), dev(c1["backward"][None]), dev(c1["forward"][None]))
    assert np.array_equal(ob[0].cpu().numpy(), c1["warp_bgr"])
    assert close(oa[0].cpu().numpy(), c1["corrected"], 1e-6)
    mask = O.occlusion_mask(c1["backward"], c1["forward"])
    assert np.array_equal(c1["corrected"] == 0, (c1["warp_alpha"] == 0) | mask)
    assert int(st[0]) == 0 and int(st[1]) == 0
    m_d, st_m = vm.pipeline.occlusion_mask(dev(c1["backward"]), dev(c1["forward"]))
    assert np.array_equal(m_d.cpu().numpy().astype(bool), mask) and int(st_m[2]) == int(mask.sum())


def test_fused_c4_cuda_graph_replay(vm):
    """The fused C4 entry point is capturable (no allocation, no synchronisation, no host round trip inside the
    library): a captured launch sequence replayed on new input values gives the eager result, and for small
    frames - where the four launches are latency bound - the replay is not slower than eager launches."""
    P = vm.pipeline
    h, w, n = 120, 160, 4
    frames, fb, ff, grids, bgs = _lean_case(h, w, n, 5)
    ctrl, coef = P.solve_grids(grids)
    fg_d, fb_d, ff_d, bg_d = dev(frames), dev(fb), dev(ff), dev(bgs)
    plan = P.get_plan((0, 0, h, w), 2, fg_d.device)
    out = torch.empty((n, h, w, 4), dtype=torch.float32, device="cuda")
    status = vm._native.new_status()
    ref, _ = P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef, plan=plan)          # also warms the library up
    scratch = torch.empty(int(vm._native.load().vm_fused_scratch_bytes(n, h, w)) + 512, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef, plan=plan, out=out, scratch=scratch, status=status)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef, plan=plan, out=out, scratch=scratch, status=status)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    # new values in the captured buffers: frames rolled by one, result must follow
    fg_d.copy_(torch.roll(fg_d, 1, dims=0)); fb_d.copy_(torch.roll(fb_d, 1, dims=0)); ff_d.copy_(torch.roll(ff_d, 1, dims=0))
    bg_d.copy_(torch.roll(bg_d, 1, dims=0)); ctrl.copy_(torch.roll(ctrl, 1, dims=0)); coef.copy_(torch.roll(coef, 1, dims=0))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, torch.roll(ref, 1, dims=0))

    def timed(fn, iters=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    t_eager = timed(lambda: P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef, plan=plan, out=out, scratch=scratch, status=status))
    t_graph = timed(g.replay)
    print(f"C4 {n} x {h}x{w}: eager {t_eager * 1e3:.1f} us per call, graph replay {t_graph * 1e3:.1f} us")
    assert t_graph <= 1.25 * t_eager

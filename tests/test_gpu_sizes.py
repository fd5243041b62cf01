"""GPU parity at the sizes BASELINE.json names, through the same public entry points the bench times:

* config 1 at native 500 x 1200 (reference-generated fixture, tests/golden/make_c1_full_golden.py),
* config 2: one 1080p frame of the fused flow warp + mask against the oracle,
* config 3: 512 x 512 x 8 with 16 control points against the oracle,
* config 4: one deformed-grid 4K frame against the oracle (the pinv truncation regime of SURVEY 8a-5 differs
  at 4K, and it is where the int index arithmetic is largest),
* the host-buffer API the end-to-end number is measured through (`flow_tps_composite_host`), incl. a ragged
  last chunk and a solver pool,
* background cycling `bg[f % n_bg]` with 1 < n_bg < n,
* the `dropin/` shims: the reference's own loader.video_load_crop running on the product's flow / reader.

Bars as in test_gpu_parity.py: uint8 and masks bit-exact, floats |got-ref| <= 1e-5*|ref| + atol; TPS knife-edge
flips are counted and bounded.
"""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

import vm_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5


def close(got, ref, atol):
    return np.allclose(got, ref, rtol=RTOL, atol=atol)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def flips(got, ref):
    return int(np.count_nonzero(~np.isclose(got, ref, rtol=RTOL, atol=1e-5)))


# ------------------------------------------------------------------------------------ config 1

def test_config1_native_resolution(vm, tmp_path, capsys):
    """in0063 warped onto in0062 (DIS stand-in flows, quantised to 1/16 px), consistency mask, composite onto
    sea.jpg on the whole 500 x 1200 frame: drop-in functions bit-equal to the unmodified reference (digests),
    fused C2 entry point equal too."""
    import cv2
    with np.load(os.path.join(ROOT, "tests", "golden", "c1_full_golden.npz")) as z:
        c = {k: z[k] for k in z.files}
    fb = c["backward_q16"].astype(np.float32) / 16.0
    ff = c["forward_q16"].astype(np.float32) / 16.0
    png = tmp_path / "in0063.png"
    png.write_bytes(c["in0063_png"].tobytes())
    alpha, bgr = vm.reader.read_fg_img(str(png))                        # uint16 branch of reader.py:13-15 on the real file
    bgr = np.ascontiguousarray(bgr)
    assert alpha.dtype == np.float64 and np.array_equal(sha(alpha), c["sha_alpha63"]) and np.array_equal(sha(bgr), c["sha_fg63"])
    bg = cv2.resize(cv2.imdecode(c["sea_jpg"], cv2.IMREAD_COLOR), dsize=(1200, 500), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(sha(bg), c["sha_bg"])
    wa = vm.flow.warp_img(alpha, fb)
    assert wa.dtype == np.float64 and np.array_equal(sha(wa), c["sha_warp_alpha"])
    wb = vm.flow.warp_bgr(bgr, fb)
    assert np.array_equal(sha(wb), c["sha_warp_bgr"])
    before = wa.copy()
    ca = vm.flow.correct_alpha(fb, ff, wa)
    assert ca is wa and int((ca != before).sum()) == int(c["n_masked"][0])
    assert np.array_equal(sha(ca), c["sha_corrected"])
    assert str(tuple(ff.shape)) in capsys.readouterr().out
    cmp_ = vm.reader.create_composite_image(wb, bg, ca)
    assert cmp_.dtype == np.float64
    assert close(cmp_[::3, ::3], c["composite_sub"].astype(np.float64), 1e-5)
    assert close(cmp_, O.create_composite_image(wb, bg, ca), 1e-5)
    # fused C2: one pass over the BGRA frame
    fg4 = np.concatenate((bgr, np.rint(alpha * 255.).astype(np.uint8)[..., None]), axis=2)
    assert np.array_equal(fg4[..., 3] / 255., alpha)
    ob, oa, st = vm.pipeline.flow_warp_mask(dev(fg4[None]), dev(fb[None]), dev(ff[None]))
    assert np.array_equal(sha(ob[0].cpu().numpy()), c["sha_warp_bgr"])
    assert close(oa[0].cpu().numpy(), ca, 1e-6)
    assert int(st[0]) == 0 and int(st[1]) == 0


# ------------------------------------------------------------------------------------ config 2

def test_config2_1080p_frame_vs_oracle(vm):
    h, w = 1080, 1920
    frame = O.synth_frame(2001, h, w)
    fb, ff = O.synth_flows(2001, h, w)
    bgr, alpha, st = vm.pipeline.flow_warp_mask(dev(frame[None]), dev(fb[None]), dev(ff[None]))
    rb, ra = O.pipeline_c2(frame, fb, ff)
    assert np.array_equal(bgr[0].cpu().numpy(), rb)
    assert close(alpha[0].cpu().numpy(), ra, 1e-6)
    assert 0.02 < float((ra == 0).mean()) and int(st[0]) == 0 and int(st[1]) == 0


# ------------------------------------------------------------------------------------ config 3

def test_config3_512_batch_16_control_points_vs_oracle(vm):
    h, w, n = 512, 512, 8
    P = vm.pipeline
    frames = np.stack([O.synth_frame(3000 + k, h, w) for k in range(n)])
    grids = [O.synth_grids(3000 + k, h, w, 4) for k in range(n)]
    bgs = np.stack([O.synth_background(k % 2, h, w) for k in range(2)])         # grass / sea stand-ins, cycled
    ctrl, coef = P.solve_grids(grids)
    out, st = P.tps_composite(dev(frames), dev(bgs), ctrl, coef)
    out = out.cpu().numpy()
    total = 0
    for k in range(n):
        rc, ra = O.pipeline_c3(frames[k], grids[k], bgs[k % 2])
        assert close(out[k][..., 3], ra, 1e-6)
        total += flips(out[k][..., :3], rc)
    assert total <= n, f"{total} composite samples beyond tolerance in {n} frames"
    assert int(st[5]) == 0


# ------------------------------------------------------------------------------------ config 4

def test_config4_4k_deformed_frame_vs_oracle(vm):
    h, w = 2160, 3840
    P = vm.pipeline
    frame = O.synth_frame(4001, h, w)
    fb, ff = O.synth_flows(4001, h, w)
    grids = O.synth_grids(4001, h, w, 5)
    bg = O.synth_background(4, h, w)
    ctrl, coef = P.solve_grids([grids])
    out, st = P.flow_tps_composite(dev(frame[None]), dev(fb[None]), dev(ff[None]), dev(bg[None]), ctrl, coef)
    got = out[0].cpu().numpy()
    rc, ra = O.pipeline_c4(frame, fb, ff, grids, bg)
    assert close(got[..., 3], ra, 1e-6)
    nf = flips(got[..., :3], rc)
    assert nf <= 8, f"{nf} composite samples differ beyond tolerance at 4K"
    assert int(st[0]) == 0 and int(st[1]) == 0 and int(st[5]) == 0


# ------------------------------------------------------------------------------------ host API, n_bg cycling

def _small_case(h, w, n, seed):
    frames = np.stack([O.synth_frame(seed + k, h, w) for k in range(n)])
    flows = [O.synth_flows(seed + k, h, w) for k in range(n)]
    grids = [O.synth_grids(seed + k, h, w, 5) for k in range(n)]
    return frames, np.stack([f[0] for f in flows]), np.stack([f[1] for f in flows]), grids


@pytest.mark.parametrize("use_pool", [False, True])
def test_host_clip_api_vs_oracle(vm, use_pool):
    """flow_tps_composite_host (what bench.py's e2e times): 7 frames in chunks of 3 -> ragged last chunk; pinned and
    pageable inputs; with and without the solver pool; equal to the device-resident path bit for bit."""
    h, w, n = 120, 168, 7
    P = vm.pipeline
    frames, fb, ff, grids = _small_case(h, w, n, 700)
    bgs = np.stack([O.synth_background(k, h, w) for k in range(n)])
    pool = P.SolverPool(2) if use_pool else None
    try:
        res = P.flow_tps_composite_host(frames, fb, ff, bgs, grids, chunk=3, pool=pool)
        pinned = [torch.from_numpy(a).pin_memory() for a in (frames, fb, ff, bgs)]
        res2 = P.flow_tps_composite_host(*pinned, grids, chunk=3, pool=pool)
    finally:
        if pool is not None:
            pool.close()
    got = res.numpy()
    assert got.shape == (n, h, w, 4) and got.dtype == np.float32
    assert torch.equal(res, res2)
    ctrl, coef = P.solve_grids(grids)
    dres, _ = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    assert torch.equal(dres.cpu(), res)
    for k in range(n):
        rc, ra = O.pipeline_c4(frames[k], fb[k], ff[k], grids[k], bgs[k])
        assert close(got[k][..., 3], ra, 1e-6)
        assert flips(got[k][..., :3], rc) <= 1


def test_background_cycling(vm):
    """frame f composites onto bg[f % n_bg] for 1 < n_bg < n (C3 and C4)."""
    h, w, n, n_bg = 96, 128, 5, 3
    P = vm.pipeline
    frames, fb, ff, grids = _small_case(h, w, n, 810)
    bgs = np.stack([O.synth_background(20 + k, h, w) for k in range(n_bg)])
    ctrl, coef = P.solve_grids(grids)
    out4, _ = P.flow_tps_composite(dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    out3, _ = P.tps_composite(dev(frames), dev(bgs), ctrl, coef)
    out4, out3 = out4.cpu().numpy(), out3.cpu().numpy()
    for k in range(n):
        rc4, ra4 = O.pipeline_c4(frames[k], fb[k], ff[k], grids[k], bgs[k % n_bg])
        rc3, ra3 = O.pipeline_c3(frames[k], grids[k], bgs[k % n_bg])
        assert close(out4[k][..., 3], ra4, 1e-6) and flips(out4[k][..., :3], rc4) <= 1
        assert close(out3[k][..., 3], ra3, 1e-6) and flips(out3[k][..., :3], rc3) <= 1


# ------------------------------------------------------------------------------------ drop-in shims

def test_dropin_shims_resolve_to_product(vm):
    """`video-matting_b200/dropin` on sys.path: the reference's bare module names import the product modules."""
    d = os.path.join(ROOT, "video-matting_b200", "dropin")
    names = ("flow", "reader", "tps", "augmentation", "loader", "data")
    saved = {n: sys.modules.pop(n, None) for n in names}
    sys.path.insert(0, d)
    try:
        import importlib
        for n in names:
            m = importlib.import_module(n)
            assert m is getattr(vm, n), f"import {n} must resolve to video_matting_b200.{n}"
    finally:
        sys.path.remove(d)
        for n, m in saved.items():
            sys.modules.pop(n, None)
            if m is not None:
                sys.modules[n] = m


def test_install_dropin_and_reference_loader(vm, tmp_path):
    """The reference's OWN loader.video_load_crop (loader.py:285-330, from baseline/_ref) running on top of the
    product's `flow` and `reader` (install_dropin) reproduces the outputs the unmodified reference stack produced
    for the committed loader fixture."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refshim
    if refshim.reference_dir() is None:
        pytest.skip("baseline/_ref not populated (python baseline/install_ref.py in the build container)")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_loader_golden as MG
    names = ("flow", "reader", "tps", "augmentation", "loader", "data")
    saved = {n: sys.modules.get(n) for n in names}
    try:
        vm.install_dropin()
        assert sys.modules["flow"] is vm.flow and sys.modules["reader"] is vm.reader
    finally:
        for n, m in saved.items():
            sys.modules.pop(n, None)
            if m is not None:
                sys.modules[n] = m
    ref_loader = refshim.load(("loader",), overrides={"flow": vm.flow, "reader": vm.reader})["loader"]
    assert ref_loader.flow is vm.flow and ref_loader.reader is vm.reader
    with np.load(os.path.join(ROOT, "tests", "golden", "loader_golden.npz")) as z:
        lg = {k: z[k] for k in z.files}
    lat = int(lg["lattice"])
    for tag in [c[0] for c in MG.CASES][:3]:
        d = tmp_path / tag
        os.makedirs(str(d), exist_ok=True)
        p = MG.write_inputs(str(d), {k: lg[f"{tag}_file_{k}"] for k in ("fg", "prev", "bg", "flo", "hw")})
        w_in, h_in, seed = (int(v) for v in lg[f"{tag}_meta"])
        np.random.seed(seed)
        got = ref_loader.video_load_crop((p["fg"], p["bg"], p["prev"], p["flo"]), (w_in, h_in))
        for name, g in zip(("cmp", "bg", "label", "warped", "fg"), got):
            ref = lg[f"{tag}_video_{name}"]
            assert np.allclose(np.asarray(g)[::lat, ::lat], ref, rtol=1e-5, atol=1e-9), f"{tag} {name}"


# ------------------------------------------------------------------------------------ tps.warp_images(order=0)

def test_tps_warp_images_order0_vs_oracle(vm):
    """interpolation_order=0 (tps.py:22 "if 0 then use nearest-neighbor"): scipy's floor(t + 1/2) rule, uint8 and
    float64 images, against the oracle (itself checked against scipy and the reference on the CPU)."""
    rng = np.random.default_rng(5)
    h, w = 61, 83
    grid, dgrid = O.synth_grids(9, h, w, 4)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    imgs = [img, img / 255.]
    for order in (0, 1):
        got = vm.tps.warp_images(grid, dgrid, imgs, (0, 0, h, w), interpolation_order=order)
        ref = O.tps_warp_images(grid, dgrid, imgs, (0, 0, h, w), interpolation_order=order)
        assert got[0].dtype == np.uint8 and got[1].dtype == np.float64 and got[0].shape == (h + 1, w + 1)
        assert np.count_nonzero(got[0] != ref[0]) <= 1                   # knife-edge budget of the float64 transform
        assert np.count_nonzero(~np.isclose(got[1], ref[1], rtol=1e-5, atol=1e-6)) <= 1
    with pytest.raises(NotImplementedError):
        vm.tps.warp_images(grid, dgrid, imgs, (0, 0, h, w), interpolation_order=3)
    # explicit-transform entry point, knife-edge coordinates
    t0 = rng.uniform(-2, h + 1, (40, 50)); t1 = rng.uniform(-2, w + 1, (40, 50))
    t0[0, :6] = [-1e-9, 0.0, 0.5, 1.5, h - 1.0, h - 1 + 1e-7]; t1[0, :6] = 2.5
    got = vm.pipeline.map_coordinates(dev(img), dev(t0), dev(t1), order=0).cpu().numpy()
    assert np.array_equal(got, O.map_coordinates_nearest(img, t0, t1))


# ------------------------------------------------------------------------------------ row f2: exact resize / HSV2BGR

def test_resize_u8_bit_exact_vs_cv2_and_oracle(vm):
    """cv2.resize(uint8, INTER_LINEAR) on the device: equal to this host's cv2 and to the oracle, up- and down-scaling,
    1 / 3 / 4 channels, the 2x INTER_AREA switch, batches; config 1's background (sea.jpg -> 500 x 1200) by digest."""
    import cv2
    P = vm.pipeline
    rng = np.random.default_rng(0)
    shapes = [(300, 400, 500, 1200), (333, 517, 500, 1200), (375, 500, 1080, 1920), (1080, 1920, 512, 512), (100, 100, 37, 53),
              (64, 64, 128, 128), (128, 128, 64, 64), (128, 130, 64, 65), (5, 7, 50, 120), (1, 1, 4, 4), (2, 3, 1, 1), (17, 31, 16, 30)]
    for sh, sw, dh, dw in shapes:
        for cn in (1, 3, 4):
            src = rng.integers(0, 256, (sh, sw, cn) if cn > 1 else (sh, sw), dtype=np.uint8)
            got = P.resize_u8(dev(src), (dw, dh)).cpu().numpy()
            assert np.array_equal(got, O.resize_linear_u8(src, (dw, dh))), (sh, sw, dh, dw, cn)
            assert np.array_equal(got, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)), (sh, sw, dh, dw, cn)
    batch = rng.integers(0, 256, (3, 90, 70, 3), dtype=np.uint8)
    got = P.resize_u8(dev(batch), (111, 64)).cpu().numpy()
    for k in range(3):
        assert np.array_equal(got[k], cv2.resize(batch[k], (111, 64), interpolation=cv2.INTER_LINEAR))
    with np.load(os.path.join(ROOT, "tests", "golden", "c1_full_golden.npz")) as z:
        jpg, want = z["sea_jpg"], z["sha_bg"]
    bg = vm.reader.resize_background(cv2.imdecode(jpg, cv2.IMREAD_COLOR), 500, 1200)
    assert isinstance(bg, np.ndarray) and np.array_equal(sha(bg), want)


def test_change_illumination_bit_exact(vm, golden):
    """augmentation.change_illumination: BGR2HSV / HSV2BGR exactly as this host's cv2 computes them (truncating SIMD
    body, rounding row tail), against cv2 itself, the oracle, the reference fixture (made on an AVX2 host: 32 pixels
    per step) and - when baseline/_ref travels - the unmodified reference function."""
    import cv2
    vec = vm._native.hsv_vec()
    assert vec == O.probe_hsv_vec()
    rng = np.random.default_rng(12)
    for (h, w), (a, b, c) in (((37, 53), (1.03, 0.8, -0.02)), ((40, 64), (0.95, 1.3, 0.07)), ((21, 100), (1.05, 0.7, -0.07)),
                              ((8, 1), (1.0, 1.0, 0.0)), ((3, 257), (0.97, 0.9, 0.05))):
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        got = vm.augmentation.change_illumination(bgr, a, b, c)
        assert np.array_equal(got, O.change_illumination(bgr, a, b, c, vec)), (h, w)
        hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
        lut = vm.augmentation.illumination_lut(a, b, c)
        hsv[..., 1] = lut[hsv[..., 1]]; hsv[..., 2] = lut[hsv[..., 2]]
        assert np.array_equal(got, cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)), (h, w)
    lut = vm.augmentation.illumination_lut(1.03, 0.8, -0.02)
    got = vm.pipeline.illumination(dev(golden["ci_bgr"]), lut, hsv_vec=32).cpu().numpy()
    assert np.array_equal(got, golden["ci_out"])
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refshim
    if refshim.reference_dir() is not None:
        aug = refshim.load(("augmentation",))["augmentation"]
        bgr = rng.integers(0, 256, (45, 71, 3), dtype=np.uint8)
        assert np.array_equal(vm.augmentation.change_illumination(bgr, 1.02, 1.1, 0.03), aug.change_illumination(bgr, 1.02, 1.1, 0.03))


# ------------------------------------------------------------------------------------ config 5 at 1080p

def test_config5_augment_1080p_vs_oracle(vm):
    """augmentation.augment on one 1080p frame (config 5's size) against the oracle with the same np.random seed: the
    background (affine + illumination) is bit-exact, the alpha within tolerance, the foreground bit-exact up to TPS
    knife-edge pixels; and augment_clip with a solver pool gives the same frame."""
    h, w = 1080, 1920
    frame = O.synth_frame(5001, h, w)
    bgr, alpha = np.ascontiguousarray(frame[..., :3]), frame[..., 3] / 255.
    bg = O.synth_background(5, h, w)
    vec = vm._native.hsv_vec()
    np.random.seed(501)
    nfg, nbg, nal = vm.augmentation.augment(bgr, bg, alpha)
    state = np.random.get_state()[1][:4].copy()
    np.random.seed(501)
    rfg, rbg, ral = O.augment(bgr, bg, alpha, vec=vec)   # the oracle's default is the fixture host's SIMD width (32)
    assert np.array_equal(state, np.random.get_state()[1][:4]), "same np.random consumption"
    assert np.array_equal(nbg, rbg)
    assert np.allclose(nal, ral, rtol=RTOL, atol=1e-6)
    assert np.count_nonzero((nfg != rfg).any(axis=2)) <= 6, "foreground differs beyond TPS knife-edge pixels"
    pool = vm.pipeline.SolverPool(2)
    try:
        np.random.seed(501)
        cfg, cbg, cal = vm.augmentation.augment_clip(frame[None], bg[None], pool=pool)
    finally:
        pool.close()
    assert np.array_equal(cbg[0], nbg) and np.allclose(cal[0], nal, rtol=RTOL, atol=1e-6)
    assert np.count_nonzero((cfg[0] != nfg).any(axis=2)) <= 6


def test_upload_many_ring_reuse(vm):
    """_native.upload_many: every array arrives intact, with its dtype and shape, also when the pinned ring wraps
    around (more calls than slots) and when a later call needs a larger staging buffer."""
    N = vm._native
    dev = torch.device("cuda", 0)
    rng = np.random.RandomState(5)
    kept = []
    for k in range(20):
        arrs = [rng.rand(3 + k % 4, 25, 2), rng.rand(3 + k % 4, 28, 2), rng.randint(0, 256, size=(7 + k, 56)).astype(np.uint8),
                rng.randint(0, 256, size=(257,)).astype(np.uint8), rng.rand(5).astype(np.float32)]
        if k == 12:
            arrs.append(rng.rand(40000))                                   # 320 KB: the ring is rebuilt with larger slots
        outs = N.upload_many(arrs, dev)
        kept.append((arrs, outs))
    torch.cuda.synchronize()
    for arrs, outs in kept:
        assert len(arrs) == len(outs)
        for a, t in zip(arrs, outs):
            assert t.is_cuda and tuple(t.shape) == a.shape and t.cpu().numpy().dtype == a.dtype
            assert np.array_equal(t.cpu().numpy(), a)

"""GPU parity tests of the batch loader: video-matting_b200/loader.py (through the
C ABI entry vm_loader_batch) against the golden vectors of the unmodified reference loader.py and
against the NumPy oracle on the same files and the same np.random seeds.

Tolerance (floats, BASELINE north_star): |got - ref| <= 1e-5 * |ref| + 1e-9 on 0..255 data."""
import os
import sys

import numpy as np
import pytest

import vm_loader_oracle as LO
import vm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_loader_golden as MG  # noqa: E402

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def lgold():
    with np.load(os.path.join(ROOT, "tests", "golden", "loader_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def files_of(lgold, tag, d):
    os.makedirs(str(d), exist_ok=True)
    return MG.write_inputs(str(d), {k: lgold[f"{tag}_file_{k}"] for k in ("fg", "prev", "bg", "flo", "hw")})


def close(got, ref, what, rtol=1e-5, atol=1e-9):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    err = np.abs(got - ref)
    assert np.all(err <= rtol * np.abs(ref) + atol), f"{what}: max err {err.max():.3e}"


@pytest.mark.parametrize("tag", [c[0] for c in MG.CASES])
def test_load_crop_variants_match_reference_and_oracle(vm, lgold, tag, tmp_path):
    L = vm.loader
    lat = int(lgold["lattice"])
    p = files_of(lgold, tag, tmp_path)
    w_in, h_in, seed = (int(v) for v in lgold[f"{tag}_meta"])
    size = (w_in, h_in)
    fg = cv2.imread(p["fg"], cv2.IMREAD_UNCHANGED)
    prev = cv2.imread(p["prev"], cv2.IMREAD_UNCHANGED)
    bg = cv2.imread(p["bg"])
    tri = cv2.imread(p["tri"], 0)
    flo, _ = O.parse_flo(open(p["flo"], "rb").read())

    np.random.seed(seed)
    got = L.video_load_crop((p["fg"], p["bg"], p["prev"], p["flo"]), size)
    ref = LO.video_sample(fg, bg, prev, flo, size, np.random.RandomState(seed))
    for name, g, r in zip(("cmp", "bg", "label", "warped", "fg"), got, ref):
        assert g.dtype == np.float64
        close(g, r, f"{tag} video {name} vs oracle")
        close(g[::lat, ::lat], lgold[f"{tag}_video_{name}"], f"{tag} video {name} vs reference")
    assert np.array_equal(got[4], ref[4]), "resized foreground is bit-equal to the oracle"

    np.random.seed(seed + 100)
    got = L.simple_load_crop((p["fg"], p["tri"], p["bg"]), size)
    for name, g in zip(("cmp", "bg", "label", "fg"), got):
        close(g[::lat, ::lat], lgold[f"{tag}_simple_{name}"], f"{tag} simple {name} vs reference")

    np.random.seed(seed + 200)
    got = L.load_and_crop((p["fg"], p["tri"], p["bg"]), size)
    ref = LO.trimap_sample(fg, tri, bg, size, np.random.RandomState(seed + 200))
    for name, g, r in zip(("inp", "label", "fg"), got, ref):
        close(g, r, f"{tag} trimap {name} vs oracle")
        close(g[::lat, ::lat], lgold[f"{tag}_trimap_{name}"], f"{tag} trimap {name} vs reference")


def batch_entries(lgold, tmp_path):
    ev, es = [], []
    for tag in lgold["batch_tags"]:
        p = files_of(lgold, str(tag), tmp_path / str(tag))
        ev.append((p["fg"], p["bg"], p["prev"], p["flo"]))
        es.append((p["fg"], p["tri"], p["bg"]))
    return ev, es


def test_batch_entry_points_match_reference(vm, lgold, tmp_path):
    L = vm.loader
    lat = 2 * int(lgold["lattice"])
    ev, es = batch_entries(lgold, tmp_path)
    np.random.seed(4242)
    for name, g in zip(("cmp", "bg", "label", "warped", "fg"), L.video_batch(ev, (320, 320))):
        close(g[:, ::lat, ::lat], lgold[f"batch_video_{name}"], f"video_batch {name}")
    np.random.seed(4243)
    for name, g in zip(("cmp", "bg", "label", "fg"), L.simple_batch(es, (320, 320))):
        close(g[:, ::lat, ::lat], lgold[f"batch_simple_{name}"], f"simple_batch {name}")
    np.random.seed(4244)                                        # rd_mirror=True: per-sample flip draws
    for name, g in zip(("inp", "label", "fg"), L.get_batch(es, (320, 320), False, True)):
        close(g[:, ::lat, ::lat], lgold[f"batch_get_{name}"], f"get_batch {name}")
    with pytest.raises(ValueError):
        L.video_batch(ev, (320, 160))
    with pytest.raises(ValueError):
        L.get_batch(es, (320, 320), True, False)


def test_device_outputs_and_float32(vm, lgold, tmp_path):
    import torch
    L = vm.loader
    ev, _ = batch_entries(lgold, tmp_path)
    np.random.seed(7)
    ref = L.video_batch(ev, (320, 320))
    np.random.seed(7)
    dev = L.video_batch(ev, (320, 320), device=True)
    np.random.seed(7)
    f32 = L.video_batch(ev, (320, 320), device=True, dtype=np.float32)
    torch.cuda.synchronize()
    for r, d, f in zip(ref, dev, f32):
        assert d.is_cuda and d.dtype == torch.float64 and f.dtype == torch.float32
        assert np.array_equal(d.cpu().numpy(), r)
        assert np.array_equal(f.cpu().numpy(), r.astype(np.float32))
    assert L.video_batch([], (320, 320))[0].shape == (0, 320, 320, 3)


def test_same_rng_state_after_call_as_the_oracle(vm, lgold, tmp_path):
    L = vm.loader
    p = files_of(lgold, "v1", tmp_path)
    fg = cv2.imread(p["fg"], cv2.IMREAD_UNCHANGED)
    bg = cv2.imread(p["bg"])
    np.random.seed(99)
    L.simple_load_crop((p["fg"], p["tri"], p["bg"]), (320, 320))
    after = np.random.uniform()
    r = np.random.RandomState(99)
    LO.simple_sample(fg, bg, (320, 320), r)
    assert after == r.uniform()


def test_psnr_and_helpers(vm, lgold, tmp_path):
    L = vm.loader
    assert abs(L.psnr(lgold["psnr_a"], lgold["psnr_b"]) - float(lgold["psnr_3"])) < 1e-9
    assert abs(L.psnr(lgold["psnr_a"][:, :, 0], lgold["psnr_b"][:, :, 0]) - float(lgold["psnr_1"])) < 1e-9
    np.random.seed(3)
    n = L.add_noise(lgold["psnr_a"][:, :, 0], var=0.05)
    assert n.shape == (40, 50, 1) and n.min() >= 0 and n.max() <= 1
    lst = tmp_path / "list.txt"
    lst.write_text("a.png t.png b.jpg\nc.png u.png d.jpg\n")
    fl = L.get_file_list("/root", str(lst))
    assert fl == [["/root/a.png", "/root/t.png", "/root/b.jpg"], ["/root/c.png", "/root/u.png", "/root/d.jpg"]]
    assert not L.epoch_is_over(fl, 2) and L.get_batch_list(fl, 1) == [["/root/c.png", "/root/u.png", "/root/d.jpg"]]
    assert L.epoch_is_over(fl, 2)
    with pytest.raises(AttributeError):
        L.simple_load_crop((files_of(lgold, "v0", tmp_path)["fg"], "x", "/nonexistent/bg.png"), (320, 320))


def test_trimap_from_matte_matches_reference_and_oracle(vm, lgold):
    import torch
    D = vm.data
    for tag in ("t0", "t1"):
        m8 = lgold[f"{tag}_matte_u8"]
        got = D.trimap_from_matte(m8 / 255.)
        assert got.dtype == np.uint8 and np.array_equal(got, lgold[f"{tag}_trimap"])
        assert np.array_equal(D.trimap_from_matte(m8), lgold[f"{tag}_trimap"])
    with pytest.raises(AssertionError):
        D.trimap_from_matte(np.zeros((4, 4), np.float32))
    # large stack against the closed-form oracle (odd sizes, tile borders)
    a8 = np.stack([O.synth_frame(s, 203, 331)[..., 3] for s in range(3)])
    got = D.trimap_from_matte(torch.from_numpy(a8).cuda()).cpu().numpy()
    for k in range(3):
        assert np.array_equal(got[k], LO.trimap_from_matte(a8[k] / 255.))
    edge = np.zeros((9, 40))
    edge[4, :] = 0.5; edge[:, 0] = 1.; edge[:, 39] = 1.; edge[0, :] = 1.; edge[8, 5:9] = 0.3
    assert np.array_equal(D.trimap_from_matte(edge), LO.trimap_from_matte_loop(edge))


def test_fg_from_u16_exhaustive(vm):
    import torch
    N = vm._native
    v = np.arange(65536 + 5, dtype=np.uint32).astype(np.uint16)          # 65541 elements: vector body + tail
    src = torch.from_numpy(v.view(np.int16)).cuda().view(torch.uint16)
    dst = torch.empty(v.size, dtype=torch.uint8, device="cuda")
    N.check(N.load().vm_fg_from_u16(N.ptr(src), v.size, N.ptr(dst), N.stream_ptr()))
    assert np.array_equal(dst.cpu().numpy(), O.fg_from_uint16(v))


@pytest.mark.parametrize("wide", [False, True])
def test_load_clip_matches_per_file_readers(vm, tmp_path, wide, capsys):
    R = vm.reader
    h, w, n = 45, 70, 5
    fgp, fbp, ffp, bgp = [], [], [], []
    rng = np.random.default_rng(3)
    for k in range(n):
        fr = O.synth_frame(k, h, w)
        img = (fr.astype(np.uint16) * 257 + rng.integers(0, 257, fr.shape).astype(np.uint16)) if wide else fr
        if wide:
            img.reshape(-1)[:8] = [0, 1, 254, 255, 256, 511, 65534, 65535]
        fgp.append(str(tmp_path / f"fg{k}.png"))
        assert cv2.imwrite(fgp[-1], img)
        fb, ff = O.synth_flows(k, h, w)
        for lst, name, f in ((fbp, "b", fb), (ffp, "f", ff)):
            lst.append(str(tmp_path / f"{name}{k}.flo"))
            O.write_flo(lst[-1], f)
    for k, shape in enumerate(((h, w), (33, 91))):
        bgp.append(str(tmp_path / f"bg{k}.png"))
        assert cv2.imwrite(bgp[-1], O.synth_background(k, *shape))
    clip = R.load_clip(fgp, fbp, ffp, bgp, chunk=2, threads=3)
    assert clip["fg"].shape == (n, h, w, 4) and clip["fg"].is_cuda
    for k in range(n):
        alpha, bgr = R.read_fg_img(fgp[k])
        got = clip["fg"][k].cpu().numpy()
        assert np.array_equal(got[..., :3], bgr) and np.array_equal(got[..., 3] / 255., alpha)
        assert np.array_equal(clip["backward"][k].cpu().numpy(), R.read_flow(fbp[k]))
        assert np.array_equal(clip["forward"][k].cpu().numpy(), R.read_flow(ffp[k]))
    assert np.array_equal(clip["bg"][0].cpu().numpy(), cv2.imread(bgp[0]))
    assert np.array_equal(clip["bg"][1].cpu().numpy(), cv2.resize(cv2.imread(bgp[1]), (w, h), interpolation=cv2.INTER_LINEAR))
    # reference conventions: short .flo -> ValueError; bad magic -> message, parsing continues
    short = str(tmp_path / "short.flo")
    open(short, "wb").write(open(fbp[0], "rb").read()[:-8])
    with pytest.raises(ValueError):
        R.load_clip(fgp[:1], [short])
    bad = bytearray(open(fbp[0], "rb").read())
    bad[0] ^= 0xFF
    open(str(tmp_path / "bad.flo"), "wb").write(bytes(bad))
    capsys.readouterr()
    c2 = R.load_clip(fgp[:1], [str(tmp_path / "bad.flo")])
    assert "ERROR: invalid key" in capsys.readouterr().out
    assert np.array_equal(c2["backward"][0].cpu().numpy(), R.read_flow(fbp[0]))


def test_device_batch_equals_file_batch(vm, lgold, tmp_path):
    import torch
    L = vm.loader
    ev, es = batch_entries(lgold, tmp_path)
    fg = [torch.from_numpy(cv2.imread(e[0], cv2.IMREAD_UNCHANGED)).cuda() for e in ev]
    bg = [torch.from_numpy(cv2.imread(e[1])).cuda() for e in ev]
    prev = [torch.from_numpy(cv2.imread(e[2], cv2.IMREAD_UNCHANGED)).cuda() for e in ev]
    flo = [torch.from_numpy(O.parse_flo(open(e[3], "rb").read())[0].copy()).cuda() for e in ev]
    np.random.seed(11)
    ref = L.video_batch(ev, (320, 320))
    np.random.seed(11)
    got = L.device_batch(fg, bg, (320, 320), prev=prev, flow=flo)
    for name, r in zip(("cmp", "bg", "label", "warped", "fg"), ref):
        assert np.array_equal(got[name].cpu().numpy(), r), name
    np.random.seed(12)
    ref = L.get_batch(es, (320, 320), False, True)
    np.random.seed(12)
    got = L.device_batch(fg, bg, (320, 320), mirror=True)
    assert got["warped"] is None
    assert np.array_equal(torch.cat((got["cmp"], got["bg"]), 3).cpu().numpy(), ref[0])
    assert np.array_equal(got["label"].cpu().numpy(), ref[1]) and np.array_equal(got["fg"].cpu().numpy(), ref[2])


def test_device_batch_random_geometries_vs_oracle(vm):
    """40 random foreground / background / output sizes (padding on either axis, every crop type, area and linear
    resize, mirror) through vm_loader_batch against the oracle with the same np.random seed."""
    import torch
    L = vm.loader
    rng = np.random.default_rng(77)
    for trial in range(40):
        fh, fw = int(rng.integers(2, 700)), int(rng.integers(2, 700))
        bh, bw = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        size = (int(rng.choice([32, 48, 160, 240, 320])), int(rng.choice([32, 48, 160, 240, 320])))
        fg = rng.integers(0, 256, size=(fh, fw, 4), dtype=np.uint8)
        prev = rng.integers(0, 256, size=(fh, fw, 4), dtype=np.uint8)
        bg = rng.integers(0, 256, size=(bh, bw, 3), dtype=np.uint8)
        flo = (rng.normal(0, 6, size=(fh, fw, 2))).astype(np.float32)
        video = trial % 2 == 0
        np.random.seed(trial)
        got = L.device_batch([torch.from_numpy(fg).cuda()], [torch.from_numpy(bg).cuda()], size,
                             prev=[torch.from_numpy(prev).cuda()] if video else None,
                             flow=[torch.from_numpy(flo).cuda()] if video else None)
        r = np.random.RandomState(trial)
        if video:
            ref = LO.video_sample(fg, bg, prev, flo, size, r)
            names = ("cmp", "bg", "label", "warped", "fg")
        else:
            ref = LO.simple_sample(fg, bg, size, r)
            names = ("cmp", "bg", "label", "fg")
        for name, rv in zip(names, ref):
            close(got[name][0].cpu().numpy(), rv, f"trial {trial} {name} ({fh}x{fw}, bg {bh}x{bw}, out {size})")


def test_generate_trimaps_and_create_list(vm, tmp_path):
    D = vm.data
    src, dst = tmp_path / "fg", tmp_path / "trimap"
    os.makedirs(src); os.makedirs(dst)
    frames = {"a.png": O.synth_frame(1, 40, 56), "b.png": O.synth_frame(2, 40, 56), "c.png": O.synth_frame(3, 33, 47)}
    for name, fr in frames.items():
        assert cv2.imwrite(str(src / name), fr)
    assert cv2.imwrite(str(dst / "b.png"), np.zeros((40, 56), np.uint8))          # existing targets are skipped (data.py:80-81)
    D.generate_trimaps(str(src), str(dst))
    for name, fr in frames.items():
        got = cv2.imread(str(dst / name), cv2.IMREAD_UNCHANGED)
        if name == "b.png":
            assert not got.any()
        else:
            assert np.array_equal(got, LO.trimap_from_matte(fr[..., 3] / 255.))
    # create_list: 100 random backgrounds per foreground, one np.random.randint(size=100) call each (data.py:87-107)
    root = tmp_path / "root"
    for d in (root / "fg" / "S", root / "trimap" / "S", root / "VOC_bg"):
        os.makedirs(d)
    for n in ("x.png", "y.png"):
        (root / "fg" / "S" / n).write_bytes(b"0"); (root / "trimap" / "S" / n).write_bytes(b"0")
    for n in ("v0.jpg", "v1.jpg", "v2.jpg"):
        (root / "VOC_bg" / n).write_bytes(b"0")
    np.random.seed(5)
    D.create_list(str(root), str(tmp_path / "list.txt"), "S")
    lines = (tmp_path / "list.txt").read_text().splitlines()
    assert len(lines) == 200
    np.random.seed(5)
    voc = os.listdir(root / "VOC_bg")
    fgs = os.listdir(root / "fg" / "S")
    ids = np.random.randint(0, len(voc), size=100)
    assert lines[0] == "{} {} {}".format(os.path.join("fg", "S", fgs[0]), os.path.join("trimap", "S", fgs[0]),
                                         os.path.join("VOC_bg", voc[ids[0]]))
    entries = vm.loader.get_file_list(str(root), str(tmp_path / "list.txt"))
    assert len(entries) == 200 and entries[0][0] == os.path.join(str(root), "fg", "S", fgs[0])


def test_load_clip_odd_frame_size_and_chunk(vm, tmp_path):
    """16-bit PNGs with an odd pixel count and an odd chunk: the converted frames land at 4-byte aligned offsets of the
    clip (the vector path of vm_fg_from_u16 needs 8) - the element-wise path must take over."""
    R = vm.reader
    h, w, n = 45, 71, 4
    rng = np.random.default_rng(11)
    fgp = []
    for k in range(n):
        img = rng.integers(0, 65536, (h, w, 4)).astype(np.uint16)
        fgp.append(str(tmp_path / f"fg{k}.png"))
        assert cv2.imwrite(fgp[-1], img)
    clip = R.load_clip(fgp, chunk=1, threads=2)
    clip3 = R.load_clip(fgp, chunk=3, threads=2)
    for k in range(n):
        alpha, bgr = R.read_fg_img(fgp[k])
        for c in (clip, clip3):
            got = c["fg"][k].cpu().numpy()
            assert np.array_equal(got[..., :3], bgr) and np.array_equal(got[..., 3] / 255., alpha)


def test_flow_entry_rejects_misaligned_buffers(vm):
    """the vectorised C entry points return VM_ERR_ARG for pointers their 16-byte accesses cannot take (no device fault)"""
    import torch
    P, Nt = vm.pipeline, vm._native
    lib = Nt.load()
    n, h, w = 1, 8, 16
    fg = torch.zeros((n, h, w, 4), dtype=torch.uint8, device="cuda")
    flow = torch.zeros(n * h * w * 2 + 2, dtype=torch.float32, device="cuda")
    bgr = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    alpha = torch.empty((n, h, w), dtype=torch.float32, device="cuda")
    rc = lib.vm_flow_warp_mask_bgra(Nt.ptr(fg), flow.data_ptr() + 8, None, n, h, w, Nt.ptr(bgr), Nt.ptr(alpha), None, Nt.stream_ptr())
    assert rc == 1 and b"unaligned" in lib.vm_last_error_string()
    rc = lib.vm_flow_warp_mask_bgra(Nt.ptr(fg), flow.data_ptr(), None, n, h, w, Nt.ptr(bgr), Nt.ptr(alpha), None, Nt.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()

"""Generate tests/golden/loader_golden.npz by running the UNMODIFIED reference ``loader.py``.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_loader_golden.py

The reference loader takes file paths.  The fixture stores the synthetic inputs at 1/4
resolution; ``write_inputs`` expands them with integer arithmetic and writes the PNG / .flo
files, and the tests call the same function to recreate the files the reference read and hand
their paths to the oracle and to the CUDA path.  Outputs are float64; to keep the fixture small
only the lattice ``out[::LAT, ::LAT]`` of every output is stored.  ``np.random.seed(seed)``
precedes every reference call, so the random crop / padding decisions are part of the pin.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

LAT = 8

# (tag, fg size, bg size, input_size, seed)
CASES = (
    ("v0", (400, 520), (300, 410), (320, 320), 3),
    ("v1", (250, 700), (640, 640), (320, 320), 8),
    ("v2", (660, 650), (320, 320), (320, 320), 2),
    ("v3", (330, 340), (211, 173), (160, 160), 11),
    ("v4", (200, 180), (480, 640), (96, 128), 5),
    ("v5", (650, 700), (500, 500), (320, 320), 21),
)


F = 4   # inputs are stored at 1/F resolution and expanded with integer arithmetic only


def expand(small, h, w, keep_extremes=False):
    """(h, w, c) uint8 image from its stored (ceil(h/F), ceil(w/F), c) form: nearest-neighbour
    blocks plus a fixed per-pixel integer pattern in [-4, 4] (so neighbouring pixels differ
    and interpolation is really exercised).  Pure integer arithmetic: version independent."""
    small = np.asarray(small)
    big = np.repeat(np.repeat(small, F, axis=0), F, axis=1)[:h, :w].astype(np.int32)
    ii, jj, cc = np.meshgrid(np.arange(h), np.arange(w), np.arange(big.shape[2]), indexing="ij")
    pat = (ii * 7 + jj * 13 + cc * 5) % 9 - 4
    out = np.clip(big + pat, 0, 255)
    if keep_extremes:                                # alpha keeps its exact 0 / 255 plateaus
        out = np.where((big == 0) | (big == 255), big, out)
    return out.astype(np.uint8)


def expand_flow(small, h, w):
    """float32 (h, w, 2) flow from its stored int32 1/8-px form at 1/F resolution."""
    q = np.repeat(np.repeat(np.asarray(small), F, axis=0), F, axis=1)[:h, :w].astype(np.int64)
    ii, jj = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    q = q + ((ii * 3 + jj * 5) % 7 - 3)[:, :, None]
    return (q.astype(np.float32) / np.float32(8)).astype(np.float32)


def make_inputs(seed, fg_hw, bg_hw):
    """Stored (1/F resolution) form of one sample: fg BGRA, previous alpha, bg, flow."""
    import vm_oracle as O
    rng = np.random.default_rng(1000 + seed)
    sm = lambda n: (n + F - 1) // F
    h, w = sm(fg_hw[0]), sm(fg_hw[1])
    fg = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    fg[..., 3] = np.clip(128 + 384 * O._smooth_noise(rng, h, w, 8), 0, 255).astype(np.uint8)
    prev_a = np.clip(128 + 384 * O._smooth_noise(rng, h, w, 8), 0, 255).astype(np.uint8)
    bg = rng.integers(0, 256, size=(sm(bg_hw[0]), sm(bg_hw[1]), 3), dtype=np.uint8)
    flo = np.stack([O._smooth_noise(rng, h, w, 8), O._smooth_noise(rng, h, w, 8)], -1) * 6.0
    floq = np.rint(flo * 8).astype(np.int32)
    floq[1, 2] = (80000, -28)                                     # far out-of-frame vectors
    return {"fg": fg, "prev": prev_a, "bg": bg, "flo": floq,
            "hw": np.array([fg_hw[0], fg_hw[1], bg_hw[0], bg_hw[1]], dtype=np.int64)}


def write_inputs(d, files):
    """Write one sample's files: fg / previous fg RGBA PNGs, bg PNG, trimap PNG, .flo."""
    import cv2
    import vm_oracle as O
    h, w, bh, bw = (int(v) for v in files["hw"])
    fg = np.concatenate((expand(files["fg"][..., :3], h, w),
                         expand(files["fg"][..., 3:], h, w, keep_extremes=True)), axis=2)
    prev = fg.copy()
    prev[..., 3:] = expand(np.asarray(files["prev"])[..., None], h, w, keep_extremes=True)
    bg = expand(files["bg"], bh, bw)
    tri = np.where(fg[..., 3] == 0, 0, np.where(fg[..., 3] == 255, 255, 128)).astype(np.uint8)
    paths = {k: os.path.join(d, f"{k}.png") for k in ("fg", "prev", "bg", "tri")}
    for k, im in (("fg", fg), ("prev", prev), ("bg", bg), ("tri", tri)):
        assert cv2.imwrite(paths[k], im)
    paths["flo"] = os.path.join(d, "flow.flo")
    O.write_flo(paths["flo"], expand_flow(files["flo"], h, w))
    return paths


def main():
    from make_golden import import_reference, REF
    import_reference()
    sys.path.insert(0, REF)
    import loader                                                 # the unmodified reference
    out = {"lattice": np.int64(LAT)}
    with tempfile.TemporaryDirectory() as d:
        for tag, fg_hw, bg_hw, input_size, seed in CASES:
            files = make_inputs(seed, fg_hw, bg_hw)
            p = write_inputs(d, files)
            for k, v in files.items():
                out[f"{tag}_file_{k}"] = v
            out[f"{tag}_meta"] = np.array([input_size[0], input_size[1], seed], dtype=np.int64)
            np.random.seed(seed)
            res = loader.video_load_crop((p["fg"], p["bg"], p["prev"], p["flo"]), input_size)
            for name, a in zip(("cmp", "bg", "label", "warped", "fg"), res):
                out[f"{tag}_video_{name}"] = np.ascontiguousarray(a[::LAT, ::LAT])
            out[f"{tag}_video_shapes"] = np.array([a.shape + (0,) * (3 - a.ndim) for a in res], dtype=np.int64)
            np.random.seed(seed + 100)
            res = loader.simple_load_crop((p["fg"], p["tri"], p["bg"]), input_size)
            for name, a in zip(("cmp", "bg", "label", "fg"), res):
                out[f"{tag}_simple_{name}"] = np.ascontiguousarray(a[::LAT, ::LAT])
            np.random.seed(seed + 200)
            res = loader.load_and_crop((p["fg"], p["tri"], p["bg"]), input_size)
            for name, a in zip(("inp", "label", "fg"), res):
                out[f"{tag}_trimap_{name}"] = np.ascontiguousarray(a[::LAT, ::LAT])
        # batch entry points on square input sizes (the reference's batch arrays are
        # (B, input_size[0], input_size[1], .), loader.py:335-339)
        tagsq = [c for c in CASES if c[3][0] == c[3][1] == 320][:3]
        entries_v, entries_s = [], []
        for tag, fg_hw, bg_hw, input_size, seed in tagsq:
            sub = os.path.join(d, tag)
            os.makedirs(sub)
            p = write_inputs(sub, {k: out[f"{tag}_file_{k}"] for k in ("fg", "prev", "bg", "flo", "hw")})
            entries_v.append((p["fg"], p["bg"], p["prev"], p["flo"]))
            entries_s.append((p["fg"], p["tri"], p["bg"]))
        out["batch_tags"] = np.array([c[0] for c in tagsq])
        np.random.seed(4242)
        for name, a in zip(("cmp", "bg", "label", "warped", "fg"), loader.video_batch(entries_v, (320, 320))):
            out[f"batch_video_{name}"] = np.ascontiguousarray(a[:, ::LAT * 2, ::LAT * 2])
        np.random.seed(4243)
        for name, a in zip(("cmp", "bg", "label", "fg"), loader.simple_batch(entries_s, (320, 320))):
            out[f"batch_simple_{name}"] = np.ascontiguousarray(a[:, ::LAT * 2, ::LAT * 2])
        np.random.seed(4244)
        for name, a in zip(("inp", "label", "fg"), loader.get_batch(entries_s, (320, 320), False, True)):
            out[f"batch_get_{name}"] = np.ascontiguousarray(a[:, ::LAT * 2, ::LAT * 2])
        # data.trimap_from_matte (reference data.py:37-67, pure-Python loop: small images only)
        import data
        for tag, (h, w), seed in (("t0", (37, 53), 1), ("t1", (24, 24), 2)):
            import vm_oracle as O
            a8 = np.clip(128 + 384 * O._smooth_noise(np.random.default_rng(seed), h, w, 6), 0, 255).astype(np.uint8)
            out[f"{tag}_matte_u8"] = a8
            out[f"{tag}_trimap"] = data.trimap_from_matte(a8 / 255.)
        # psnr
        rng = np.random.default_rng(9)
        a, b = rng.uniform(0, 1, (40, 50, 3)), rng.uniform(0, 1, (40, 50, 3))
        out.update({"psnr_a": a, "psnr_b": b, "psnr_3": np.float64(loader.psnr(a, b)),
                    "psnr_1": np.float64(loader.psnr(a[:, :, 0], b[:, :, 0]))})
    path = os.path.join(HERE, "loader_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()

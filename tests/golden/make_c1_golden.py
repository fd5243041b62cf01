"""Generate tests/golden/c1_golden.npz: BASELINE config 1 on the reference's own test images.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_c1_golden.py

Config 1 = "test_data in0062/in0063 + forward.flo/backward.flo: backward-warp in0063 onto in0062 with
consistency mask and composite onto sea.jpg at native resolution".  The two .flo files are missing from the
reference checkout (.MISSING_LARGE_BLOBS), so deterministic stand-ins are computed as SURVEY 8c prescribes:
single-threaded DIS optical flow (PRESET_MEDIUM) between the grey versions of the reference's own composites
cmp1.png / cmp2.png.  The UNMODIFIED reference functions then run the pipeline of flow.py:68-79 (warp_img,
warp_bgr, correct_alpha) and reader.create_composite_image on a 160 x 224 window of the 500 x 1200 frames (the
pure-Python loop of correct_alpha takes ~10 us per pixel); window inputs and outputs are the fixture.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
WIN = (slice(140, 300), slice(500, 724))


def main():
    import cv2
    from make_golden import import_reference, REF
    reader, flow, tps, augmentation = import_reference()
    cv2.setNumThreads(1)
    td = os.path.join(REF, "test_data")
    cmp1, cmp2 = cv2.imread(os.path.join(td, "cmp1.png")), cv2.imread(os.path.join(td, "cmp2.png"))
    g1, g2 = cv2.cvtColor(cmp1, cv2.COLOR_BGR2GRAY), cv2.cvtColor(cmp2, cv2.COLOR_BGR2GRAY)
    dis = cv2.DISOpticalFlow_create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
    f12 = dis.calc(g1, g2, None)            # lives on frame 62, points into frame 63: the "backward" flow of warp_img
    f21 = dis.calc(g2, g1, None)            # lives on frame 63, points into frame 62: the "forward" flow of correct_alpha
    again = cv2.DISOpticalFlow_create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM).calc(g1, g2, None)
    assert np.array_equal(f12, again), "DIS flow must be reproducible"
    a62, fg62 = reader.read_fg_img(os.path.join(td, "in0062.png"))
    a63, fg63 = reader.read_fg_img(os.path.join(td, "in0063.png"))
    sea = cv2.resize(cv2.imread(os.path.join(td, "sea.jpg")), dsize=(fg62.shape[1], fg62.shape[0]), interpolation=cv2.INTER_LINEAR)
    fb = np.ascontiguousarray(f12[WIN]).astype(np.float32)
    ff = np.ascontiguousarray(f21[WIN]).astype(np.float32)
    a63w, fg63w = np.ascontiguousarray(a63[WIN]), np.ascontiguousarray(fg63[WIN])
    bgw = np.ascontiguousarray(sea[WIN])
    walpha = flow.warp_img(a63w, fb)
    wbgr = flow.warp_bgr(fg63w, fb)
    with contextlib.redirect_stdout(io.StringIO()):
        calpha = flow.correct_alpha(fb, ff, walpha.copy())
    cmp_ = reader.create_composite_image(wbgr, bgw, calpha)
    # sanity of the stand-in flows at native resolution: warping alpha63 onto frame 62 must bring it
    # close to alpha62.  (Inside the window fixture the 28 px flows read zeros beyond the window border, so the
    # window outputs are a parity vector, not a registration result.)
    mae_before = float(np.abs(a63 - a62).mean())
    mae_after = float(np.abs(flow.warp_img(a63, f12) - a62).mean())
    assert mae_after < 0.2 * mae_before
    out = {"fg63_bgra": np.concatenate((fg63w, np.rint(a63w * 255.).astype(np.uint8)[..., None]), axis=2),
           "alpha62": np.rint(a62[WIN] * 255.).astype(np.uint8), "backward": fb, "forward": ff, "bg": bgw,
           "warp_alpha": walpha, "warp_bgr": wbgr, "corrected": calpha, "composite": cmp_,
           "mae": np.array([mae_before, mae_after])}
    assert np.array_equal(out["fg63_bgra"][..., 3] / 255., a63w)
    path = os.path.join(HERE, "c1_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; full-frame MAE of alpha63 vs alpha62 {mae_before:.4f} -> "
          f"{mae_after:.4f} after the warp; {int((calpha != walpha).sum())} window pixels masked")


if __name__ == "__main__":
    main()

"""Generate tests/golden/c1_full_golden.npz: BASELINE config 1 at NATIVE resolution (500 x 1200).

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_c1_full_golden.py

Same pipeline as make_c1_golden.py (flow.py:68-79: warp_img, warp_bgr, correct_alpha, then
reader.create_composite_image onto sea.jpg) but on the whole frame, run by the UNMODIFIED reference
(the pure-Python loop of correct_alpha takes ~7 s here).  To keep the fixture small:

* inputs: the bytes of in0063.png and sea.jpg as they lie in the reference's test_data (decoded by the
  product's own reader in the test - this also exercises the uint16 branch of read_fg_img on the real
  file) and the two DIS stand-in flows quantised to 1/16 px and stored as int16 - the
  reference runs on exactly those quantised float32 flows, so nothing depends on re-running DIS;
* outputs that must match bit for bit (warped alpha float64, warped BGR, corrected alpha, occlusion
  mask) as SHA-256 digests plus a few summary numbers, so a failure can be localised;
* the float64 composite (tolerance 1e-5) as float32 on every third row and column; the full composite is
  compared with the oracle, which the CPU suite pins with this same fixture.
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
SUB = 3


def digest(a):
    a = np.ascontiguousarray(a)
    return np.frombuffer(hashlib.sha256(a.tobytes()).digest(), dtype=np.uint8).copy()


def main():
    import cv2
    from make_golden import import_reference, REF
    reader, flow, tps, augmentation = import_reference()
    cv2.setNumThreads(1)
    td = os.path.join(REF, "test_data")
    cmp1, cmp2 = cv2.imread(os.path.join(td, "cmp1.png")), cv2.imread(os.path.join(td, "cmp2.png"))
    g1, g2 = cv2.cvtColor(cmp1, cv2.COLOR_BGR2GRAY), cv2.cvtColor(cmp2, cv2.COLOR_BGR2GRAY)
    dis = cv2.DISOpticalFlow_create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
    q12 = np.rint(dis.calc(g1, g2, None) * 16.0).astype(np.int16)       # on frame 62, points into 63: "backward"
    q21 = np.rint(dis.calc(g2, g1, None) * 16.0).astype(np.int16)       # on frame 63, points into 62: "forward"
    fb, ff = q12.astype(np.float32) / 16.0, q21.astype(np.float32) / 16.0
    a63, fg63 = reader.read_fg_img(os.path.join(td, "in0063.png"))
    a62, _ = reader.read_fg_img(os.path.join(td, "in0062.png"))
    h, w = a63.shape
    assert (h, w) == (500, 1200) and fb.shape == (h, w, 2)
    sea = cv2.resize(cv2.imread(os.path.join(td, "sea.jpg")), dsize=(w, h), interpolation=cv2.INTER_LINEAR)
    walpha = flow.warp_img(a63, fb)
    wbgr = flow.warp_bgr(fg63, fb)
    with contextlib.redirect_stdout(io.StringIO()):
        calpha = flow.correct_alpha(fb, ff, walpha.copy())
    mask = (calpha != walpha)
    cmp_ = reader.create_composite_image(wbgr, sea, calpha)
    mae_before = float(np.abs(a63 - a62).mean())
    mae_after = float(np.abs(walpha - a62).mean())
    assert mae_after < 0.2 * mae_before
    with open(os.path.join(td, "in0063.png"), "rb") as f:
        png = np.frombuffer(f.read(), dtype=np.uint8).copy()
    with open(os.path.join(td, "sea.jpg"), "rb") as f:
        jpg = np.frombuffer(f.read(), dtype=np.uint8).copy()
    out = {"in0063_png": png, "sea_jpg": jpg, "backward_q16": q12, "forward_q16": q21,
           "sha_fg63": digest(fg63), "sha_alpha63": digest(a63), "sha_bg": digest(sea),
           "sha_warp_alpha": digest(walpha), "sha_warp_bgr": digest(wbgr), "sha_corrected": digest(calpha),
           "sha_masked": digest(mask.astype(np.uint8)),
           "n_masked": np.array([int(mask.sum())]), "sum_corrected": np.array([calpha.sum()]),
           "sum_warp_bgr": np.array([int(wbgr.astype(np.int64).sum())]),
           "composite_sub": cmp_[::SUB, ::SUB].astype(np.float32), "sum_composite": np.array([cmp_.sum()]),
           "mae": np.array([mae_before, mae_after])}
    path = os.path.join(HERE, "c1_full_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; MAE {mae_before:.4f} -> {mae_after:.4f}; "
          f"{int(mask.sum())} pixels zeroed by correct_alpha")


if __name__ == "__main__":
    main()

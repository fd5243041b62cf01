"""(Round 1; since round 2 `bench.py` carries these configs in its own line - `configs`.)
Device-resident throughput of the BASELINE.json configs 2-4 (one GPU): frames/s, algorithmic GB/s and
fraction of the measured HBM peak.  python scripts/bench_configs.py [iters] > profiles/rNN_configs.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import __graft_entry__ as ge
import bench
import vm_oracle as O

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
vm = ge.load_package()
P = vm.pipeline
dev = torch.device("cuda", 0)
peak, peak_src = bench.measured_peak()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case(name, which, h, w, n, n_ctrl, bpp):
    fg, fb, ff, bg = bench.make_clip(torch, 77, n, h, w, dev)
    st = vm._native.new_status(dev)
    if which == "c2":
        ob = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        oa = torch.empty((n, h, w), dtype=torch.float32, device=dev)
        ms = timed(lambda: P.flow_warp_mask(fg, fb, ff, out_bgr=ob, out_alpha=oa, status=st))
    else:
        grids = [O.synth_grids(5000 + k, h, w, n_ctrl) for k in range(n)]
        ctrl, coef = P.solve_grids(grids, dev)
        out = torch.empty((n, h, w, 4), dtype=torch.float32, device=dev)
        if which == "c3":
            ms = timed(lambda: P.tps_composite(fg, bg, ctrl, coef, out=out, status=st))
        else:
            ms = timed(lambda: P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st))
    gbs = bpp * h * w * n / (ms / 1e3) / 1e9
    rec = {"config": name, "height": h, "width": w, "frames_per_launch": n, "control_points": n_ctrl * n_ctrl if which != "c2" else None,
           "ms_per_launch": ms, "frames_per_s": n / (ms / 1e3), "algorithmic_bytes_per_px": bpp, "algorithmic_GBps": gbs,
           "frac_of_hbm_peak": gbs / peak, "working_set_MB": bpp * h * w * n / 1e6}
    print(json.dumps(rec), flush=True)
    del fg, fb, ff, bg
    torch.cuda.empty_cache()


print(json.dumps({"peak_GBps": peak, "peak_source": peak_src, "iters": iters, "gpu": torch.cuda.get_device_name(0)}))
case("C2 flow warp + fwd/bwd mask, 1080p x 64", "c2", 1080, 1920, 64, 0, 27)
case("C3 TPS (16 control points) + composite, 512x512 x 256", "c3", 512, 512, 256, 4, 23)
case("C3 TPS (25 control points) + composite, 1080p x 64", "c3", 1080, 1920, 64, 5, 23)
case("C4 flow warp + mask + TPS + composite, 1080p x 64 (headline)", "c4", 1080, 1920, 64, 5, 39)
case("C4 flow warp + mask + TPS + composite, 4K x 16", "c4", 2160, 3840, 16, 5, 39)


def case_augment(h, w, iters):
    """C5 (informational): the drop-in augmentation.augment on device-resident tensors, one frame per call -
    host RNG draws, the host TPS solve (twice, as the reference) and two host syncs (object_size, fg_center) included."""
    import time
    import numpy as np
    frame = O.synth_frame(4242, h, w)
    fg = torch.from_numpy(np.ascontiguousarray(frame[..., :3])).to(dev)
    alpha = torch.from_numpy(frame[..., 3] / 255.).to(dev)
    bgf = torch.from_numpy(O.synth_background(7, h, w)).to(dev)
    np.random.seed(1)
    for _ in range(2):
        vm.augmentation.augment(fg, bgf, alpha)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        vm.augmentation.augment(fg, bgf, alpha)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / iters
    gbs = 17 * h * w / dt / 1e9
    print(json.dumps({"config": "C5 augmentation.augment drop-in, 1080p, one frame per call (host orchestration included)",
                      "height": h, "width": w, "ms_per_frame": dt * 1e3, "frames_per_s": 1 / dt, "algorithmic_bytes_per_px": 17,
                      "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}), flush=True)


case_augment(1080, 1920, 20)


def case_augment_clip(h, w, n, iters):
    """C5, batched: augmentation.augment_clip on a device-resident clip (host RNG + pinv per frame included)."""
    import time
    import numpy as np
    fg, _, _, bg = bench.make_clip(torch, 99, n, h, w, dev)
    bgn = bg[torch.arange(n) % bg.shape[0]].contiguous()
    np.random.seed(1)
    vm.augmentation.augment_clip(fg, bgn)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        vm.augmentation.augment_clip(fg, bgn)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / iters
    gbs = 17 * h * w * n / dt / 1e9
    print(json.dumps({"config": f"C5 augmentation.augment_clip, 1080p x {n} per call (host RNG + pinv per frame included)",
                      "height": h, "width": w, "frames_per_launch": n, "ms_per_call": dt * 1e3, "frames_per_s": n / dt,
                      "algorithmic_bytes_per_px": 17, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}), flush=True)


case_augment_clip(1080, 1920, 64, 5)


def case_augment_clip_stats(h, w, n, iters):
    """C5, batched, the same foregrounds augmented again and again (the reference writes 50 variants per foreground,
    augmentation.py:140): alpha statistics computed once, so a call has no host synchronisation and its host work
    (RNG plan + pinv) overlaps the kernels of the previous call."""
    import time
    import numpy as np
    fg, _, _, bg = bench.make_clip(torch, 99, n, h, w, dev)
    bgn = bg[torch.arange(n) % bg.shape[0]].contiguous()
    stats = vm.augmentation.alpha_stats(fg)
    np.random.seed(1)
    vm.augmentation.augment_clip(fg, bgn, stats=stats)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        vm.augmentation.augment_clip(fg, bgn, stats=stats)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / iters
    gbs = 17 * h * w * n / dt / 1e9
    print(json.dumps({"config": f"C5 augmentation.augment_clip(stats=...), 1080p x {n} per call, alpha statistics reused (no sync per call)",
                      "height": h, "width": w, "frames_per_launch": n, "ms_per_call": dt * 1e3, "frames_per_s": n / dt,
                      "algorithmic_bytes_per_px": 17, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}), flush=True)


case_augment_clip_stats(1080, 1920, 64, 8)

"""Clip ingest on one GPU: PNG + .flo files -> reader.load_clip -> fused C4 pipeline.

    python scripts/bench_ingest.py [frames] [iters] > profiles/rNN_ingest.json

Reports the frames/s of load_clip alone (decode threads -> pinned memory -> chunked async H2D), of load_clip
followed by the C4 kernels (output left on the device), and of the per-file reference readers on one core
(reader.read_fg_img / read_flow semantics via the oracle + cv2 decode) as the CPU baseline.  Files are
1080p RGBA PNGs (compression level 1) and Middlebury .flo files in a temporary directory (page cache)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import __graft_entry__ as ge
import bench
import vm_oracle as O
import cv2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 1080, 1920
vm = ge.load_package()
R, P = vm.reader, vm.pipeline
tmp = tempfile.mkdtemp(prefix="vm_ingest_bench_")
fgp, fbp, ffp, bgp = [], [], [], []
for k in range(n):
    fr = O.synth_frame(100 + k, H, W)
    fr[..., :3] = O.synth_background(k, H, W)                 # natural-image-like colour: realistic PNG sizes
    fb, ff = O.synth_flows(100 + k, H, W)
    fgp.append(os.path.join(tmp, f"fg{k}.png")); cv2.imwrite(fgp[-1], fr, [cv2.IMWRITE_PNG_COMPRESSION, 1])
    fbp.append(os.path.join(tmp, f"b{k}.flo")); O.write_flo(fbp[-1], fb)
    ffp.append(os.path.join(tmp, f"f{k}.flo")); O.write_flo(ffp[-1], ff)
for k in range(2):
    bgp.append(os.path.join(tmp, f"bg{k}.png")); cv2.imwrite(bgp[-1], O.synth_background(50 + k, H, W))
file_bytes = sum(os.path.getsize(p) for p in fgp + fbp + ffp + bgp)

grids = bench.make_grids(vm, 1, n, H, W)
dev = torch.device("cuda", 0)
ctrl, coef = P.solve_grids(grids, dev)
out = torch.empty((n, H, W, 4), dtype=torch.float32, device=dev)
st = vm._native.new_status(dev)


def ingest():
    return R.load_clip(fgp, fbp, ffp, bgp)


def ingest_and_run():
    c = R.load_clip(fgp, fbp, ffp, bgp)
    P.flow_tps_composite(c["fg"], c["backward"], c["forward"], c["bg"], ctrl, coef, out=out, status=st)
    torch.cuda.synchronize()


def timed(fn):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / iters


t_ing, t_all = timed(ingest), timed(ingest_and_run)
m = min(n, 4)
t = time.perf_counter()
for k in range(m):
    img = cv2.imread(fgp[k], cv2.IMREAD_UNCHANGED)
    O.split_fg(img)
    O.parse_flo(open(fbp[k], "rb").read())
    O.parse_flo(open(ffp[k], "rb").read())
t_cpu = (time.perf_counter() - t) / m
print(json.dumps({
    "workload": f"ingest of a {n}-frame 1080p clip: RGBA PNG + backward/forward .flo per frame, 2 backgrounds",
    "gpu": torch.cuda.get_device_name(0), "host_threads": min(32, os.cpu_count() or 1),
    "file_MB_per_frame": file_bytes / n / 1e6, "device_MB_per_frame": (4 + 16) * H * W / 1e6,
    "load_clip": {"s_per_clip": t_ing, "frames_per_s": n / t_ing, "file_GBps": file_bytes / t_ing / 1e9},
    "load_clip_plus_c4": {"s_per_clip": t_all, "frames_per_s": n / t_all},
    "cpu_baseline": {"frames_per_s": 1.0 / t_cpu, "cores": 1, "kind": "port",
                     "sample": f"{m} frames: cv2.imread + oracle split_fg / parse_flo (the reference's per-file readers)"},
}))

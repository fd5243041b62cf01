"""Throughput of the batch-loader row on one GPU.

    python scripts/bench_loader.py [batch] [iters] > profiles/rNN_loader.json

Three numbers for loader.video_batch on 1080p foregrounds / 720p backgrounds, input_size 320x320:
  * device: vm_loader_batch alone on staged samples (CUDA events), with the bytes it must move
    (decoded uint8 windows in, 13 float64 planes out) as a fraction of the measured HBM peak;
  * end to end: loader.video_batch(file list) - PNG / .flo decode in the host thread pool, planning,
    one H2D copy, the kernel, D2H of the float64 batch;
  * CPU: the NumPy oracle of the same samples on one host core (decode included), as the baseline.
Synthetic files are written to a temporary directory first (not timed)."""
import ctypes
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
import vm_loader_oracle as LO
import vm_oracle as O

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
H, W, BH, BW, SIZE = 1080, 1920, 720, 1280, (320, 320)

vm = ge.load_package()
L = vm.loader
N = vm._native
import cv2

peak, peak_src = bench.measured_peak()
tmp = tempfile.mkdtemp(prefix="vm_loader_bench_")
entries = []
for k in range(batch):
    fg = O.synth_frame(100 + k, H, W)
    prev = O.synth_frame(500 + k, H, W)
    flo, _ = O.synth_flows(100 + k, H, W)
    bg = O.synth_background(k, BH, BW)
    p = [os.path.join(tmp, f"{name}_{k}.{ext}") for name, ext in (("fg", "png"), ("bg", "png"), ("prev", "png"), ("flow", "flo"))]
    cv2.imwrite(p[0], fg, [cv2.IMWRITE_PNG_COMPRESSION, 1])
    cv2.imwrite(p[1], bg, [cv2.IMWRITE_PNG_COMPRESSION, 1])
    cv2.imwrite(p[2], prev, [cv2.IMWRITE_PNG_COMPRESSION, 1])
    O.write_flo(p[3], flo)
    entries.append(tuple(p))

# ---- end to end ---------------------------------------------------------------------------------
np.random.seed(0)
L.video_batch(entries, SIZE)
t0 = time.perf_counter()
for _ in range(iters):
    out = L.video_batch(entries, SIZE)
e2e_s = (time.perf_counter() - t0) / iters
t0 = time.perf_counter()
for _ in range(iters):
    L._decode_all("video", entries)
decode_s = (time.perf_counter() - t0) / iters

# ---- device only --------------------------------------------------------------------------------
np.random.seed(0)
decoded = L._decode_all("video", entries)
records = np.zeros(batch, dtype=L.SAMPLE_DTYPE)
in_bytes = 0
for d, r in zip(decoded, records):
    r["fgv"], r["bgv"] = L._plan_sample(H, W, BH, BW, SIZE)
    win = int(r["fgv"]["win_h"]) * int(r["fgv"]["win_w"])
    in_bytes += win * (4 + 8 + 1) + BH * BW * 3            # fg BGRA + flow + previous alpha at the window, whole bg
dev, host, table = L._stage(decoded, records)
outs = [torch.empty((batch, SIZE[1], SIZE[0], c), dtype=torch.float64, device="cuda") for c in (3, 3, 1, 3, 3)]
mean = (ctypes.c_double * 3)(*L.VGG_MEAN)
lib = N.load()


def launch():
    N.check(lib.vm_loader_batch(ctypes.c_void_p(table), batch, SIZE[1], SIZE[0], mean, N.VM_F64,
                                *[N.ptr(o) for o in (outs[0], outs[1], outs[2], outs[3], outs[4])], N.stream_ptr()))


for _ in range(3):
    launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    launch()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
out_bytes = batch * SIZE[0] * SIZE[1] * 13 * 8
gbs = (in_bytes + out_bytes) / (ms / 1e3) / 1e9

# ---- CPU oracle on one core ---------------------------------------------------------------------
ncpu = min(batch, 4)
t0 = time.perf_counter()
rng = np.random.RandomState(0)
for e in entries[:ncpu]:
    fg = cv2.imread(e[0], cv2.IMREAD_UNCHANGED)
    bg = cv2.imread(e[1])
    prev = cv2.imread(e[2], cv2.IMREAD_UNCHANGED)
    flo, _ = O.parse_flo(open(e[3], "rb").read())
    LO.video_sample(fg, bg, prev, flo, SIZE, rng)
cpu_s = (time.perf_counter() - t0) / ncpu

print(json.dumps({
    "workload": f"loader.video_batch: {batch} samples, 1080p RGBA foreground + previous frame + .flo, 720p background, input_size 320x320",
    "gpu": torch.cuda.get_device_name(0),
    "device": {"ms_per_batch": ms, "samples_per_s": batch / (ms / 1e3), "algorithmic_bytes_per_batch": in_bytes + out_bytes,
               "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak, "peak_GBps": peak, "peak_source": peak_src,
               "note": "launch-size bound: one batch is only %.1f MB" % ((in_bytes + out_bytes) / 1e6)},
    "e2e": {"s_per_batch": e2e_s, "samples_per_s": batch / e2e_s, "decode_s_per_batch": decode_s,
            "decode_threads": L.DECODE_THREADS, "note": "host PNG/.flo decode dominates; files in the page cache"},
    "cpu_baseline": {"samples_per_s": 1.0 / cpu_s, "cores": 1, "kind": "port",
                     "sample": f"{ncpu} samples through oracle.vm_loader_oracle.video_sample incl. cv2 decode"},
}))

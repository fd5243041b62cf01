"""Per-stage durations (CUDA events inside the library, vm_lean_stage_ms) of the lean pipeline for one config:
    python scripts/stage_times.py [c3|c4] [h] [w] [frames] [n_ctrl_side]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import __graft_entry__ as ge
import bench
import vm_oracle as O

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
h, w, n, nc = (int(v) for v in (sys.argv[2:6] + ["1080", "1920", "64", "5"][len(sys.argv) - 2:]))
vm = ge.load_package()
P, N = vm.pipeline, vm._native
for kv in filter(None, os.environ.get("VM_OPTS", "").split(",")):
    k, v = kv.split("=")
    N.set_option(k, int(v))
dev = torch.device("cuda", 0)
fg, fb, ff, bg = bench.make_clip(torch, 77, n, h, w, dev)
grids = [O.synth_grids(5000 + k, h, w, nc) for k in range(n)]
ctrl, coef = P.solve_grids(grids, dev)
out = torch.empty((n, h, w, 4), dtype=torch.float32, device=dev)
st = N.new_status(dev)
run = (lambda: P.tps_composite(fg, bg, ctrl, coef, out=out, status=st)) if which == "c3" else \
      (lambda: P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st))
for _ in range(3):
    run()
torch.cuda.synchronize()
N.set_option("lean_timing", 1)
acc = [0.0] * 4
iters = 10
for _ in range(iters):
    run()
    torch.cuda.synchronize()
    ms = (ctypes.c_float * 4)()
    N.check(N.load().vm_lean_stage_ms(ms))
    acc = [a + float(m) for a, m in zip(acc, ms)]
N.set_option("lean_timing", 0)
names = ["spline", "boxes", "flow stage", "resampling"]
tot = sum(acc) / iters
print(f"{which} {n} x {h}x{w}, {nc * nc} control points: " + ", ".join(f"{nm} {a / iters * 1e3 / n:.2f} us" for nm, a in zip(names, acc)) +
      f" per frame; total {tot * 1e3 / n:.2f} us per frame = {n / tot * 1e3:.0f} frames/s; status {st.tolist()}")

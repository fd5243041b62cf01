"""Per-stage durations of the C4 headline clip (1080p x 64, device-resident) under several option sets, with a
bit-identity check of the output against the first set:
    python scripts/stage_sweep.py "" "flow_stage_layout=2" "lean_minb=3,flow_stage_layout=2" ...
Each argument is a comma-separated list of vm_set_option key=value pairs applied on top of the defaults."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import bench

vm = ge.load_package()
P, Nt = vm.pipeline, vm._native
lib = Nt.load()
dev = torch.device("cuda", 0)
H, W, n = bench.H, bench.W, int(os.environ.get("VM_FRAMES", 64))
fg, fb, ff, bg = bench.make_clip(torch, 1234, n, H, W, dev)
ctrl, coef = P.solve_grids(bench.make_grids(vm, 1, n, H, W), dev)
out = torch.empty((n, H, W, 4), dtype=torch.float32, device=dev)
ref = None
DEFAULTS = {"fused_variant": P.DEFAULT_VARIANT, "flow_stage_layout": 0, "lean_minb": 4, "lean_fine_rows": 8, "lean_tmap": 1}
names = ["spline", "boxes", "flow", "resample"]
iters = int(os.environ.get("VM_ITERS", 10))
for arg in sys.argv[1:] or [""]:
    opts = dict(DEFAULTS)
    opts.update({k: int(v) for k, v in (kv.split("=") for kv in arg.split(",") if kv)})
    try:
        for k, v in opts.items():
            Nt.set_option(k, v)
    except Exception as e:                                   # option not in this build
        print(f"{arg:50s} skipped: {e}", flush=True)
        continue
    st = Nt.new_status(dev)
    run = lambda: P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(out, ref))
    Nt.set_option("lean_timing", 1)
    acc = [0.0] * 4
    for _ in range(5):
        run()
        torch.cuda.synchronize()
        t = (ctypes.c_float * 4)()
        Nt.check(lib.vm_lean_stage_ms(t))
        acc = [a + float(m) / 5 for a, m in zip(acc, t)]
    Nt.set_option("lean_timing", 0)
    print(f"{arg:50s} {ms:7.3f} ms  frac {39 * H * W * n / ms / 1e6 / bench.measured_peak()[0]:.3f}  " +
          " ".join(f"{nm} {a:.3f}" for nm, a in zip(names, acc)) + f"  same={same}", flush=True)

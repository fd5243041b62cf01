"""Time the C4 headline clip (1080p x 64, device-resident) under several option sets:
    python scripts/opt_sweep.py "fused_variant=4" "fused_variant=4,lean_pipe=1" ...
Each argument is a comma-separated list of vm_set_option key=value pairs applied on top of the defaults."""
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
DEFAULTS = {"fused_variant": P.DEFAULT_VARIANT, "lean_chunk": 0, "lean_sub": 0, "lean_b1_warps": 16,
            "lean_b1_dyr": 1, "lean_b1_ctas": 0, "lean_minb": 4, "lean_fine_rows": 8, "lean_rb": 0, "lean_box_cap": 0}
for arg in sys.argv[1:] or ["fused_variant=4"]:
    opts = dict(DEFAULTS)
    opts.update({k: int(v) for k, v in (kv.split("=") for kv in arg.split(",") if kv)})
    for k, v in opts.items():
        Nt.set_option(k, v)
    st = Nt.new_status(dev)
    run = lambda: P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = int(os.environ.get("VM_ITERS", 10))
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(out, ref))
    print(f"{arg:60s} {ms:7.3f} ms = {ms / n * 1e3:6.1f} us/frame  frac {39 * H * W * n / ms / 1e6 / bench.measured_peak()[0]:.3f}  same={same}", flush=True)

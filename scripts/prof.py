"""Small fixed workload for ncu / timing: a few launches of the fused kernels at 1080p.

    python scripts/prof.py [c4|c2|c3] [frames] [iters]
"""
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
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H, W = bench.H, bench.W
vm = ge.load_package()
P = vm.pipeline
if os.environ.get("VM_VARIANT"):
    P.set_fused_variant(int(os.environ["VM_VARIANT"]))
for kv in filter(None, os.environ.get("VM_OPTS", "").split(",")):
    k, v = kv.split("=")
    vm._native.set_option(k, int(v))
dev = torch.device("cuda", 0)
fg, fb, ff, bg = bench.make_clip(torch, 1234, n, H, W, dev)
grids = bench.make_grids(vm, 1, n, H, W)
ctrl, coef = P.solve_grids(grids, dev)
out = torch.empty((n, H, W, 4), dtype=torch.float32, device=dev)
ob = torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev)
oa = torch.empty((n, H, W), dtype=torch.float32, device=dev)
st = vm._native.new_status(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(iters + 1):
    if it == 1:
        e0.record()
    if which == "c4":
        P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st)
    elif which == "c3":
        P.tps_composite(fg, bg, ctrl, coef, out=out, status=st)
    else:
        P.flow_warp_mask(fg, fb, ff, out_bgr=ob, out_alpha=oa, status=st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
bpp = {"c4": 39, "c3": 23, "c2": 27}[which]
print(f"{which}: {n} frames {ms:.3f} ms/launch = {ms / n * 1e3:.1f} us/frame, "
      f"{bpp * H * W * n / ms / 1e6:.0f} GB/s algorithmic, status {st.tolist()}")

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
for key in ("pipe_lead", "pipe_ring_rows", "pipe_cring_rows", "pipe_blocks", "pipe_roles", "chunk_frames"):
    if os.environ.get("VM_" + key.upper()):
        vm._native.set_option(key, int(os.environ["VM_" + key.upper()]))
for kv in filter(None, os.environ.get("VM_OPTS", "").split(",")):
    k, v = kv.split("=")
    vm._native.set_option(k, int(v))
if os.environ.get("VM_TILE_H"):
    vm._native.set_option("tile_h", int(os.environ["VM_TILE_H"]))
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

if os.environ.get("VM_TIMING"):
    import numpy as np
    sc = next(iter(P._scratch_cache.values()))
    off = (-sc.data_ptr()) % 256
    hdr = sc[off:off + 64].cpu().numpy().view(np.uint32)
    names = {2: "A n", 3: "A ns", 4: "A wait", 5: "R n", 6: "R ns", 8: "B n", 9: "B loop", 10: "B waitR", 11: "B waitA+tc"}
    its = iters + 1
    for k, nm in names.items():
        print(f"   {nm:12s} {hdr[k]}")
    print(f"   per item: A {hdr[3]/max(hdr[2],1):.0f} ns (+wait {hdr[4]/max(hdr[2],1):.0f}), R {hdr[6]/max(hdr[5],1):.0f} ns, "
          f"B loop {hdr[9]/max(hdr[8],1):.0f} ns, waitR {hdr[10]/max(hdr[8],1):.0f}, waitA+tc {hdr[11]/max(hdr[8],1):.0f}")

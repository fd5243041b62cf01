"""End-to-end (pinned host -> device -> pinned host) rate of the C4 pipeline for several chunk sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import __graft_entry__ as ge
import bench
import vm_oracle as O
vm = ge.load_package(); P = vm.pipeline
dev = torch.device("cuda", 0)
n, H, W = 64, bench.H, bench.W
fg, fb, ff, bg = bench.make_clip(torch, 1, n, H, W, dev)
grids = bench.make_grids(vm, 1, n, H, W)
host = [t.cpu().pin_memory() for t in (fg, fb, ff)]
bg_h = bg[torch.arange(n) % bg.shape[0]].cpu().pin_memory()
out_h = torch.empty((n, H, W, 4), dtype=torch.float32).pin_memory()
for chunk in [int(a) for a in sys.argv[1:]] or [2, 4, 8, 16]:
    P.flow_tps_composite_host(host[0], host[1], host[2], bg_h, grids, out=out_h, chunk=chunk)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        P.flow_tps_composite_host(host[0], host[1], host[2], bg_h, grids, out=out_h, chunk=chunk)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print(f"chunk {chunk}: {n / dt:.0f} frames/s ({dt * 1e3:.1f} ms per 64 frames)")

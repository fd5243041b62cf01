"""Cycles per phase of k_fuse_c4 (library built with VM_NVCC_EXTRA=-DVF_TIMING):
    VM_NVCC_EXTRA=-DVF_TIMING python video-matting_b200/_build.py --force && python scripts/fuse_phases.py [frames]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
vm = ge.load_package()
P = vm.pipeline
P.set_fused_variant(5)                # the single-pass kernel (the default is the lean split pipeline)
lib = vm._native.load()
for kv in filter(None, os.environ.get("VM_OPTS", "").split(",")):
    k, v = kv.split("=")
    vm._native.set_option(k, int(v))
dev = torch.device("cuda", 0)
H, W = bench.H, bench.W
fg, fb, ff, bg = bench.make_clip(torch, 1234, n, H, W, dev)
ctrl, coef = P.solve_grids(bench.make_grids(vm, 1, n, H, W), dev)
out = torch.empty((n, H, W, 4), dtype=torch.float32, device=dev)
st = vm._native.new_status(dev)
buf = (ctypes.c_ulonglong * 16)()
lib.vm_fuse_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(3):
    P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st)
    torch.cuda.synchronize()
    lib.vm_fuse_prof_read(buf, 1)
v = list(buf)
ts, tp = max(v[2], 1), max(v[10], 1)
print(f"spline role: {ts} tiles; per tile: wait-for-buffer {v[0]/ts:.0f} cyc, compute {v[1]/ts:.0f} cyc")
names = ["wait for T", "P2 (axis + Cs)", "P3 own work", "P3 barrier wait", "P4 own work", "P4 barrier wait"]
tot = sum(v[4:10]) / tp
print(f"pixel role (thread 0 of the role): {tp} tiles, {tot:.0f} cyc per tile")
for nm, x in zip(names, v[4:10]):
    print(f"   {nm:18s} {x/tp:8.0f} cyc  {100*x/tp/tot:5.1f} %")

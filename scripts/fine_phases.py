"""Cycles per phase of the resampling CTAs k_lean_fine (library built with VM_NVCC_EXTRA=-DVL_TIMING):
    VM_NVCC_EXTRA=-DVL_TIMING python video-matting_b200/_build.py --force && python scripts/fine_phases.py [frames]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vm = ge.load_package()
P = vm.pipeline
lib = vm._native.load()
dev = torch.device("cuda", 0)
H, W = bench.H, bench.W
fg, fb, ff, bg = bench.make_clip(torch, 1234, n, H, W, dev)
ctrl, coef = P.solve_grids(bench.make_grids(vm, 1, n, H, W), dev)
out = torch.empty((n, H, W, 4), dtype=torch.float32, device=dev)
st = vm._native.new_status(dev)
buf = (ctypes.c_ulonglong * 16)()
lib.vm_lean_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(3):
    P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=out, status=st)
    torch.cuda.synchronize()
    lib.vm_lean_prof_read(buf, 1)
v = list(buf)
names = ["P0 set-up, axis entries, 2 CTA barriers", "P1 issue of the bulk copies", "P2 transform window -> Cs", "wait for the bulk copies + barrier", "P3 resampling + composite"]
for who, o in (("thread 0 (warp 0: issues and polls the copies)", 0), ("last thread (warp 7)", 8)):
    tiles = max(v[o + 7], 1)
    tot = sum(v[o:o + 5]) / tiles
    print(f"{who}: {tiles} tiles, {tot:.0f} cycles per tile")
    for nm, x in zip(names, v[o:o + 5]):
        print(f"   {nm:42s} {x / tiles:8.0f} cyc  {100 * x / tiles / tot:5.1f} %")

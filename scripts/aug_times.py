"""Workload for the per-kernel times of augmentation.augment_clip (BASELINE config 5): 1080p x 64, a few calls.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/aug_times.py
    python scripts/launch_times.py out.csv k_
Without ncu it prints the wall-clock time per call (solver pool of 8 workers, alpha statistics reused)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
import bench

vm = ge.load_package()
P = vm.pipeline
dev = torch.device("cuda", 0)
h, w, n = 1080, 1920, int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fg, fb, ff, bg = bench.make_clip(torch, 77, n, h, w, dev)
bgn = bg[torch.arange(n) % bg.shape[0]].contiguous()
stats = vm.augmentation.alpha_stats(fg)
np.random.seed(1)
pool = P.SolverPool(8)
try:
    for _ in range(2):
        vm.augmentation.augment_clip(fg, bgn, stats=stats, pool=pool)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        vm.augmentation.augment_clip(fg, bgn, stats=stats, pool=pool)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t) / iters
    print(f"augment_clip {n} x {h}x{w}: {ms:.2f} ms per call = {n / ms:.2f} kfps")
finally:
    pool.close()

"""CPU tests (no GPU needed): the C-ABI library loads and exports every symbol declared in
include/vm_b200.h, the host-side logic (TPS solve, axis tables, RNG order, LUT, sharding)
matches the oracle, and the product refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import vm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vm):
    lib = ctypes.CDLL(vm._native.lib_path() if os.path.exists(vm._native.lib_path()) else vm._build.build())
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vm_b200.h but not exported"
    assert set(vm._native.SIGNATURES) == set(syms), "ctypes binding and header disagree"
    bound = vm._native.load()
    assert bound.vm_version() >= 100
    assert bound.vm_last_error_string() is not None


def test_no_cpu_fallback(vm):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(vm._native.VmError):
        vm.flow.warp_img(np.zeros((4, 4)), np.zeros((4, 4, 2), np.float32))
    with pytest.raises(vm._native.VmError):
        vm.reader.create_composite_image(np.zeros((2, 2, 3), np.uint8), np.zeros((2, 2, 3), np.uint8), np.zeros((2, 2)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "video-matting_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "vm_oracle" not in txt and "oracle/" not in txt, f"{f} references the oracle"


def test_tps_solve_bit_equal_to_oracle(vm):
    for seed, (h, w), n in ((1, (1080, 1920), 5), (2, (512, 512), 4), (3, (61, 81), 5), (4, (2160, 3840), 5)):
        grid, dgrid = O.synth_grids(seed, h, w, n)
        assert np.array_equal(vm.pipeline.tps_solve(dgrid, grid), O.tps_coefficients(dgrid, grid))
        assert np.array_equal(vm.pipeline._tps_kernel_matrix(dgrid), O.tps_system(dgrid))


def test_batched_solve_bit_equal_to_per_frame_solve(vm):
    """solve_grids' stacked / threaded pinv == the reference's per-frame np.dot(pinv(L), V) (tps.py:119)."""
    P = vm.pipeline
    grids = [O.synth_grids(900 + k, 1080, 1920, 5) for k in range(20)] + [O.synth_grids(5, 512, 512, 4)] * 0
    ref = np.stack([P.tps_solve(d, g) for (g, d) in grids])
    for chunk in (grids, grids[:3]):
        got = np.concatenate([P._solve_chunk(chunk[i:i + 8]) for i in range(0, len(chunk), 8)])
        assert np.array_equal(got, ref[:len(chunk)])


def test_axis_tables_match_reference_index_math(vm):
    for h in (61, 64, 500, 1080, 1081):
        xs = h / 2
        tab = vm.pipeline.axis_table(0, h, xs)
        ni = np.arange(h + 1)
        frac, idx = np.modf((xs - 1) * ni / float(h))
        assert np.array_equal(tab["frac"], frac) and np.array_equal(tab["i0"], idx.astype(int))
        assert np.array_equal(tab["i1"], (idx.astype(int) + 1).clip(0, xs - 1).astype(int))
        assert tab["i1"].max() <= int(xs) - 1 and tab.dtype.itemsize == 16


def test_deform_grid_rng_order(vm, golden):
    np.random.seed(1234)
    g, d = vm.augmentation.deform_grid(108, 192)
    assert np.array_equal(g, golden["grid_108x192_n5"]) and np.array_equal(d, golden["defgrid_108x192_n5"])
    np.random.seed(1234)
    g, d = vm.augmentation.deform_grid(64, 64, n=4)
    assert np.array_equal(g, golden["grid_64x64_n4"]) and np.array_equal(d, golden["defgrid_64x64_n4"])


def test_illumination_lut_and_rotation_matrix(vm):
    A = vm.augmentation
    x = np.arange(256, dtype=np.uint8)
    for (a, b, c) in ((1.03, 0.8, -0.02), (0.95, 1.3, 0.07), (1.05, 0.7, -0.07)):
        assert np.array_equal(A.illumination_lut(a, b, c), O.illumination_sv(x, a, b, c))
    assert np.array_equal(A._rotation_matrix((52, 31), 7.5, 1.1), O.rotation_matrix_2d((52, 31), 7.5, 1.1))
    assert np.array_equal(A.identity(3, 4)[2, 3], [3, 4])


def test_shard_range_partitions(vm):
    for n in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [vm.pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as ge
vm = ge.load_package()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
lo, hi = vm.pipeline.shard_range(13, rank, 2)
# host-side aggregation used by bench.py: units summed, elapsed = max over ranks
units = torch.tensor([float(hi - lo)]); ms = torch.tensor([10.0 + rank])
dist.all_reduce(units, op=dist.ReduceOp.SUM); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
assert units.item() == 13 and ms.item() == 11.0
covered = torch.zeros(13); covered[lo:hi] = 1
dist.all_reduce(covered)
assert bool((covered == 1).all())
dist.destroy_process_group()
print("ok", rank)
'''


def test_clip_sharding_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_batched_augment_plan_is_the_sequential_plan(vm):
    """One random_sample block for a clip == the reference's 40 uniform() calls per frame, value for value, and
    leaves np.random in the same state."""
    A = vm.augmentation
    for (h, w), n in (((1080, 1920), 9), ((61, 83), 4), ((512, 512), 2)):
        stats = np.stack([np.array([1000 + 37 * k, (1000 + 37 * k) * (h // 3), (1000 + 37 * k) * (w // 2 + k)]) for k in range(n)])
        np.random.seed(99)
        seq = A._augment_plan(stats, h, w, before_frame=lambda k: None)        # a hook forces the frame-by-frame path
        after_seq = np.random.random_sample()
        np.random.seed(99)
        bat = A._augment_plan(stats, h, w)
        assert np.random.random_sample() == after_seq
        for name, x, y in zip(("bg params", "fg params", "luts"), seq[:3], bat[:3]):
            assert x.tobytes() == y.tobytes(), name
        for (g0, d0), (g1, d1) in zip(seq[3], bat[3]):
            assert np.array_equal(g0, g1) and np.array_equal(d0, d1)
    with pytest.raises(ValueError):
        A._augment_plan(np.array([[5, 1, 1], [0, 0, 0]]), 64, 64)


def test_stacked_kernel_matrices_bit_equal_to_per_frame(vm):
    P, A = vm.pipeline, vm.augmentation
    np.random.seed(3)
    for (h, w), n in (((1080, 1920), 5), ((2160, 3840), 5), ((512, 512), 4), ((61, 83), 3)):
        pts = np.stack([A.deform_grid(h, w, n)[1] for _ in range(33)])
        L = P._tps_kernel_matrices(pts)
        ref = np.stack([P._tps_kernel_matrix(p) for p in pts])
        assert L.tobytes() == ref.tobytes()
    grids = [A.deform_grid(1080, 1920) for _ in range(40)]
    P._stacked_kernel_ok[0] = None
    a = P._solve_chunk(grids[:16])            # first stack of the process: checked against the per-frame build
    assert P._stacked_kernel_ok[0] is True
    b = P._solve_chunk(grids[:16])
    P._stacked_kernel_ok[0] = False
    c = P._solve_chunk(grids[:16])
    P._stacked_kernel_ok[0] = True
    assert a.tobytes() == b.tobytes() == c.tobytes()

"""GPU tests of the single-pass C4 kernel (csrc/vm_fuse.cu, fused_variant 5 of
vm_flow_tps_composite_bgra): bit-identical to the lean split pipeline (same arithmetic, different
schedule), equal to the oracle, and exact on its slow paths (source box larger than shared memory, control
points on coarse grid points, frames smaller than a tile, run-time control-point counts, no forward flow).
"""
import numpy as np
import pytest
import torch

import vm_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def case(h, w, n, n_ctrl, seed=300, stretch=1.0):
    frames = np.stack([O.synth_frame(seed + k, h, w) for k in range(n)])
    flows = [O.synth_flows(seed + k, h, w) for k in range(n)]
    grids = []
    for k in range(n):
        g, d = O.synth_grids(seed + k, h, w, n_ctrl)
        grids.append((g, g + (d - g) * stretch))
    bgs = np.stack([O.synth_background(k, h, w) for k in range(n)])
    return frames, np.stack([f[0] for f in flows]), np.stack([f[1] for f in flows]), grids, bgs


def run(vm, variant, args, forward=True):
    P = vm.pipeline
    fg, fb, ff, bg, ctrl, coef = args
    P.set_fused_variant(variant)
    try:
        out, st = P.flow_tps_composite(fg, fb, ff if forward else None, bg, ctrl, coef)
        torch.cuda.synchronize()
    finally:
        P.set_fused_variant(P.DEFAULT_VARIANT)
    return out, st.cpu().numpy()


def report_diff(a, b):
    d = (a != b).any(dim=-1)
    idx = torch.nonzero(d)
    first = idx[:5].tolist()
    return f"{int(d.sum())} differing pixels, first (frame, row, col): {first}; " \
           f"got {a[tuple(idx[0])].tolist() if len(idx) else None} vs {b[tuple(idx[0])].tolist() if len(idx) else None}"


@pytest.mark.parametrize("h,w,n,n_ctrl", [(200, 336, 3, 5), (61, 83, 2, 5), (120, 168, 5, 4), (59, 60, 2, 5), (60, 61, 1, 5),
                                          (4, 6, 2, 2), (128, 96, 2, 3), (333, 200, 2, 6), (540, 960, 2, 5)])
def test_fuse_equals_lean_bit_for_bit(vm, h, w, n, n_ctrl):
    frames, fb, ff, grids, bgs = case(h, w, n, n_ctrl)
    ctrl, coef = vm.pipeline.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    for forward in (True, False):
        a, sta = run(vm, 5, args, forward)
        b, stb = run(vm, 4, args, forward)
        assert torch.equal(a, b), f"{h}x{w} forward={forward}: " + report_diff(a, b)
        assert sta[3] == stb[3], "same number of samples outside the source"
        assert sta[6] == stb[6], "same number of near-knife-edge samples (status word 6)"
        assert sta[5] == 0


def test_fuse_cta_count_does_not_change_the_result(vm):
    """the number of persistent CTAs only changes the schedule (tiles are handed out by stride)"""
    frames, fb, ff, grids, bgs = case(200, 336, 4, 5, seed=40)
    ctrl, coef = vm.pipeline.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    base, _ = run(vm, 4, args)
    Nt = vm._native
    try:
        for ctas in (0, 1, 3, 7):
            Nt.set_option("fuse_ctas", ctas)
            got, _ = run(vm, 5, args)
            assert torch.equal(got, base), f"ctas={ctas}: " + report_diff(got, base)
    finally:
        Nt.set_option("fuse_ctas", 0)


def test_fuse_vs_oracle_and_knife_count(vm):
    h, w, n = 200, 333, 3
    frames, fb, ff, grids, bgs = case(h, w, n, 5, seed=50)
    ctrl, coef = vm.pipeline.solve_grids(grids)
    out, st = run(vm, 5, (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef))
    out = out.cpu().numpy()
    for k in range(n):
        rc, ra = O.pipeline_c4(frames[k], fb[k], ff[k], grids[k], bgs[k])
        assert np.allclose(out[k][..., 3], ra, rtol=RTOL, atol=1e-6)
        assert np.count_nonzero(~np.isclose(out[k][..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1
    assert st[0] == 0 and st[1] == 0 and st[5] == 0
    assert st[6] <= 2, "near-knife-edge samples are counted (status word 6) and rare"


def test_fuse_stretched_grid_slow_tiles(vm):
    """30 % displacements: source boxes beyond the shared-memory capacity (taps evaluated one by one) and control
    points close to coarse grid points (generic spline path) - still the lean pipeline's bits and the oracle's values."""
    h, w = 256, 320
    frames, fb, ff, grids, bgs = case(h, w, 1, 5, seed=77, stretch=6.0)
    ctrl, coef = vm.pipeline.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    a, sta = run(vm, 5, args)
    b, _ = run(vm, 4, args)
    assert torch.equal(a, b), report_diff(a, b)
    rc, ra = O.pipeline_c4(frames[0], fb[0], ff[0], grids[0], bgs[0])
    got = a[0].cpu().numpy()
    assert np.allclose(got[..., 3], ra, rtol=RTOL, atol=1e-6)
    assert np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1


def test_fuse_folded_grid_box_overflow(vm):
    """A grid scaled so that one 60 x 60 tile reads a source region larger than the 6144-entry box: the tile must
    take the tap-by-tap path (status word 4) and still match."""
    h, w = 512, 640
    frames, fb, ff, grids, bgs = case(h, w, 1, 5, seed=78)
    g, d = grids[0]
    c = np.array([h / 2.0, w / 2.0])
    d2 = c + (g - c) * 0.45 + (d - g)              # the output frame looks at a 2.2 x magnified source region
    grids = [(g, d2)]
    ctrl, coef = vm.pipeline.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    a, sta = run(vm, 5, args)
    b, _ = run(vm, 4, args)
    assert torch.equal(a, b), report_diff(a, b)
    assert sta[4] > 0, "expected tiles whose source box exceeds shared memory"
    rc, ra = O.pipeline_c4(frames[0], fb[0], ff[0], grids[0], bgs[0])
    got = a[0].cpu().numpy()
    assert np.allclose(got[..., 3], ra, rtol=RTOL, atol=1e-6)
    assert np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1


def test_fuse_control_point_on_grid_point(vm):
    h, w = 128, 160
    frames, fb, ff, grids, bgs = case(h, w, 1, 4, seed=55)
    g, d = grids[0]
    d = d.copy()
    sx, sy = h / float(h // 2 - 1), w / float(w // 2 - 1)          # coarse grid steps (tps.py:47)
    d[5] = (20 * sx, 31 * sy)                                       # exactly on coarse point (20, 31): r = 0 -> U = 0
    d[6] = (33 * sx + 1e-3, 40 * sy - 2e-3)                         # r^2 below the log table
    grids = [(g, d)]
    ctrl, coef = vm.pipeline.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    a, _ = run(vm, 5, args)
    b, _ = run(vm, 4, args)
    assert torch.equal(a, b), report_diff(a, b)
    rc, ra = O.pipeline_c4(frames[0], fb[0], ff[0], grids[0], bgs[0])
    got = a[0].cpu().numpy()
    assert np.allclose(got[..., 3], ra, rtol=RTOL, atol=1e-6)
    assert np.count_nonzero(~np.isclose(got[..., :3], rc, rtol=RTOL, atol=1e-5)) <= 1


def test_fuse_wild_flows_match_lean(vm):
    """flows of sigma 50 / 500 px and NaN / Inf components: same bits as the lean pipeline, error pixels flagged"""
    h, w, n = 72, 100, 2
    frames, fb, ff, grids, bgs = case(h, w, n, 5, seed=61)
    rng = np.random.default_rng(9)
    fb = (fb + rng.normal(0, 50, fb.shape)).astype(np.float32)
    ff = (ff + rng.normal(0, 500, ff.shape)).astype(np.float32)
    fb[0, 5, 7, 0] = np.nan; fb[1, 9, 3, 1] = np.inf; fb[1, 20, 30, 0] = 3e9
    ctrl, coef = vm.pipeline.solve_grids(grids)
    args = (dev(frames), dev(fb), dev(ff), dev(bgs), ctrl, coef)
    a, sta = run(vm, 5, args)
    b, stb = run(vm, 4, args)
    assert torch.equal(a, b), report_diff(a, b)
    assert (sta[1] > 0) == (stb[1] > 0) and (sta[0] > 0) == (stb[0] > 0)


def test_fuse_launch_count_and_graph(vm):
    """one kernel per call, capturable in a CUDA graph"""
    P, Nt = vm.pipeline, vm._native
    lib = Nt.load()
    frames, fb, ff, grids, bgs = case(120, 160, 4, 5)
    ctrl, coef = P.solve_grids(grids)
    fg_d, fb_d, ff_d, bg_d = dev(frames), dev(fb), dev(ff), dev(bgs)
    P.set_fused_variant(5)
    before = lib.vm_fuse_launch_count()
    ref, _ = P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef)
    assert lib.vm_fuse_launch_count() - before == 1
    out = torch.empty_like(ref)
    status = Nt.new_status()
    plan = P.get_plan((0, 0, 120, 160), 2, fg_d.device)
    scratch = torch.empty(int(lib.vm_fused_scratch_bytes(4, 120, 160)) + 512, dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef, plan=plan, out=out, scratch=scratch, status=status)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        P.flow_tps_composite(fg_d, fb_d, ff_d, bg_d, ctrl, coef, plan=plan, out=out, scratch=scratch, status=status)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    P.set_fused_variant(P.DEFAULT_VARIANT)
    assert torch.equal(out, ref)

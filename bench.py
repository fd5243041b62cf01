"""bench.py - throughput of the B200 video-matting data path.

Headline workload (BASELINE.json metric "1080p frames/s (warp+TPS+composite)"): per GPU, clips
of 64 synthetic 1080p BGRA frames go through flow warp + forward/backward occlusion mask +
thin-plate-spline deformation (25 control points, fresh grid per frame) + composite onto a
background - SURVEY 8(d) "C4 pipeline", 39 algorithmic bytes per pixel.  A step is one pass
over one clip per rank; clips are independent, so ranks share nothing (weak scaling, no
collective on the data path; torch.distributed only for the barrier and the max-over-ranks).

    python bench.py --gpus N --steps K --warmup W            # ours
    python bench.py --impl reference --gpus N ...            # the reference's own CPU implementation

Besides the contract's keys the line carries: `value_with_solve` (fresh TPS grids every step, solved on the
host by a worker pool while the previous step's kernels run), `configs` (the other BASELINE.json configs, short
device-timed runs), and in `e2e` the copy-only ceiling of the same byte volume.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, CLIP, NCTRL = 1080, 1920, 64, 5
BYTES_PER_PX = {"c4": 39, "c3": 23, "c2": 27, "c5": 17}
METRIC = "1080p frames/s (warp+TPS+composite)"
WORKLOAD = ("C4 1080p: flow warp + fwd/bwd occlusion mask + TPS (25 control points, fresh grid per frame) "
            "+ composite, BGRA uint8 in, float32x4 out, clip of 64 frames per GPU per step")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """DRAM bytes of the kernels of one 64-frame launch from the committed ncu --set full summary."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------- clocks sampling

class ClockSampler:
    FIELDS = ("uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid, self.proc = uuid, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8 or (self.uuid and self.uuid not in p[0]):
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------- host placement

def _cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def _gpu_numa_node(torch, idx):
    try:
        pr = torch.cuda.get_device_properties(idx)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            return int(f.read())
    except Exception:
        return -1


def place_rank(torch, local, world):
    """Pin this rank to its own slice of the host cores - those of its GPU's NUMA node when sysfs tells, split
    among the ranks whose GPUs sit on the same node - so that N ranks do not oversubscribe one another and
    pinned buffers are allocated NUMA-locally.  Returns a description for the bench line."""
    try:
        avail = sorted(os.sched_getaffinity(0))
    except Exception:
        return {"cores": None}
    nodes = [_gpu_numa_node(torch, g) for g in range(world)] if world > 1 else [_gpu_numa_node(torch, local)]
    node = nodes[local] if world > 1 else nodes[0]
    cores = avail
    if node >= 0:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                local_cores = [c for c in _cpulist(f.read()) if c in avail]
            if local_cores:
                cores = local_cores
        except Exception:
            node = -1
    if world > 1:
        peers = [g for g in range(world) if nodes[g] == node] if node >= 0 else list(range(world))
        k, share = peers.index(local), len(peers)
        per = max(1, len(cores) // share)
        mine = cores[k * per:(k + 1) * per] or cores
        try:
            os.sched_setaffinity(0, mine)
            cores = mine
        except Exception:
            pass
    return {"cores": len(cores), "numa_node": node}


# ------------------------------------------------------------------------- synthetic inputs

def smooth(torch, gen, n, h, w, cell, device):
    import torch.nn.functional as F
    g = torch.randn((n, 1, h // cell + 2, w // cell + 2), generator=gen, device=device)
    return F.interpolate(g, size=(h, w), mode="bicubic", align_corners=True)[:, 0]


def make_clip(torch, seed, n, h, w, device):
    """Synthetic clip on the device: white-noise colour, blobby alpha, smooth ~8 px
    backward flow, forward = -backward + 0.5 px noise with a 10 % rectangle offset by +25 px."""
    gen = torch.Generator(device=device); gen.manual_seed(seed)
    fg = torch.randint(0, 256, (n, h, w, 4), generator=gen, device=device, dtype=torch.uint8)
    fg[..., 3] = (128 + 384 * smooth(torch, gen, n, h, w, 32, device)).clamp(0, 255).to(torch.uint8)
    amp = 8.0 * (w / 1920.0)
    fb = torch.stack([smooth(torch, gen, n, h, w, 32, device), smooth(torch, gen, n, h, w, 32, device)], -1) * amp
    ff = -fb + 0.5 * torch.stack([smooth(torch, gen, n, h, w, 32, device), smooth(torch, gen, n, h, w, 32, device)], -1)
    rh, rw = int(h * 0.316), int(w * 0.316)
    ff[:, h // 5:h // 5 + rh, w // 4:w // 4 + rw] += 25.0
    bg = (128 + 64 * torch.stack([smooth(torch, gen, 4, h, w, 16, device) for _ in range(3)], -1)).clamp(0, 255).to(torch.uint8)
    return fg.contiguous(), fb.float().contiguous(), ff.float().contiguous(), bg.contiguous()


def make_grids(vm, seed, n, h, w, n_ctrl=NCTRL):
    """n (regular grid, deformed grid) pairs: np.random.seed(s); augmentation.deform_grid(h, w, n_ctrl) -
    the reference's own generator (augmentation.py:24-41, the product's host copy of it), SURVEY 8d."""
    import numpy as np
    state = np.random.get_state()
    try:
        out = []
        for k in range(n):
            np.random.seed(seed * 1000 + k)
            out.append(vm.augmentation.deform_grid(h, w, n_ctrl))
        return out
    finally:
        np.random.set_state(state)


# ------------------------------------------------------------------------------- our arm

def run_ours(args):
    import numpy as np
    import torch
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    placement = place_rank(torch, local, world)
    torch.set_num_threads(1)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout when the communicator is created (NCCL_DEBUG=WARN/VERSION on this
        # pool): send fd 1 to stderr while that happens, so that rank 0's stdout is the ONE JSON line of the contract
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    vm = ge.load_package()
    P, Nt = vm.pipeline, vm._native
    lib = Nt.load()
    peak, peak_src = measured_peak()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ev = lambda: torch.cuda.Event(enable_timing=True)
    ncores = placement.get("cores") or os.cpu_count() or 2
    solve_workers = max(1, min(8, ncores if ncores <= 4 else ncores - 1))       # the launching thread mostly waits: few cores -> use them all
    pool = P.SolverPool(solve_workers)

    fg, fb, ff, bg = make_clip(torch, 1234 + rank, CLIP, H, W, dev)
    n_sets = 4
    grid_sets = [make_grids(vm, 10 * rank + 1 + s, CLIP, H, W) for s in range(n_sets)]
    grids = grid_sets[0]
    ctrl, coef = P.solve_grids(grids, dev, pool=pool)
    plan = P.get_plan((0, 0, H, W), 2, dev)
    out = torch.empty((CLIP, H, W, 4), dtype=torch.float32, device=dev)
    status = Nt.new_status(dev)
    scratch = torch.empty(lib.vm_fused_scratch_bytes(CLIP, H, W), dtype=torch.uint8, device=dev)

    def launch(c, k):
        Nt.check(lib.vm_flow_tps_composite_bgra(
            fg.data_ptr(), fb.data_ptr(), ff.data_ptr(), bg.data_ptr(), bg.shape[0], c.data_ptr(),
            k.data_ptr(), NCTRL * NCTRL, plan.nx, plan.ny, plan.step_x, plan.step_y, plan.rows.data_ptr(),
            plan.cols.data_ptr(), CLIP, H, W, out.data_ptr(), scratch.data_ptr(), status.data_ptr(),
            torch.cuda.current_stream().cuda_stream))

    marks = []

    def step(timed):
        e1, e2 = ev(), ev()
        e1.record()
        launch(ctrl, coef)
        e2.record()
        if timed:
            marks.append((e1, e2))

    for _ in range(args.warmup):
        step(False)
    stage_info = kernel_stages(vm, lib, step, torch)
    launches0 = lib.vm_lean_launch_count() + lib.vm_fuse_launch_count()
    sampler = ClockSampler(str(torch.cuda.get_device_properties(local).uuid)) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start(); time.sleep(0.3)
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    barrier()
    launches = int(lib.vm_lean_launch_count() + lib.vm_fuse_launch_count() - launches0)
    ms_total = max_over_ranks(t0.elapsed_time(t1))
    frames = world * CLIP * args.steps
    value = frames / (ms_total / 1e3)
    k_ms = statistics.mean(a.elapsed_time(b) for (a, b) in marks)
    ref_out = out.clone() if CLIP * H * W * 16 < 3e9 else None

    # ---- the same loop with FRESH grids every step: the host solve of step s+1 (worker pool) overlaps the
    # kernels of step s; coefficients go up through pinned memory on the launch stream -----------------------
    ws_steps = max(4, min(args.steps, 40))
    pin = [(torch.empty((CLIP, NCTRL * NCTRL, 2), dtype=torch.float64).pin_memory(),
            torch.empty((CLIP, NCTRL * NCTRL + 3, 2), dtype=torch.float64).pin_memory()) for _ in range(2)]
    dbuf = [(torch.empty_like(ctrl), torch.empty_like(coef)) for _ in range(2)]
    consumed = [ev(), ev()]
    barrier()
    w0 = time.perf_counter()
    s0, s1 = ev(), ev()
    handle = pool.submit(grid_sets[0])
    solve_wait = 0.0
    s0.record()
    for s in range(ws_steps):
        tw = time.perf_counter()
        c_h, k_h = pool.collect(handle)
        solve_wait += time.perf_counter() - tw
        if s + 1 < ws_steps:
            handle = pool.submit(grid_sets[(s + 1) % n_sets])
        b = s & 1
        if s >= 2:
            consumed[b].synchronize()                      # the pinned / device pair of step s-2 is free again
        pin[b][0].copy_(torch.from_numpy(c_h)); pin[b][1].copy_(torch.from_numpy(k_h))
        dbuf[b][0].copy_(pin[b][0], non_blocking=True); dbuf[b][1].copy_(pin[b][1], non_blocking=True)
        launch(dbuf[b][0], dbuf[b][1])
        consumed[b].record()
    s1.record()
    barrier()
    ws_ms = max_over_ranks(s0.elapsed_time(s1))
    ws_wall = max_over_ranks(time.perf_counter() - w0)
    value_with_solve = world * CLIP * ws_steps / (max(ws_ms / 1e3, 0.0) or ws_wall)

    # ---- e2e: same clip from pinned host memory through the public host API ----------------
    e2e_steps = max(1, min(args.steps, 3))
    host = [t.cpu().pin_memory() for t in (fg, fb, ff)]
    bg_h = bg[torch.arange(CLIP) % bg.shape[0]].cpu().pin_memory()
    out_h = torch.empty((CLIP, H, W, 4), dtype=torch.float32).pin_memory()
    P.flow_tps_composite_host(host[0], host[1], host[2], bg_h, grids, out=out_h, pool=pool)      # warm-up
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        P.flow_tps_composite_host(host[0], host[1], host[2], bg_h, grids, out=out_h, pool=pool)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - w0)
    clocks = sampler.stop() if sampler else None     # sampled over the device-timed steps and the end-to-end steps
    runner = next(iter(P._runners.values()))
    h2d, d2h, solve_s = int(runner.h2d_bytes), int(runner.d2h_bytes), float(runner.solve_s)
    same = bool(torch.equal(out_h.to(dev), ref_out)) if ref_out is not None else None
    # copy-only ceiling: the same H2D / D2H volume through the same three streams, no solve, no kernels
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        runner.run(host[0], host[1], host[2], bg_h, grids, out_h, kernels=False)
        torch.cuda.current_stream().synchronize()
    barrier()
    copy_s = max_over_ranks(time.perf_counter() - w0)
    e2e = {"value": world * CLIP * e2e_steps / e2e_s, "unit": "frames/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "copy_only_value": world * CLIP * e2e_steps / copy_s,
           "frac_of_copy_only": copy_s / e2e_s,
           "h2d_GBps_per_rank": h2d * e2e_steps / e2e_s / 1e9, "d2h_GBps_per_rank": d2h * e2e_steps / e2e_s / 1e9,
           "copy_only_h2d_GBps_per_rank": h2d * e2e_steps / copy_s / 1e9,
           "solve_ms_per_step": 1e3 * solve_s, "solve_workers": solve_workers, "host_placement": placement,
           "note": "pinned host clip -> H2D -> kernels -> D2H float32x4, TPS solve on host (worker pool) included; "
                   "copy_only_value = the same copies without solve and kernels, same streams, all ranks at once"}
    del host, bg_h, out_h, ref_out
    cfgs = bench_configs(vm, torch, dev, rank, world, barrier, max_over_ranks, peak, pool)
    pool.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    alg = BYTES_PER_PX["c4"] * H * W * CLIP
    ach = alg / (k_ms / 1e3) / 1e9
    traffic = load_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 in / f64 TPS transform / f32 out", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "frames_per_step_per_gpu": CLIP,
                   "control_points": NCTRL * NCTRL, "sharding": f"clip-sharded replicas x{world}, no collective",
                   "l2": "inputs+outputs 4.8 GB per step >> 126 MB L2 (no flush needed)"},
        "value_with_solve": value_with_solve,
        "with_solve": {"steps": ws_steps, "solve_cores": solve_workers, "ratio_to_value": value_with_solve / value,
                       "host_wait_ms_per_step": 1e3 * solve_wait / ws_steps,
                       "note": "fresh deform_grid per frame and step; np.linalg.pinv systems (reference tps.py:113-119) solved by "
                               f"{solve_workers} worker processes per rank while the previous step's kernels run"},
        "roofline": {"bound": "hbm", "kernel": stage_info["kernel"],
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic.get("traffic_per_launch"), "traffic_source": traffic.get("source"),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                     "step_frac": alg / ((ms_total / args.steps) / 1e3) / 1e9 / peak, "stages": stage_info["stages"],
                     "single_pass_kernel": stage_info.get("single_pass_kernel")},
        "e2e": e2e, "e2e_matches_device_path": same,
        "configs": cfgs,
        "gpu_launches": launches, "clocks": clocks,
        "status_words": [int(v) for v in status.cpu()],
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def kernel_stages(vm, lib, step, torch):
    """Per-kernel durations of one (untimed) step of the default path (lean split pipeline, four kernels) from CUDA
    events recorded inside the library on the launch stream, with the committed ncu DRAM traffic of each kernel
    beside them; and, for comparison, the same step through the single-pass kernel (fused_variant 5: no
    intermediate in HBM, but slower - DESIGN.md 5f)."""
    import ctypes
    Nt, P = vm._native, vm.pipeline
    peak, _ = measured_peak()
    traffic = load_traffic()
    names = traffic.get("kernels") or ["k_lean_coarse<25> (TPS on the coarse grid, float64)", "k_lean_boxes<1> (source box per tile)",
                                       "k_flow_warp_mask_bgra<1,2> (flow warp + consistency mask)",
                                       "k_lean_fine_tm<1,4> (resampling + composite, tensor-map staging)"]
    info = {"kernel": traffic.get("kernel_note") or
            "vm_flow_tps_composite_bgra = k_lean_coarse + k_lean_boxes + k_flow_warp_mask_bgra + k_lean_fine_tm, timed as one unit "
            "(39 B/px is defined for the whole pipeline); the dominant kernel is k_lean_fine_tm, see stages", "stages": []}
    try:
        Nt.set_option("lean_timing", 1)
        rows = []
        for _ in range(3):
            step(False)
            torch.cuda.synchronize()
            buf = (ctypes.c_float * 4)()
            Nt.check(lib.vm_lean_stage_ms(ctypes.cast(buf, ctypes.c_void_p)))
            rows.append(list(buf))
        stage_ms = [statistics.median(col) for col in zip(*rows)]
        stage_bpp = traffic.get("stage_algorithmic_bpp") or [0, 0, 20, 19]
        stage_tr = traffic.get("stage_traffic") or [None] * len(stage_ms)
        total = sum(stage_ms) or 1.0
        for nm, ms_k, bpp, tr in zip(names, stage_ms, stage_bpp, stage_tr):
            rec = {"kernel": nm, "ms": ms_k, "share_of_step": ms_k / total,
                   "algorithmic_GBps": (bpp * H * W * CLIP / (ms_k / 1e3) / 1e9) if bpp else None}
            if tr:
                rec["dram_GBps"] = tr / (ms_k / 1e3) / 1e9
                rec["dram_frac_of_peak"] = rec["dram_GBps"] / peak
            info["stages"].append(rec)
    except Exception as e:          # noqa: BLE001 - stage timing is diagnostic only
        info["stages"].append({"error": str(e)})
    finally:
        Nt.set_option("lean_timing", 0)
    try:
        P.set_fused_variant(5)
        ev = lambda: torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            step(False)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(3):
            step(False)
        b.record()
        torch.cuda.synchronize()
        ms5 = a.elapsed_time(b) / 3
        alg = BYTES_PER_PX["c4"] * H * W * CLIP
        info["single_pass_kernel"] = {"kernel": "k_fuse_c4<25, true, 8> (fused_variant 5: flow warp + mask + TPS + resampling + composite "
                                      "in one persistent warp-specialised kernel, no intermediate in HBM)", "ms": ms5,
                                      "frac": alg / (ms5 / 1e3) / 1e9 / peak, "traffic": traffic.get("single_pass_traffic_per_launch"),
                                      "traffic_source": traffic.get("single_pass_source")}
    except Exception as e:          # noqa: BLE001
        info["single_pass_kernel"] = {"error": str(e)}
    finally:
        P.set_fused_variant(P.DEFAULT_VARIANT)
    return info


def bench_configs(vm, torch, dev, rank, world, barrier, max_over_ranks, peak, pool, iters=8):
    """The other BASELINE.json configs, device-resident, short runs (3 warm-up + `iters` timed launches each,
    CUDA events, max over ranks): C2 1080p x 64, C3 512^2 x 256 (16 control points), C4 4K x 16 per GPU
    (config 4 = 256 frames clip-sharded: 16 clips of 16 frames, each rank times its share), C5 augment_clip."""
    import numpy as np
    P, Nt = vm.pipeline, vm._native
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = []

    def timed(fn, n_it=iters):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(n_it):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n_it

    def rec(name, which, h, w, n, ms, extra=None):
        bpp = BYTES_PER_PX[which]
        gbs = bpp * h * w * n / (ms / 1e3) / 1e9
        r = {"config": name, "height": h, "width": w, "frames_per_launch_per_gpu": n, "ms": ms,
             "fps": world * n / (ms / 1e3), "algorithmic_bytes_per_px": bpp, "frac": gbs / peak}
        r.update(extra or {})
        out.append(r)

    try:
        h, w, n = 1080, 1920, 64
        fg, fb, ff, bg = make_clip(torch, 77 + rank, n, h, w, dev)
        st = Nt.new_status(dev)
        ob = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        oa = torch.empty((n, h, w), dtype=torch.float32, device=dev)
        rec("C2 flow warp + fwd/bwd mask, 1080p x 64", "c2", h, w, n,
            timed(lambda: P.flow_warp_mask(fg, fb, ff, out_bgr=ob, out_alpha=oa, status=st)))
        del ob, oa
        # C5: augmentation.augment_clip on the same clip (host RNG plan + pinv per frame included, alpha statistics reused)
        bgn = bg[torch.arange(n) % bg.shape[0]].contiguous()
        stats = vm.augmentation.alpha_stats(fg)
        np.random.seed(1 + rank)
        barrier()
        c5_iters = 3 * iters                            # a long-running writer: the first calls after a pause wait for the solver pool to wake up
        for _ in range(4):
            vm.augmentation.augment_clip(fg, bgn, stats=stats, pool=pool)
        barrier()
        t = time.perf_counter()
        for _ in range(c5_iters):
            vm.augmentation.augment_clip(fg, bgn, stats=stats, pool=pool)
        barrier()
        ms5 = 1e3 * max_over_ranks(time.perf_counter() - t) / c5_iters
        rec("C5 augmentation.augment_clip (RNG plan + host pinv on the solver pool + TPS + 2 affine passes + illumination), 1080p x 64, wall clock",
            "c5", h, w, n, ms5, {"host_bound": True})
        del fg, fb, ff, bg, bgn
        torch.cuda.empty_cache()

        h, w, n = 512, 512, 256
        fg, fb, ff, bg = make_clip(torch, 78 + rank, n, h, w, dev)
        ctrl, coef = P.solve_grids(make_grids(vm, 500 + rank, n, h, w, 4), dev, pool=pool)
        o = torch.empty((n, h, w, 4), dtype=torch.float32, device=dev)
        rec("C3 TPS (16 control points) + composite, 512x512 x 256", "c3", h, w, n,
            timed(lambda: P.tps_composite(fg, bg, ctrl, coef, out=o, status=st)))
        del fg, fb, ff, bg, o
        torch.cuda.empty_cache()

        h, w, n = 2160, 3840, 16
        fg, fb, ff, bg = make_clip(torch, 79 + rank, n, h, w, dev)
        ctrl, coef = P.solve_grids(make_grids(vm, 600 + rank, n, h, w, 5), dev, pool=pool)
        o = torch.empty((n, h, w, 4), dtype=torch.float32, device=dev)
        rec("C4 flow warp + mask + TPS + composite, 4K, clips of 16 frames (config 4: 256 frames clip-sharded over the ranks)",
            "c4", h, w, n, timed(lambda: P.flow_tps_composite(fg, fb, ff, bg, ctrl, coef, out=o, status=st)),
            {"equivalent_1080p_fps": None})
        out[-1]["equivalent_1080p_fps"] = out[-1]["fps"] * 4
        out[-1]["seconds_for_256_frames"] = 256 / out[-1]["fps"]
        del fg, fb, ff, bg, o
        torch.cuda.empty_cache()
    except Exception as e:          # noqa: BLE001 - the headline must still be printed
        out.append({"error": repr(e)})
    return out


# --------------------------------------------------------------- CPU baseline / reference arm

def _synth_host_frame(seed):
    """One synthetic 1080p frame, flows, grids and background on the host (SURVEY 8d generators)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vm_oracle as O
    return (O.synth_frame(seed, H, W),) + tuple(O.synth_flows(seed, H, W)) + (O.synth_grids(seed, H, W, NCTRL),
                                                                              O.synth_background(seed, H, W))


def _load_reference():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import refshim
    return refshim.load(("reader", "flow", "tps", "augmentation"))


def _reference_frame(seed, threads=None):
    """The C4 pipeline of SURVEY 8d through the UNMODIFIED reference functions (baseline/_ref): warp_bgr,
    warp_img, correct_alpha (the shipped pure-Python loop), warp_image(thin=grids) twice, create_composite_image.
    Returns (seconds, seconds spent in correct_alpha)."""
    import contextlib
    import io
    import numpy as np
    import cv2
    if threads is not None:
        cv2.setNumThreads(threads)
    R = _load_reference()
    fgf, fb, ff, grids, bgf = _synth_host_frame(seed)
    alpha, bgr = fgf[..., 3] / 255., np.ascontiguousarray(fgf[..., :3])
    params = ((0, 0), 0., 1., (W // 2, H // 2))
    t = time.perf_counter()
    b = R["flow"].warp_bgr(bgr, fb)
    a = R["flow"].warp_img(alpha, fb)
    tc = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        a = R["flow"].correct_alpha(fb, ff, a)
    tc = time.perf_counter() - tc
    b2 = R["augmentation"].warp_image(b, params, thin=grids)
    a2 = R["augmentation"].warp_image(a, params, thin=grids)
    R["reader"].create_composite_image(b2, bgf, a2)
    return time.perf_counter() - t, tc


def _port_frame(seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vm_oracle as O
    fgf, fb, ff, grids, bgf = _synth_host_frame(seed)
    t = time.perf_counter()
    O.pipeline_c4(fgf, fb, ff, grids, bgf)
    return time.perf_counter() - t, 0.0


def _have_reference():
    try:
        _load_reference()
        return True
    except Exception:
        return False


def cpu_baseline():
    """Rank 0, N=1: the reference AS SHIPPED (one process, cv2's default thread pool) on one synthetic 1080p frame
    when baseline/_ref is present, else the oracle port on four frames."""
    if _have_reference():
        import cv2
        secs, tc = _reference_frame(10)
        return {"value": 1.0 / secs, "unit": "frames/s", "cores": int(cv2.getNumThreads()), "kind": "reference",
                "sample": "1 synthetic 1080p frame through the unmodified reference modules (baseline/_ref via the A.0 shim), "
                          "one process, cv2 default threads; correct_alpha (pure-Python loop, flow.py:41-48) took "
                          f"{tc:.1f} s = {1e6 * tc / (H * W):.2f} us/px of {secs:.1f} s"}
    frames = 4
    secs = [_port_frame(10 + k)[0] for k in range(frames)]
    return {"value": frames / sum(secs), "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{frames} synthetic 1080p frames through oracle.pipeline_c4 (vectorised NumPy restatement; "
                      "baseline/_ref absent)"}


def _ref_worker(seed):
    return _reference_frame(seed, threads=1)


def run_reference(args):
    """The reference's own CPU implementation with all host cores: one process per core (cv2.setNumThreads(1)),
    one 1080p frame per process and step; before that, two frames as shipped (one process,
    cv2's default pool).  Bounded to ~4 minutes whatever --steps says."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, 64))
    use_ref = _have_reference()
    fn = _ref_worker if use_ref else _port_frame
    kind = "reference" if use_ref else "port"
    shipped = None
    if use_ref:
        import cv2
        rows = [_reference_frame(20 + k) for k in range(2)]
        secs, tc = sum(r[0] for r in rows), sum(r[1] for r in rows)
        shipped = {"frames": 2, "fps": 2 / secs, "s_per_frame": secs / 2, "cv2_threads": int(cv2.getNumThreads()),
                   "correct_alpha_s_per_frame": tc / 2, "correct_alpha_us_per_px": 1e6 * tc / 2 / (H * W)}
    ctx = mp.get_context("spawn")
    times = []
    budget_s, t_start = 170.0, time.perf_counter()
    warm = 0 if use_ref else min(args.warmup, 1)         # a reference step is ~25 s of pure Python: nothing to warm up
    with ctx.Pool(workers) as pool:
        for s in range(warm + args.steps):
            t = time.perf_counter()
            pool.map(fn, [1000 * s + k for k in range(workers)])
            dt = time.perf_counter() - t
            if s >= warm:
                times.append(dt)
            if times and time.perf_counter() - t_start + dt > budget_s:
                break
    total = sum(times)
    steps_run = len(times)
    value = workers * steps_run / total
    sample = (f"{workers} synthetic 1080p frames per step, one per worker process with cv2.setNumThreads(1) (host has {cores} usable cores); "
              + ("unmodified reference modules from baseline/_ref (flow.warp_bgr / warp_img / correct_alpha as shipped, "
                 "augmentation.warp_image(thin=grids) x2, reader.create_composite_image)" if use_ref else
                 "oracle.pipeline_c4 = NumPy restatement (baseline/_ref absent)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": steps_run, "steps_requested": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * total / steps_run, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 in / f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "height": H, "width": W, "frames_per_step": workers},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": workers, "kind": kind, "sample": sample,
                             "as_shipped": shipped},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

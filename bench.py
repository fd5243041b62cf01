"""bench.py - throughput of the B200 video-matting data path.

Headline workload (BASELINE.json metric "1080p frames/s (warp+TPS+composite)"): per GPU, clips
of 64 synthetic 1080p BGRA frames go through flow warp + forward/backward occlusion mask +
thin-plate-spline deformation (25 control points, fresh grid per frame) + composite onto a
background - SURVEY 8(d) "C4 pipeline", 39 algorithmic bytes per pixel.  A step is one pass
over one clip per rank; clips are independent, so ranks share nothing (weak scaling, no
collective on the data path; torch.distributed only for the barrier and the max-over-ranks).

    python bench.py --gpus N --steps K --warmup W            # ours
    python bench.py --impl reference --gpus N ...            # CPU restatement (oracle) arm
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
BYTES_PER_PX = {"c4": 39, "c2": 27}
# dram__bytes_read.sum + dram__bytes_write.sum of the four kernels of one 64-frame launch, from the ncu
# --set full capture summarised in profiles/ (None until measured for the current kernels)
TRAFFIC_PER_LAUNCH = 8486719488
TRAFFIC_SOURCE = "profiles/r01_lean_traffic.txt"
# the same capture per kernel (spline, tile boxes, flow stage, resampling), bytes per 64-frame launch
STAGE_TRAFFIC = [600568960, 133293312, 3683009920, 4069848192]
# what binds each stage: the spline stage has no HBM traffic to speak of, it is bound by the
# float64 pipe - 10.85 DP instructions per (coarse point, control point) evaluation (ncu instruction
# histogram), one DP instruction per 2 cycles and SM sub-partition
STAGE_BOUND = ["fp64 pipe", "latency", "hbm", "hbm + instruction issue"]
DP_PER_EVAL = 10.85
METRIC = "1080p frames/s (warp+TPS+composite)"
WORKLOAD = ("C4 1080p: flow warp + fwd/bwd occlusion mask + TPS (25 control points, fresh grid per frame) "
            "+ composite, BGRA uint8 in, float32x4 out, clip of 64 frames per GPU per step")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
        # "under load" = samples in the upper half of the observed range (idle samples before/after excluded)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


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


def make_grids(O, seed, n, h, w):
    return [O.synth_grids(seed * 1000 + k, h, w, NCTRL) for k in range(n)]


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
    P = vm.pipeline
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vm_oracle as O       # input-grid generator + cpu_baseline checker leg only

    fg, fb, ff, bg = make_clip(torch, 1234 + rank, CLIP, H, W, dev)
    grids = make_grids(O, rank + 1, CLIP, H, W)
    ctrl, coef = P.solve_grids(grids, dev)
    plan = P.get_plan((0, 0, H, W), 2, dev)
    out = torch.empty((CLIP, H, W, 4), dtype=torch.float32, device=dev)
    status = vm._native.new_status(dev)
    lib = vm._native.load()
    scratch = torch.empty(lib.vm_fused_scratch_bytes(CLIP, H, W), dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    def step(timed):
        e1, e2 = ev(), ev()
        e1.record()
        vm._native.check(lib.vm_flow_tps_composite_bgra(
            fg.data_ptr(), fb.data_ptr(), ff.data_ptr(), bg.data_ptr(), bg.shape[0], ctrl.data_ptr(),
            coef.data_ptr(), NCTRL * NCTRL, plan.nx, plan.ny, plan.step_x, plan.step_y, plan.rows.data_ptr(),
            plan.cols.data_ptr(), CLIP, H, W, out.data_ptr(), scratch.data_ptr(), status.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        e2.record()
        if timed:
            marks.append((e1, e2))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    # per-stage durations of one (untimed) step, CUDA events recorded inside the library on the launch stream
    import ctypes
    vm._native.set_option("lean_timing", 1)
    stage_ms = []
    for _ in range(3):
        step(False)
        torch.cuda.synchronize()
        buf = (ctypes.c_float * 4)()
        vm._native.check(lib.vm_lean_stage_ms(ctypes.cast(buf, ctypes.c_void_p)))
        stage_ms.append(list(buf))
    vm._native.set_option("lean_timing", 0)
    stage_ms = [statistics.median(col) for col in zip(*stage_ms)]
    launches0 = lib.vm_lean_launch_count()
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
    launches = int(lib.vm_lean_launch_count() - launches0)
    ms = t0.elapsed_time(t1)
    tms = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    frames = world * CLIP * args.steps
    value = frames / (ms_total / 1e3)
    k_ms = statistics.mean(a.elapsed_time(b) for (a, b) in marks)

    # ---- e2e: same clip from pinned host memory through the public host API ----------------
    e2e_steps = max(1, min(args.steps, 3))
    host = [t.cpu().pin_memory() for t in (fg, fb, ff)]
    bg_h = bg[torch.arange(CLIP) % bg.shape[0]].cpu().pin_memory()
    out_h = torch.empty((CLIP, H, W, 4), dtype=torch.float32).pin_memory()
    P.flow_tps_composite_host(host[0], host[1], host[2], bg_h, grids, out=out_h)      # warm-up
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        P.flow_tps_composite_host(host[0], host[1], host[2], bg_h, grids, out=out_h)
    barrier()
    e2e_s = time.perf_counter() - w0
    clocks = sampler.stop() if sampler else None     # sampled over the device-timed steps and the end-to-end steps
    te = torch.tensor([e2e_s], device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    runner = next(iter(P._runners.values()))
    e2e = {"value": world * CLIP * e2e_steps / float(te.item()), "unit": "frames/s",
           "h2d_bytes_per_step": int(runner.h2d_bytes), "d2h_bytes_per_step": int(runner.d2h_bytes),
           "steps": e2e_steps, "note": "pinned host clip -> H2D -> kernels -> D2H float32x4, TPS solve on host included"}
    same = bool(torch.equal(out_h.to(dev), out))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    alg = BYTES_PER_PX["c4"] * H * W * CLIP
    ach = alg / (k_ms / 1e3) / 1e9
    names = ["k_lean_coarse<25> (TPS on the coarse grid, float64)", "k_lean_boxes<1> (source box per tile)",
             "k_flow_warp_mask_bgra<1,2> (flow warp + consistency mask)", "k_lean_fine<1,4> (resampling + composite)"]
    # algorithmic bytes each stage moves per pixel (SURVEY 8d layouts; intermediates are not algorithmic)
    stage_bpp = [0, 0, 8 + 8 + 4, 3 + 16]
    stages = [{"kernel": nm, "ms": ms_k, "share_of_step": ms_k / sum(stage_ms), "bound": bound,
               "algorithmic_GBps": (bpp * H * W * CLIP / (ms_k / 1e3) / 1e9) if bpp else None,
               "dram_GBps": tr / (ms_k / 1e3) / 1e9, "dram_frac_of_peak": tr / (ms_k / 1e3) / 1e9 / peak}
              for nm, ms_k, bpp, tr, bound in zip(names, stage_ms, stage_bpp, STAGE_TRAFFIC, STAGE_BOUND)]
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_hz = (((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)) * 1e6
    evals = CLIP * (H // 2) * (W // 2) * NCTRL * NCTRL
    stages[0]["fp64_pipe_frac"] = (evals / 32 * DP_PER_EVAL) / ((stage_ms[0] / 1e3) * sm_count * 4 * 0.5 * sm_hz)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 in / f64 TPS transform / f32 out", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "frames_per_step_per_gpu": CLIP,
                   "control_points": NCTRL * NCTRL, "sharding": f"clip-sharded replicas x{world}, no collective",
                   "l2": "inputs+outputs 4.8 GB per step >> 126 MB L2 (no flush needed)"},
        "roofline": {"bound": "hbm",
                     "kernel": "vm_flow_tps_composite_bgra = k_lean_coarse + k_lean_boxes + k_flow_warp_mask_bgra + "
                               "k_lean_fine, timed as one unit (39 B/px is defined for the whole pipeline); the "
                               "dominant kernel is k_lean_fine, see stages",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": TRAFFIC_PER_LAUNCH,
                     "traffic_source": TRAFFIC_SOURCE,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                     "step_frac": alg / ((ms_total / args.steps) / 1e3) / 1e9 / peak, "stages": stages},
        "e2e": e2e, "e2e_matches_device_path": same,
        "gpu_launches": launches, "clocks": clocks,
        "status_words": [int(v) for v in status.cpu()],
    }
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(O, frames=4)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------- CPU baseline / reference arm

def _cpu_frame(seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vm_oracle as O
    fgf = O.synth_frame(seed, H, W)
    b, f = O.synth_flows(seed, H, W)
    g = O.synth_grids(seed, H, W, NCTRL)
    bgf = O.synth_background(seed, H, W)
    t = time.perf_counter()
    O.pipeline_c4(fgf, b, f, g, bgf)
    return time.perf_counter() - t


def cpu_baseline(O, frames=2):
    """Oracle (NumPy restatement of the reference pipeline) on this box's host, one process."""
    secs = [_cpu_frame(10 + k) for k in range(frames)]
    return {"value": frames / sum(secs), "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{frames} synthetic 1080p frames through oracle.pipeline_c4 (vectorised NumPy restatement; "
                      "the shipped reference adds a pure-Python per-pixel loop in correct_alpha, ~10 us/px)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    ctx = mp.get_context("spawn")
    times = []
    # bounded sample: every step is `workers` frames (one per process); at most ~150 s of wall clock in total
    budget_s, t_start = 150.0, time.perf_counter()
    warm = min(args.warmup, 1)
    with ctx.Pool(workers) as pool:
        for s in range(warm + args.steps):
            t = time.perf_counter()
            pool.map(_cpu_frame, [1000 * s + k for k in range(workers)])
            dt = time.perf_counter() - t
            if s >= warm:
                times.append(dt)
            if times and time.perf_counter() - t_start + dt > budget_s:
                break
    total = sum(times)
    steps_run = len(times)
    value = workers * steps_run / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": steps_run, "steps_requested": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * total / steps_run, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 in / f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "height": H, "width": W, "frames_per_step": workers},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": workers, "kind": "port",
                             "sample": f"{workers} 1080p frames per step, one per worker process (host has {cores} cores); "
                                       "oracle.pipeline_c4 = NumPy restatement of the reference pipeline (the Python "
                                       "reference and its cv2/scipy wheels cannot be compiled into oracle/_ref)"},
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

#!/usr/bin/env python
"""Benchmark of the two hot paths on B200 (contract: see the task prompt).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload rt_cornell_4k|rt_tess100k_4k|rast_cornell_4k|rast_soup_4k]

One JSON line on stdout (rank 0).  The headline line is the raytracer on
BASELINE config 3 (Cornell box at 3840x2160, 9 spp, shadow rays, rows banded over
the GPUs): metric Mrays/s = (primary + shadow rays) / s, whole job.  The same
line carries the rasteriser figure (frames/s on BASELINE config 4, the 1M-triangle
soup at 3840x2160) under "raster" when --workload is left at its default.

A "step" is one frame.  `value` times the device-resident path (scene already in
HBM) with CUDA events on the launching stream; `e2e` times the host-pointer C-ABI
call (H2D of the scene from pinned memory, render, D2H of the packed framebuffer
into pinned memory).  N > 1: one process per GPU (torchrun), each rank renders a
band of rows; the bands are assembled on rank 0 by an NCCL gather inside the
timed region.

--impl reference times the reference's own CPU renderer (oracle/_ref, the
unmodified reference compiled as a library) on the box's host cores, on a bounded
sample of rows of the same frame, with one process per core.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOP_PER_RT_TEST = 39.0          # always-executed stage of one ray-triangle test
RAST_BYTES_PER_PIXEL = 16.0      # RGB f32 + depth f32 written once
RAST_BYTES_PER_TRI = 84.0        # input triangle read once

WORKLOADS = {
    # name: (kind, W, H, focal)
    "rt_cornell_4k": ("rt", 3840, 2160, 2160.0),
    "rt_tess100k_4k": ("rt", 3840, 2160, 2160.0),
    "rt_cornell_default": ("rt", 320, 256, 256.0),
    "rast_cornell_4k": ("rast", 3840, 2160, 1536.0),
    "rast_soup_4k": ("rast", 3840, 2160, 1536.0),
    "rast_cornell_default": ("rast", 900, 720, 512.0),
}
RT_CAM = (0.0, 0.0, -3.0, 1.0)
RT_LIGHTS = [((0.0, -0.5, -0.7, 1.0), (14.0, 14.0, 14.0))]
RAST_CAM = (0.0, 0.0, -3.001, 1.0)
RAST_LIGHT = dict(pos=(0.0, -0.5, 0.0, 1.0), power=(20.0, 20.0, 20.0), indirect=(0.2, 0.2, 0.2))


def captured_traffic(workload):
    """DRAM bytes per launch of the workload's dominant kernel, from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "r01", "traffic.json")
    try:
        with open(path) as f:
            t = json.load(f).get(workload)
        return (t["traffic_bytes"], t["kernel"]) if t else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores
# ------------------------------------------------------------------------------------------
def _ref_rt_rows_worker(args):
    """Renders rows [y0, y0+h) of the W x H frame with the reference's own Draw.  The
    crop is expressed through the translation column of R (dir = R * (x, y, f, 1),
    raytracer/Source/skeleton.cpp:126-128), exact for an identity rotation."""
    import helpers as h
    W, H, focal, y0, hh, tris_b, sph_b = args
    tris = np.frombuffer(tris_b, h.RT_TRI) if tris_b is not None else None
    sph = np.frombuffer(sph_b, h.RT_SPHERE) if sph_b is not None else None
    R = h.identity_R()
    R[13] = float(y0 - H // 2 + hh // 2)
    t0 = time.perf_counter()
    h.ref_rt_draw(W, hh, focal, h.f32(*RT_CAM), R, h.lights_array(RT_LIGHTS), tris, sph)
    return time.perf_counter() - t0


def _oracle_rt_count_worker(args):
    import helpers as h
    W, H, focal, y0, hh, tris, sph = args
    o = h.oracle_rt_render(W, H, focal, h.f32(*RT_CAM), h.identity_R(), h.lights_array(RT_LIGHTS),
                           np.frombuffer(tris, h.RT_TRI), np.frombuffer(sph, h.RT_SPHERE), y0, y0 + hh, want=())
    return o["primary"] + o["shadow"]


def sample_windows(H, n_windows, rows_each):
    """Evenly spaced row windows covering the frame top to bottom."""
    n_windows = max(1, min(n_windows, H // rows_each))
    if n_windows == 1:
        return [((H - rows_each) // 2, rows_each)]      # the middle of the frame
    step = (H - rows_each) / max(1, n_windows - 1) if n_windows > 1 else 0
    return [(int(round(i * step)), rows_each) for i in range(n_windows)]


def run_reference(args, workload):
    import helpers as h
    from multiprocessing import get_context
    kind, W, H, focal = WORKLOADS[workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    b200 = importlib.import_module("computer-graphics_b200")
    if kind != "rt":
        return run_reference_rast(args, workload, b200, cores)
    if workload == "rt_tess100k_4k":
        tris, sph = b200.scene_cornell_rt_tessellated(60)
        rows_each, per_proc = 1, 1          # ~0.9 s per row per core at 100 800 triangles
    else:
        tris, sph = b200.scene_cornell_rt()
        rows_each, per_proc = 16, 2
    wins = sample_windows(H, cores * per_proc, rows_each)
    jobs = [(W, H, focal, y0, hh, tris.tobytes(), sph.tobytes()) for (y0, hh) in wins]
    ctx = get_context("fork")
    with ctx.Pool(cores) as pool:
        rays = sum(pool.map(_oracle_rt_count_worker, jobs))   # untimed: ray count of the sample
        times = []
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_rt_rows_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = rays / (ms * 1e-3) / 1e6
    sample = f"{len(wins)} windows x {rows_each} rows of the {W}x{H} frame ({rays} rays) per step"
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "width": W, "height": H, "focal": focal, "spp": 9,
                       "triangles": int(len(tris)), "spheres": int(len(sph))},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def run_reference_rast(args, workload, b200, cores):
    """Rasteriser reference arm: whole reference Draw (geometry + clip + triangle loop +
    post) at the workload's resolution; single-threaded by construction, so `cores`
    independent frames are rendered concurrently and frames/s is their aggregate."""
    import helpers as h
    from multiprocessing import get_context
    kind, W, H, focal = WORKLOADS[workload]
    n_proc = max(1, min(cores, 8))      # each process owns ~365 MB of static frame buffers at 4K
    warmup, steps = min(args.warmup, 1), min(args.steps, 3)   # a soup frame takes seconds on a core
    ctx = get_context("fork")
    with ctx.Pool(n_proc) as pool:
        times = []
        for step in range(warmup + steps):
            t0 = time.perf_counter()
            pool.map(_ref_rast_worker, [(workload,)] * n_proc, chunksize=1)
            dt = time.perf_counter() - t0
            if step >= warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_proc / (ms * 1e-3)
    line = {"impl": "reference", "metric": "frames/s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "width": W, "height": H, "focal": focal},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": n_proc, "kind": "reference",
                             "sample": f"{n_proc} whole frames per step, one per process"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


_rast_scene_cache = {}


def rast_scene(b200, workload):
    if workload not in _rast_scene_cache:
        if workload == "rast_soup_4k":
            _rast_scene_cache[workload] = (b200.scene_soup_rast(1_000_000), np.zeros(0, b200.RAST_TRI))
        else:
            _rast_scene_cache[workload] = b200.scene_cornell_rast()
    return _rast_scene_cache[workload]


def _ref_rast_worker(args):
    import helpers as h
    (workload,) = args
    kind, W, H, focal = WORKLOADS[workload]
    b200 = importlib.import_module("computer-graphics_b200")
    room, boxes = rast_scene(b200, workload)
    lib = h.ref_lib(h.ref_rast_name(W, H))
    n = h.c_i(0)
    t0 = time.perf_counter()
    rc = lib.ref_rast_draw(h.c_f(focal), h.ptr(h.f32(*RAST_CAM)), h.ptr(h.identity_R()), h.ptr(h.f32(*RAST_LIGHT["pos"])),
                           h.ptr(h.f32(*RAST_LIGHT["power"])), h.ptr(h.f32(*RAST_LIGHT["indirect"])),
                           h.ptr(room), h.c_i(len(room)), h.ptr(boxes), h.c_i(len(boxes)),
                           None, None, None, None, None, None, None, None, h.c_i(0), None, None)
    assert rc == 0
    return time.perf_counter() - t0


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def cpu_baseline_rt(b200, r, workload, W, H, focal, tris, sph):
    """The unmodified reference (oracle/_ref), one core, on a bounded sample of rows of
    the same frame; the sample's ray count comes from the (parity-tested) device counters."""
    import helpers as h
    if not h.have_ref("libref_rt.so"):
        return {"value": None, "unit": "Mrays/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref missing"}
    if workload == "rt_tess100k_4k":
        wins = sample_windows(H, 1, 1)      # one row of 3840 pixels: ~4.7e9 ray-triangle tests, about a minute on one core
    else:
        wins = sample_windows(H, 30, 16)
    cam = b200.make_camera(RT_CAM, focal, h.identity_R(), W, H)
    rays, secs = 0, 0.0
    for (y0, hh) in wins:
        r.render_raytrace(tris, sph, cam, RT_LIGHTS, y0, y0 + hh, want=())
        st = r.stats()
        rays += st["primary_rays"] + st["shadow_rays"]
        secs += _ref_rt_rows_worker((W, H, focal, y0, hh, tris.tobytes(), sph.tobytes()))
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "reference",
            "sample": f"{len(wins)} windows x {wins[0][1]} rows of the {W}x{H} frame, {rays} rays, {secs:.1f} s"}


def cpu_baseline_rast(workload):
    import helpers as h
    kind, W, H, focal = WORKLOADS[workload]
    if not h.have_ref(h.ref_rast_name(W, H)):
        return {"value": None, "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref missing"}
    n = 2 if workload == "rast_soup_4k" else 6
    _ref_rast_worker((workload,))
    secs = [_ref_rast_worker((workload,)) for _ in range(n)]
    return {"value": 1.0 / float(np.median(secs)), "unit": "frames/s", "cores": 1, "kind": "reference",
            "sample": f"{n} whole frames (median), reference Draw incl. geometry and post pass"}


def run_b200(args, workloads):
    """Measures the given workloads in one process (one per GPU under torchrun); returns
    the list of result dicts on rank 0."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = [run_b200_one(args, w) for w in workloads]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_b200_one(args, workload):
    import torch
    import torch.distributed as dist
    import helpers as h
    b200 = importlib.import_module("computer-graphics_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    kind, W, H, focal = WORKLOADS[workload]
    peaks, peaks_src = measured_peaks()
    r = b200.Renderer(local)
    # one non-default stream carries the kernels, the NCCL gathers and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    r.set_stream(stream.cuda_stream)
    if "B200_RAST_TILE_LOG2" in os.environ:      # tuning knob, default chosen by the library
        r.set_option(b200.OPT_RAST_TILE_LOG2, int(os.environ["B200_RAST_TILE_LOG2"]))
    row0, row1 = rank * H // world, (rank + 1) * H // world
    rows = row1 - row0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- where each rank's band goes --------------------------------------------------
    # N == 1: plain device buffers.  N > 1: the frame lives on rank 0 and
    #   "p2p_store"  (default) every rank is handed a peer mapping of rank 0's frame
    #                (torch symmetric memory over NVLink) and its kernels store the band
    #                straight into it -- the exchange happens in the epilogue of the render
    #                kernels, followed by one device-side barrier;
    #   "nccl"       each rank renders into a local band, then an NCCL gather to rank 0.
    # The device entry points address outputs as full frames (pixel (x, y) at y*W + x).
    gather, gather_note, symm_handles = "none", None, []
    band_rgb = band_depth = full_rgb = full_depth = None
    if world > 1 and args.gather == "p2p":
        try:
            import torch.distributed._symmetric_memory as symm
            frame_rgb = symm.empty((H, W, 3), dtype=torch.float32, device="cuda")
            frame_depth = symm.empty((H, W), dtype=torch.float32, device="cuda")
            h_rgb = symm.rendezvous(frame_rgb, dist.group.WORLD)
            h_depth = symm.rendezvous(frame_depth, dist.group.WORLD)
            p_rgb, p_depth = int(h_rgb.buffer_ptrs[0]), int(h_depth.buffer_ptrs[0])
            symm_handles = [h_rgb, h_depth, frame_rgb, frame_depth]
            gather = "p2p_store"
        except Exception as e:                      # both branches are GPU paths; say which one ran
            gather_note = f"symmetric memory unavailable ({type(e).__name__}: {e}); NCCL gather used"
    if gather != "p2p_store":
        band_rgb = torch.empty((rows, W, 3), dtype=torch.float32, device="cuda")
        band_depth = torch.empty((rows, W), dtype=torch.float32, device="cuda")
        p_rgb = band_rgb.data_ptr() - row0 * W * 12     # frame origin above the band
        p_depth = band_depth.data_ptr() - row0 * W * 4
        if world > 1:
            gather = "nccl_gather"
            if rank == 0:
                full_rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
                full_depth = torch.empty((H, W), dtype=torch.float32, device="cuda")

    def exchange():
        if gather == "p2p_store":
            symm_handles[0].barrier(channel=0)      # every band has landed in rank 0's frame
        elif gather == "nccl_gather":
            dist.gather(band_rgb, list(full_rgb.view(world, rows, W, 3).unbind(0)) if rank == 0 else None, dst=0)
            dist.gather(band_depth, list(full_depth.view(world, rows, W).unbind(0)) if rank == 0 else None, dst=0)

    result = {}
    pinned = torch.empty(rows * W, dtype=torch.int32).pin_memory()
    if kind == "rt":
        tris, sph = b200.scene_cornell_rt_tessellated(60) if workload == "rt_tess100k_4k" else b200.scene_cornell_rt()
        cam = b200.make_camera(RT_CAM, focal, h.identity_R(), W, H)
        r.rt_upload_scene(tris, sph)

        # N > 1 with the peer-mapped frame: the ranks share the frame by interleaved 16-row blocks
        # (better balanced than contiguous bands); each stores its blocks straight into rank 0's frame
        interleaved = world > 1 and gather == "p2p_store"
        if interleaved:
            r.set_option(b200.OPT_RT_INTERLEAVE_N, world)
            r.set_option(b200.OPT_RT_INTERLEAVE_R, rank)

        def step():
            if interleaved:
                r.rt_render_device(cam, RT_LIGHTS, 0, H, p_rgb, p_depth)
            else:
                r.rt_render_device(cam, RT_LIGHTS, row0, row1, p_rgb, p_depth)
            exchange()

        tris_pin = torch.from_numpy(tris.view(np.uint8).copy()).pin_memory()
        sph_pin = torch.from_numpy(sph.view(np.uint8).copy()).pin_memory()
        tris_h = tris_pin.numpy().view(b200.RT_TRI)
        sph_h = sph_pin.numpy().view(b200.RT_SPHERE)

        def step_e2e():
            r.draw_raytrace_band(tris_h, sph_h, cam, RT_LIGHTS, row0, row1, pinned.data_ptr())

        h2d = tris.nbytes + sph.nbytes
        d2h = rows * W * 4
    else:
        room, boxes = rast_scene(b200, workload)
        cam = b200.make_camera(RAST_CAM, focal, h.identity_R(), W, H)
        L = b200.make_rast_light(RAST_LIGHT["pos"], RAST_LIGHT["power"], RAST_LIGHT["indirect"])
        r.rast_upload_scene(room, boxes)
        # frames sized from the last verified frame instead of mid-frame read-backs; the
        # verification (b200_synchronize) is part of the step, so a frame that had to be
        # rendered twice would be timed twice
        r.set_option(b200.OPT_RAST_PIPELINED, 0 if args.rast_sync else 1)
        respec0 = r.stats()["respeculated"]

        def step():
            r.rast_draw_device(cam, L, row0, row1, p_rgb, p_depth)
            r.synchronize()
            exchange()

        room_pin = torch.from_numpy(room.view(np.uint8).copy()).pin_memory()
        boxes_pin = torch.from_numpy(boxes.view(np.uint8).copy() if len(boxes) else np.zeros(84, np.uint8)).pin_memory()
        room_h = room_pin.numpy().view(b200.RAST_TRI)
        boxes_h = boxes_pin.numpy().view(b200.RAST_TRI)[: len(boxes)]

        def step_e2e():
            r.draw_raster_band(room_h, boxes_h, cam, L, row0, row1, pinned.data_ptr())

        h2d = room.nbytes + boxes.nbytes
        d2h = rows * W * 4

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        total_ms, launches, last_stats = 0.0, 0, None
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)                       # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            last_stats = r.stats()
            launches += last_stats["kernel_launches"]
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, launches, last_stats, clocks, wall

    ms, launches, st, clocks, wall = timed(step, args.steps, max(args.warmup, 3))
    # whole-job units per step (sum over ranks of what each rank processed)
    if kind == "rt":
        units = torch.tensor([st["primary_rays"] + st["shadow_rays"], st["prim_tests"]], dtype=torch.float64, device="cuda")
    else:
        units = torch.tensor([1.0 / world, float(st["fragments"])], dtype=torch.float64, device="cuda")
    kernel_ms = torch.tensor([st["gpu_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    ms_e2e, _, _, _, _ = timed(step_e2e, max(2, args.steps // 2), 3)

    if kind == "rt":
        rays, tests = float(units[0]), float(units[1])
        value, unit, metric = rays / (ms * 1e-3) / 1e6, "Mrays/s", "Mrays/s"
        e2e_value = rays / (ms_e2e * 1e-3) / 1e6
        fp32_peak = r.measure_fp32_peak()
        achieved = tests * FLOP_PER_RT_TEST / (float(kernel_ms) * 1e-3) / 1e12 / world   # per GPU
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak, "traffic": None,
                    "kernel": "rt_prep_planes_kernel + rt_filtered_kernel", "kernel_ms": float(kernel_ms),
                    "peak_source": "FFMA microbenchmark run in this process (b200_measure_fp32_peak)",
                    "algorithmic_work": f"{FLOP_PER_RT_TEST:.0f} flop per ray-primitive test x {tests:.0f} tests",
                    "exact_evals": st["exact_evals"]}
        config = {"workload": workload, "width": W, "height": H, "focal": focal, "spp": 9,
                  "triangles": int(len(tris)), "spheres": int(len(sph)), "lights": len(RT_LIGHTS),
                  "rays_per_frame": rays,
                  "parallelism": (f"interleaved 16-row blocks x{world}" if world > 1 and gather == "p2p_store" else f"row bands x{world}"),
                  "gather": gather, "gather_note": gather_note,
                  "l2": "flushed between timed steps (256 MiB write); outputs 133 MB > L2"}
    else:
        frames = 1.0
        value, unit, metric = frames / (ms * 1e-3), "frames/s", "frames/s"
        e2e_value = frames / (ms_e2e * 1e-3)
        n_in = len(room) + len(boxes)
        bytes_alg = W * H * RAST_BYTES_PER_PIXEL + n_in * RAST_BYTES_PER_TRI
        achieved = bytes_alg / world / (float(kernel_ms) * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": None, "kernel": "whole raster frame (all kernels)",
                    "kernel_ms": float(kernel_ms), "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks_src})",
                    "algorithmic_work": f"{bytes_alg / 1e6:.1f} MB per frame (16 B/pixel + 84 B/triangle)"}
        config = {"workload": workload, "width": W, "height": H, "focal": focal, "triangles_in": int(n_in),
                  "fragments_per_frame": float(units[1]), "parallelism": f"row bands x{world}", "gather": gather, "gather_note": gather_note,
                  "pipelined": not args.rast_sync, "frames_rendered_twice": int(r.stats()["respeculated"] - respec0),
                  "l2": "flushed between timed steps (256 MiB write)"}

    roofline["traffic"], roofline["traffic_kernel"] = captured_traffic(workload)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r.set_stream(None)
        cpu = cpu_baseline_rt(b200, r, workload, W, H, focal, tris, sph) if kind == "rt" else cpu_baseline_rast(workload)

    if rank == 0:
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d) * world,
                        "d2h_bytes_per_step": int(d2h) * world, "ms_per_step": ms_e2e,
                        "call": "draw_raytrace_band" if kind == "rt" else "draw_raster_band"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        result = line
    r.set_stream(None)
    r.close()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="default", choices=["default"] + sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rast-sync", action="store_true",
                    help="rasteriser: read list/table sizes back mid-frame instead of pipelined frames")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: peer stores into rank 0's frame (default) or an NCCL gather")
    args = ap.parse_args()
    # default: the headline raytracer line (BASELINE config 3) carrying the rasteriser figure
    # (BASELINE config 4) as a nested "raster" object
    workloads = ["rt_cornell_4k", "rast_soup_4k"] if args.workload == "default" else [args.workload]
    if args.impl == "reference":
        lines = [run_reference(args, w) for w in workloads]
    else:
        lines = run_b200(args, workloads)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    line = lines[0]
    if len(lines) > 1 and lines[1]:
        keep = ("metric", "value", "unit", "ms_per_step", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "steps", "warmup")
        line["raster"] = {k: lines[1][k] for k in keep if k in lines[1]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

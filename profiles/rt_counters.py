"""Diagnostic: per-level counts of the RT filter hierarchy (library built with
-DRT_PROFILE_COUNTERS).  Usage: python profiles/rt_counters.py [workload]"""
import ctypes
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import helpers as h  # noqa: E402
import bench  # noqa: E402

b = importlib.import_module("computer-graphics_b200")
workload = sys.argv[1] if len(sys.argv) > 1 else "rt_cornell_4k"
kind, W, H, f = bench.WORKLOADS[workload]
r = b.Renderer(0)
tris, sph = b.scene_cornell_rt_tessellated(60) if workload == "rt_tess100k_4k" else b.scene_cornell_rt()
cam = b.make_camera(bench.RT_CAM, f, h.identity_R(), W, H)
r.rt_upload_scene(tris, sph)
rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
r.rt_render_device(cam, bench.RT_LIGHTS, 0, H, rgb.data_ptr())
c = (ctypes.c_uint64 * 8)()
assert b.load_library().b200_debug_counters(r.ctx, c) == 0
names = ["shadow rays", "exact evals", "primary L1 tests", "primary L1 passes", "shadow L1 tests", "shadow L1 passes",
         "shadow L2 tests", "shadow L2 non-miss"]
npx = W * H
for n, v in zip(names, c):
    print(f"{n:20s} {v:14d}  {v / npx:8.2f} per pixel")

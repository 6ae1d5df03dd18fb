"""One workload, a few frames through the device-resident path: the command line
profiled with ncu (see profiles/README.md)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import helpers as h  # noqa: E402
import bench  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rt_cornell_4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
b200 = importlib.import_module("computer-graphics_b200")
kind, W, H, focal = bench.WORKLOADS[workload]
r = b200.Renderer(0)
rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
depth = torch.empty((H, W), dtype=torch.float32, device="cuda")
if kind == "rt":
    tris, sph = b200.scene_cornell_rt_tessellated(60) if workload == "rt_tess100k_4k" else b200.scene_cornell_rt()
    cam = b200.make_camera(bench.RT_CAM, focal, h.identity_R(), W, H)
    r.rt_upload_scene(tris, sph)
    for _ in range(frames):
        r.rt_render_device(cam, bench.RT_LIGHTS, 0, H, rgb.data_ptr(), depth.data_ptr())
        st = r.stats()
else:
    room, boxes = bench.scenes_rast(workload)
    cam = b200.make_camera(bench.RAST_CAM, focal, h.identity_R(), W, H)
    L = b200.make_rast_light(bench.RAST_LIGHT["pos"], bench.RAST_LIGHT["power"], bench.RAST_LIGHT["indirect"])
    r.rast_upload_scene(room, boxes)
    r.set_option(b200.OPT_RAST_PIPELINED, 1)      # as bench.py runs it
    for _ in range(frames):
        r.rast_draw_device(cam, L, 0, H, rgb.data_ptr(), depth.data_ptr())
        r.synchronize()
        st = r.stats()
print(workload, st)

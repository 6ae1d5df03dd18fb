"""Ordered-tile raster path: frame time for the tile sizes B200_OPT_RAST_TILE_LOG2 offers.
python profiles/rast_tile_probe.py <workload>"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as h, bench
b = importlib.import_module("computer-graphics_b200")
w = sys.argv[1]
kind, W, H, f = bench.WORKLOADS[w]
r = b.Renderer(0)
rgb = torch.empty((H, W, 3), device="cuda"); depth = torch.empty((H, W), device="cuda")
room, boxes = bench.scenes_rast(w)
cam = b.make_camera(bench.RAST_CAM, f, h.identity_R(), W, H)
L = b.make_rast_light(bench.RAST_LIGHT["pos"], bench.RAST_LIGHT["power"], bench.RAST_LIGHT["indirect"])
r.rast_upload_scene(room, boxes)
r.set_option(b.OPT_RAST_PIPELINED, 1)
for ts in (5, 4, 3):
    r.set_option(b.OPT_RAST_TILE_LOG2, ts)
    ms = []
    for i in range(8):
        r.rast_draw_device(cam, L, 0, H, rgb.data_ptr(), depth.data_ptr()); r.synchronize()
        ms.append(r.stats()["gpu_ms"])
    print(w, "tile", 1 << ts, "gpu_ms", round(min(ms[3:]), 4))

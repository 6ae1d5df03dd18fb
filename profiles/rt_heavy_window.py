"""The long-running patch of the 100 800-triangle frame on its own: a 64 x 32 pixel window of the 4K frame around
pixel (1952, 1400), rendered through the window trick of bench.py (same rays as in the full frame).  Profile with
ncu --set full --import-source on to see where a heavy warp's instructions go."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as h, bench
b = importlib.import_module("computer-graphics_b200")
W, H, f = 3840, 2160, 2160.0
x0, y0, w, hh = (int(sys.argv[1]), int(sys.argv[2]), 64, 32) if len(sys.argv) > 2 else (1936, 1392, 64, 32)
r = b.Renderer(0)
tris, sph = bench.scenes_rt("rt_tess100k_4k")
cam = b.make_camera(bench.RT_CAM, f, bench.window_R(W, H, x0, y0, w, hh), w, hh)
r.rt_upload_scene(tris, sph)
rgb = torch.empty((hh, w, 3), device="cuda"); depth = torch.empty((hh, w), device="cuda")
for _ in range(3):
    r.rt_render_device(cam, bench.RT_LIGHTS, 0, hh, rgb.data_ptr(), depth.data_ptr()); st = r.stats()
print("window", (x0, y0, w, hh), "gpu_ms", st["gpu_ms"], "exact evals per pixel", st["exact_evals"] / (w * hh))

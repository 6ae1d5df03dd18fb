"""Diagnostic: the plain Cornell box through the GRID kernel (B200_OPT_RT_GRID = 1), for ncu."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import helpers as h  # noqa: E402
import bench  # noqa: E402
b = importlib.import_module("computer-graphics_b200")
r = b.Renderer(0)
W, H, f = 3840, 2160, 2160.0
rgb = torch.empty((H, W, 3), device="cuda"); depth = torch.empty((H, W), device="cuda")
cam = b.make_camera(bench.RT_CAM, f, h.identity_R(), W, H)
tris, sph = b.scene_cornell_rt()
r.rt_upload_scene(tris, sph)
r.set_option(b.OPT_RT_GRID, 1)
for _ in range(2):
    r.rt_render_device(cam, bench.RT_LIGHTS, 0, H, rgb.data_ptr(), depth.data_ptr()); r.synchronize()
print(r.stats())

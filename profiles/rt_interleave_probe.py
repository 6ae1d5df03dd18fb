"""One GPU playing rank r of N in turn (B200_OPT_RT_INTERLEAVE_N/R): device time of each rank's share
of the frame -- the balance of the interleaved 16-row blocks and the part of the frame that does
not shrink with N.  python profiles/rt_interleave_probe.py <workload> <N>"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers as h, bench
b = importlib.import_module("computer-graphics_b200")
w = sys.argv[1]; N = int(sys.argv[2])
kind, W, H, f = bench.WORKLOADS[w]
r = b.Renderer(0)
tris, sph = bench.scenes_rt(w)
cam = b.make_camera(bench.RT_CAM, f, h.identity_R(), W, H)
r.rt_upload_scene(tris, sph)
rgb = torch.empty((H, W, 3), device="cuda"); depth = torch.empty((H, W), device="cuda"); argb = torch.empty((H, W), dtype=torch.int32, device="cuda")
for n in (1, N):
    r.set_option(b.OPT_RT_INTERLEAVE_N, n)
    ms = []
    for rank in range(n):
        r.set_option(b.OPT_RT_INTERLEAVE_R, rank)
        for _ in range(3):
            r.rt_render_device(cam, bench.RT_LIGHTS, 0, H, rgb.data_ptr(), depth.data_ptr(), None, argb.data_ptr())
            st = r.stats()
        ms.append(round(st["gpu_ms"], 3))
    print(w, "N", n, "gpu_ms per rank", ms, "sum", round(sum(ms), 3))
r.set_option(b.OPT_RT_INTERLEAVE_N, 1)
ms = []
for rank in range(N):
    a, e = rank * H // N, (rank + 1) * H // N
    for _ in range(3):
        r.rt_render_device(cam, bench.RT_LIGHTS, a, e, rgb.data_ptr(), depth.data_ptr(), None, argb.data_ptr())
        st = r.stats()
    ms.append(round(st["gpu_ms"], 3))
print(w, "contiguous bands", N, "gpu_ms per band", ms, "sum", round(sum(ms), 3))

"""Library multi-GPU context: wall time, band edges and per-device cost of consecutive whole-frame
calls (diagnostics of the adaptive split).  python profiles/multi_frames.py <workload> [n_gpus] [frames]"""
import ctypes, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as h, bench
b = importlib.import_module("computer-graphics_b200")
w = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count(); frames = int(sys.argv[3]) if len(sys.argv) > 3 else 12
kind, W, H, f = bench.WORKLOADS[w]
m = b.Renderer(n_gpus=n)
out = torch.empty(H * W, dtype=torch.int32).pin_memory()
if kind == "rt":
    tris, sph = bench.scenes_rt(w)
    tp = torch.from_numpy(tris.view(np.uint8).copy()).pin_memory(); sp = torch.from_numpy(sph.view(np.uint8).copy()).pin_memory()
    th, sh = tp.numpy().view(b.RT_TRI), sp.numpy().view(b.RT_SPHERE)
    cam = b.make_camera(bench.RT_CAM, f, h.identity_R(), W, H)
    call = lambda: m.draw_raytrace_band(th, sh, cam, bench.RT_LIGHTS, 0, H, out.data_ptr())
else:
    room, boxes = bench.scenes_rast(w)
    rp = torch.from_numpy(room.view(np.uint8).copy()).pin_memory()
    rh = rp.numpy().view(b.RAST_TRI)
    cam = b.make_camera(bench.RAST_CAM, f, h.identity_R(), W, H)
    L = b.make_rast_light(bench.RAST_LIGHT["pos"], bench.RAST_LIGHT["power"], bench.RAST_LIGHT["indirect"])
    call = lambda: m.draw_raster_band(rh, boxes, cam, L, 0, H, out.data_ptr())
edges = (ctypes.c_int * 16)(); cost = (ctypes.c_float * 16)()
for i in range(frames):
    t0 = time.perf_counter(); call(); dt = (time.perf_counter() - t0) * 1e3
    k = m.lib.b200_debug_multi_bands(m.ctx, 0 if kind == "rt" else 1, edges, cost, 16)
    print(f"frame {i}: {dt:7.3f} ms wall, gpu_ms max {m.stats()['gpu_ms']:.3f}; edges {list(edges[:k])} cost {[round(c, 3) for c in cost[:k]]}")
m.close()

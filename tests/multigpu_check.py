"""Launched under torchrun by tests/test_multigpu.py: every rank renders its row band
straight into rank 0's frame (peer stores through a symmetric-memory mapping) and,
separately, into a local band that is gathered with NCCL; rank 0 compares both
assembled frames with a single-GPU render of the whole frame.  Bit-exact."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h  # noqa: E402

b200 = importlib.import_module("computer-graphics_b200")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
r = b200.Renderer(local)
r.set_stream(stream.cuda_stream)

W, H = 640, 360
row0, row1 = rank * H // world, (rank + 1) * H // world
rows = row1 - row0
import torch.distributed._symmetric_memory as symm  # noqa: E402
frame = symm.empty((H, W, 3), dtype=torch.float32, device="cuda")
depth = symm.empty((H, W), dtype=torch.float32, device="cuda")
frame.zero_(); depth.zero_()
hf, hd = symm.rendezvous(frame, dist.group.WORLD), symm.rendezvous(depth, dist.group.WORLD)
band = torch.zeros((rows, W, 3), dtype=torch.float32, device="cuda")
ok = True


def check(name, render_full, render_band):
    """render_band(row0, row1, rgb_ptr, depth_ptr) renders with full-frame addressing."""
    global ok
    frame.zero_(); depth.zero_()
    torch.cuda.synchronize(); dist.barrier()
    render_band(row0, row1, int(hf.buffer_ptrs[0]), int(hd.buffer_ptrs[0]))     # peer stores into rank 0
    hf.barrier(channel=0)
    render_band(row0, row1, band.data_ptr() - row0 * W * 12, None)              # local band + NCCL gather
    parts = [torch.empty_like(band) for _ in range(world)] if rank == 0 else None
    dist.gather(band, parts, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        full_rgb, full_depth = render_full()
        a = np.array_equal(frame.cpu().numpy().view(np.uint32), full_rgb.view(np.uint32))
        b = np.array_equal(depth.cpu().numpy().view(np.uint32), full_depth.view(np.uint32))
        c = np.array_equal(torch.cat(parts).cpu().numpy().view(np.uint32), full_rgb.view(np.uint32))
        print(f"{name}: p2p rgb {a} depth {b}; nccl rgb {c}", flush=True)
        ok = ok and a and b and c


# raytracer
tris, sph = b200.scene_cornell_rt()
cam = b200.make_camera((0, 0, -3, 1), 360.0, h.identity_R(), W, H)
r.rt_upload_scene(tris, sph)


def rt_full():
    o = r.render_raytrace(tris, sph, cam, h.DEFAULT_RT_LIGHTS)
    return o["rgb"], o["depth"]


check("raytracer", rt_full, lambda a, b, p, q: r.rt_render_device(cam, h.DEFAULT_RT_LIGHTS, a, b, p, q))

# rasteriser (whole Draw, Cornell with shadow volumes: ordered-tile path with its band halo)
room, boxes = b200.scene_cornell_rast()
rcam = b200.make_camera(h.DEFAULT_RAST_CAM, 256.0, h.identity_R(), W, H)
L = b200.make_rast_light(h.DEFAULT_RAST_LIGHT["pos"], h.DEFAULT_RAST_LIGHT["power"], h.DEFAULT_RAST_LIGHT["indirect"])
r.rast_upload_scene(room, boxes)


def rast_full():
    o = r.render_raster(room, boxes, rcam, L)
    r.rast_upload_scene(room, boxes)
    return o["rgb"], o["depth"]


check("rasteriser cornell", rast_full, lambda a, b, p, q: r.rast_draw_device(rcam, L, a, b, p, q))

# rasteriser, shadow-free list: scatter/resolve path with its band halo
soup = b200.scene_soup_rast(30000, edge=0.03)
none = np.zeros(0, b200.RAST_TRI)
r.rast_upload_scene(soup, none)


def soup_full():
    o = r.render_raster(soup, none, rcam, L)
    r.rast_upload_scene(soup, none)
    return o["rgb"], o["depth"]


check("rasteriser soup", soup_full, lambda a, b, p, q: r.rast_draw_device(rcam, L, a, b, p, q))

flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTIGPU_OK" if ok else "MULTIGPU_MISMATCH", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)

import os
import sys, os, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import helpers as h, bench
b=importlib.import_module("computer-graphics_b200")
w=sys.argv[1]
kind,W,H,f=bench.WORKLOADS[w]
r=b.Renderer(0)
rgb=torch.empty((H,W,3),device="cuda"); depth=torch.empty((H,W),device="cuda")
room,boxes=bench.scenes_rast(w)
cam=b.make_camera(bench.RAST_CAM,f,h.identity_R(),W,H)
L=b.make_rast_light(bench.RAST_LIGHT["pos"],bench.RAST_LIGHT["power"],bench.RAST_LIGHT["indirect"])
r.rast_upload_scene(room,boxes)
r.set_option(b.OPT_RAST_PIPELINED,int(sys.argv[2]) if len(sys.argv) > 2 else 1)
for i in range(4):
    if i==3: sys.stderr.write("---- last frame\n")
    r.rast_draw_device(cam,L,0,H,rgb.data_ptr(),depth.data_ptr()); r.synchronize()

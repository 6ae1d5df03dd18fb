import os
import sys, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import helpers as h, bench
b=importlib.import_module("computer-graphics_b200")
r=b.Renderer(0)
W,H,f=3840,2160,2160.0
rgb=torch.empty((H,W,3),device="cuda"); depth=torch.empty((H,W),device="cuda")
cam=b.make_camera(bench.RT_CAM,f,h.identity_R(),W,H)
for n in (1,2,3,4,6,8):
    tris,sph=b.scene_cornell_rt_tessellated(n)
    r.rt_upload_scene(tris,sph)
    res=[]
    for mode in (2,1):
        r.set_option(b.OPT_RT_GRID,mode)
        for _ in range(2): r.rt_render_device(cam,bench.RT_LIGHTS,0,H,rgb.data_ptr(),depth.data_ptr()); r.synchronize()
        t=0
        for _ in range(5):
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); r.rt_render_device(cam,bench.RT_LIGHTS,0,H,rgb.data_ptr(),depth.data_ptr()); r.synchronize(); e1.record(); e1.synchronize()
            t+=e0.elapsed_time(e1)
        res.append(t/5)
    print(f"n={n} tris={len(tris)} stream {res[0]:.3f} ms  grid {res[1]:.3f} ms")

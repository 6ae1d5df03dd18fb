import os
import sys, os, importlib, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import helpers as h, bench
b=importlib.import_module("computer-graphics_b200")
w=sys.argv[1]
kind,W,H,f=bench.WORKLOADS[w]
r=b.Renderer(0)
st=torch.cuda.Stream(); torch.cuda.set_stream(st); r.set_stream(st.cuda_stream)
rgb=torch.empty((H,W,3),device="cuda"); depth=torch.empty((H,W),device="cuda")
room,boxes=bench.scenes_rast(w)
cam=b.make_camera(bench.RAST_CAM,f,h.identity_R(),W,H)
L=b.make_rast_light(bench.RAST_LIGHT["pos"],bench.RAST_LIGHT["power"],bench.RAST_LIGHT["indirect"])
r.rast_upload_scene(room,boxes)
flush=torch.empty(256<<20,dtype=torch.uint8,device="cuda")
for pipe in (0,1):
    r.set_option(b.OPT_RAST_PIPELINED,pipe)
    for do_flush in (True,False):
        for _ in range(3):
            r.rast_draw_device(cam,L,0,H,rgb.data_ptr(),depth.data_ptr()); r.synchronize()
        tot=0; gms=0; wall=0
        for _ in range(20):
            if do_flush: flush.fill_(1)
            torch.cuda.synchronize()
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            t0=time.perf_counter()
            e0.record(st); r.rast_draw_device(cam,L,0,H,rgb.data_ptr(),depth.data_ptr()); t1=time.perf_counter(); e1.record(st); e1.synchronize()
            tot+=e0.elapsed_time(e1); gms+=r.stats()["gpu_ms"]; wall+=t1-t0
        print(w,"pipelined",pipe,"flush",do_flush,"ms",round(tot/20,4),"gpu_ms",round(gms/20,4),"host enqueue ms",round(wall/20*1e3,4))

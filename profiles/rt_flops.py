"""Turns the launch lists written by profiles/capture_flops.sh into profiles/<round>/rt_flops.json:
executed FP32 flop (thread-level FADD + FMUL + 2 x FFMA) of ONE whole frame per RT workload, summed over
the frame's launches (the last frame of the capture).
Usage: python profiles/rt_flops.py r02 gpurun_out/flops_rt_cornell_4k.csv [...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def frame_flops(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = next(r for r in rows if r[0] == "ID")
    body = [r for r in rows if r[0].isdigit()]
    name_i, metric_i, value_i, id_i = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), 0
    launches = {}
    for r in body:
        d = launches.setdefault(int(r[id_i]), {"name": r[name_i]})
        d[r[metric_i]] = float(r[value_i].replace(",", ""))
    order = [launches[k] for k in sorted(launches)]
    order = [l for l in order if not l["name"].startswith(("void at::", "b200_ffma"))]
    starts = [i for i, l in enumerate(order) if l["name"].startswith("rt_prep_planes_kernel")]
    frame = order[starts[-1]:]
    out = {"flop_per_frame": 0.0, "fadd": 0.0, "fmul": 0.0, "ffma": 0.0, "thread_inst": 0.0, "warp_inst": 0.0,
           "kernel_ns_under_ncu": 0.0, "launches": [l["name"][:48] for l in frame]}
    for l in frame:
        fa = l.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", 0.0)
        fm = l.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", 0.0)
        ff = l.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", 0.0)
        out["fadd"] += fa; out["fmul"] += fm; out["ffma"] += ff
        out["flop_per_frame"] += fa + fm + 2 * ff
        out["thread_inst"] += l.get("smsp__thread_inst_executed.sum", 0.0)
        out["warp_inst"] += l.get("smsp__inst_executed.sum", 0.0)
        out["kernel_ns_under_ncu"] += l.get("gpu__time_duration.sum", 0.0)
    return out


if __name__ == "__main__":
    rnd, res = sys.argv[1], {"_comment": "executed FP32 flop of one whole frame (FADD + FMUL + 2 x FFMA, thread level, predicated-on), "
                                         "ncu counters summed over the frame's launches; see profiles/capture_flops.sh"}
    for p in sys.argv[2:]:
        w = os.path.basename(p)[len("flops_"):-len(".csv")]
        res[w] = frame_flops(p)
        print(w, res[w]["flop_per_frame"], res[w]["kernel_ns_under_ncu"])
    os.makedirs(os.path.join(ROOT, "profiles", rnd), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", rnd, "rt_flops.json"), "w") as f:
        json.dump(res, f, indent=1)

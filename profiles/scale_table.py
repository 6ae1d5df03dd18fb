"""Scaling table from bench.py JSON lines of several N (gpurun_out/b1.log, b2.log, b4.log, b8.log): the top-level
line (rt_cornell_4k) and the nested raster / rt_tess100k lines.  python profiles/scale_table.py gpurun_out/b{1,2,4,8}.log"""
import json
import sys

rows = {}
for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        for name, s in (("rt_cornell_4k", d), ("rast_soup_4k", d.get("raster")), ("rt_tess100k_4k", d.get("rt_tess100k"))):
            if s:
                rows[(name, d["n_gpus"])] = s
print("| workload | N | ms/step | kernel ms | e2e ms | value | speed-up | efficiency | parity |")
print("|---|---|---|---|---|---|---|---|---|")
for n in sorted({k[1] for k in rows}):
    for name in ("rt_cornell_4k", "rast_soup_4k", "rt_tess100k_4k"):
        s = rows.get((name, n))
        if not s:
            continue
        base = rows.get((name, 1))
        sp = s["value"] / base["value"] if base else float("nan")
        par = s.get("parity") or {}
        ptxt = "-" if not par else f"{par.get('assembled_frame_vs_single_gpu')} / {par.get('library_multi_frame_vs_single_gpu')}"
        print(f"| {name} | {n} | {s['ms_per_step']:.4f} | {s['roofline']['kernel_ms']:.4f} | {s['e2e']['ms_per_step']:.3f} | "
              f"{s['value']:.1f} {s['unit']} | {sp:.2f}x | {sp / n:.2f} | {ptxt} |")

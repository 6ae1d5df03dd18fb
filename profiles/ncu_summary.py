"""Reads .ncu-rep files (ncu -i ... --page raw --csv) and prints the per-kernel
figures quoted in DESIGN.md / bench.py; also summarises launch-list CSVs.
Usage: python profiles/ncu_summary.py report.ncu-rep [...]   |   launches.csv [...]"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "XU pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
])


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    last = {}                                   # a capture spans several frames: keep each kernel's last launch
    for r in rows[2:]:
        last[(r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")])] = r
    for r in last.values():
        name = r[hdr.index("Kernel Name")]
        print(f"### {name}  (grid {r[hdr.index('Grid Size')]}, block {r[hdr.index('Block Size')]})")
        for k, label in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                print(f"- {label}: {r[i]} {units[i]}  (`{k}`)")
        print()


def launches(path):
    """Summarises the last frame of bench.py's TIMED region: frames start at rt_prep_planes_kernel /
    rast_geom_kernel<0> (two-pass geometry) or <2> (single pass, pipelined frames); the end-to-end frames that follow (sliced, smaller grids) and the FFMA peak
    microbenchmark are not part of it."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    rows = [r for r in rows if not r[4].startswith(("b200_ffma_peak_kernel", "void at::"))]
    starts = [i for i, r in enumerate(rows) if r[4].startswith(("rt_prep_planes_kernel", "void rast_geom_kernel<0>",
                                                                 "void rast_geom_kernel<2>"))]
    frames = [rows[a:b] for a, b in zip(starts, starts[1:] + [len(rows)])]

    def blocks(fr):   # thread blocks of the frame's final kernel: the sliced end-to-end frames have fewer
        g = [int(v) for v in fr[-1][8].strip("()").split(",")]
        return g[0] * g[1] * g[2]
    full = max(blocks(fr) for fr in frames)
    frame = [fr for fr in frames if blocks(fr) == full][-1]
    total = sum(float(r[-1]) for r in frame)
    print(f"### {path}: last timed frame, {len(frame)} launches, {total / 1e3:.1f} us in kernels")
    print("| kernel | grid | block | us | share |")
    print("|---|---|---|---|---|")
    for r in frame:
        print(f"| `{r[4][:60]}` | {r[8]} | {r[7]} | {float(r[-1]) / 1e3:.1f} | {100 * float(r[-1]) / total:.1f}% |")
    print()


for p in sys.argv[1:]:
    (report if p.endswith(".ncu-rep") else launches)(p)

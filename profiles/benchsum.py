"""Prints the key fields of bench.py JSON lines read from stdin (one per line)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(d.get("impl", "b200"), d["config"]["workload"], "gpus", d["n_gpus"], "value", round(d["value"], 3), d["unit"],
          "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(r.get("kernel_ms", 0) or 0, 4),
          "frac", round(r.get("frac", 0) or 0, 4), "e2e_ms", round(e.get("ms_per_step", 0) or 0, 3),
          "launches", d.get("gpu_launches"), sys.argv[1] if len(sys.argv) > 1 else "")

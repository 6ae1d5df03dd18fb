#!/usr/bin/env bash
# Executed-FP32 counters of one whole RT frame per workload (the roofline numerator bench.py reports):
#   ncu --metrics smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on.sum over every launch of
#   profiles/prof_frame.py <workload> 2 (two frames; the second is summarised by profiles/rt_flops.py).
# Run under gpurun (one GPU); each capture follows a plain run of the same command that exited 0.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for w in "$@"; do
  python profiles/prof_frame.py "$w" 2 > "$OUT/plain_flops_$w.log" 2>&1 &&
  ncu --metrics "$M" --clock-control none --csv --log-file "$OUT/flops_$w.csv" python profiles/prof_frame.py "$w" 2 > "$OUT/ncu_flops_$w.log" 2>&1
  echo "flops capture $w rc=$?"
done

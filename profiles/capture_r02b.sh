#!/usr/bin/env bash
# Round-2 evidence, final code: launch lists (gpu__time_duration per launch of bench.py's timed frames) of the five
# workloads and full captures of the two whose kernels changed after profiles/capture_r02.sh ran (the ordered raster
# path: rast_geom_warp_kernel, rast_short_kernel, fused fill / resolve; the gridded raytracer: pair lists, planned
# frames).  Summarised on the GPU box by profiles/ncu_summary.py (the .ncu-rep files are deleted).
# Run under gpurun, one GPU:   bash profiles/capture_r02b.sh   -> gpurun_out/launches_<w>.csv,
# gpurun_out/ncu_summary_r02b_body.md.  Every ncu pass follows a plain run of the same command that exited 0.
set -uo pipefail
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/ncu_summary_r02b_body.md
for w in rt_cornell_4k rt_tess100k_4k rast_soup_4k rast_cornell_4k rast_cornell_default; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $OUT/plain_$w.log 2>&1 || { echo "plain $w failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_l_$w.log 2>&1
  echo "launch list $w rc=$?"
  case $w in
    rt_tess100k_4k|rast_cornell_4k)
      python profiles/prof_frame.py $w 3 > $OUT/plain_pf_$w.log 2>&1 || { echo "plain prof_frame $w failed"; continue; }
      ncu --set full --clock-control none -c 48 -o $OUT/full_$w -f python profiles/prof_frame.py $w 3 > $OUT/ncu_f_$w.log 2>&1
      echo "full capture $w rc=$?"
      { echo "## $w"; echo; python profiles/ncu_summary.py $OUT/full_$w.ncu-rep $OUT/launches_$w.csv; } >> $OUT/ncu_summary_r02b_body.md
      rm -f $OUT/full_$w.ncu-rep $OUT/ncu_f_$w.log ;;
    *)
      { echo "## $w"; echo; python profiles/ncu_summary.py $OUT/launches_$w.csv; } >> $OUT/ncu_summary_r02b_body.md ;;
  esac
  rm -f $OUT/ncu_l_$w.log
done
du -sh $OUT

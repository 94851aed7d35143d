#!/bin/bash
# Round 2, GPU call O (1 GPU): DRAM traffic of the dominant kernels of C1, C2, C3, C5 (both arithmetic modes) for the
# `traffic` fields of bench.py's config blocks: ncu with the two DRAM byte counters only (one pass per kernel).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for w in c1 c2 c3 c5; do
  case $w in
    c1) K='regex:direct_fast_kernel|direct_strict_split_kernel|direct_strict_kernel'; C=8;;
    c2) K='regex:tp_multistep_kernel'; C=6;;
    c3) K='regex:direct_fast_kernel|direct_strict_kernel'; C=4;;
    c5) K='regex:walk_rec_kernel|tree_collision_kernel'; C=24;;
  esac
  timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" -c $C --csv --log-file gpurun_out/o_traffic_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/o_traffic_$w.log 2>&1
  echo "$w rc=$? lines=$(wc -l < gpurun_out/o_traffic_$w.csv)"
done

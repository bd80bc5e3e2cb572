#!/bin/bash
# Executed warp-instructions and duration of one R = 9 launch of the fused kernel per variant library (ncu, two metrics):
#   tools/gpu_inst.sh <tag> lib1.so lib2.so ...
tag=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --clips 1 --frames 24 --no-e2e --no-cpu-baseline"
for v in "$@"; do
  VOS_LIB_NAME=$v $CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed for $v"; continue; }
  VOS_LIB_NAME=$v ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none -k regex:vos_affinity_idx -s 40 -c 1 --csv --log-file gpurun_out/${tag}_$v.csv $CMD > /dev/null 2>&1
  echo "== $v: $(tail -4 gpurun_out/${tag}_$v.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | tr '\n' ' ')"
done

#!/bin/bash
# Round-2 evidence: launch list + ncu --set full captures of every kernel of the path (one B200).  Each ncu run follows a
# plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
B="python bench.py --workload uniform --clips 1 --frames 24 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-records"
S="python tools/sweep_configs.py"
run() { # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  "$@" > gpurun_out/${name}_plain.log 2>&1 || { echo "plain run failed: $name"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  echo "$name: $(tail -1 gpurun_out/${name}_ncu.log)"
}
$B > gpurun_out/launches_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vos_ -s 60 -c 90 --csv --log-file gpurun_out/r2_launches_vos.csv $B > /dev/null 2>&1
run r2_affinity_idx_f16 vos_affinity_idx 40 2 $B
run r2_affinity_idx_split3 vos_affinity_idx 40 2 $B --precision split3
run r2_merge_writeback vos_merge_writeback 40 2 $B
run r2_append vos_append 40 2 $B
run r2_topk_scan vos_topk_scan 4 2 $S topk1
run r2_topk_threshold vos_topk_threshold 2 1 $S topk1
run r2_topk_finish vos_topk_finish 2 1 $S topk1
run r2_affinity_tc vos_affinity_tc 2 1 $S prob1
ls -la gpurun_out/*.ncu-rep

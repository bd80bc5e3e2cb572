#!/bin/bash
# End-of-round-2 evidence on one B200: GPU suite, A/B of variant libraries, the bench lines (DAVIS-30 and YouTube-VOS-shaped sets),
# the secondary-configuration sweep, the end-to-end breakdown and ncu captures of the kernels changed late in the round.
#   tools/gpu_evidence_r2.sh <tag> [variant libs for the A/B ...]
# Every ncu run follows a plain run of the same command that exited 0; nothing printed under ncu is a bench value.
set -u
tag=$1; shift
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${tag}_gputest.log 2>&1; echo "gpu tests rc=$? $(tail -1 $O/${tag}_gputest.log)"
[ $# -gt 0 ] && bash tools/gpu_ab.sh $tag "$@"
python bench.py > $O/${tag}_bench.json 2> $O/${tag}_bench.err; echo "bench rc=$?"
python bench.py --workload ytvos --no-cpu-baseline > $O/${tag}_bench_ytvos.json 2> $O/${tag}_bench_ytvos.err; echo "bench ytvos rc=$?"
timeout 300 python tools/sweep_configs.py > $O/${tag}_sweep_configs.jsonl 2> $O/${tag}_sweep.err; echo "sweep rc=$?"
timeout 200 python tools/e2e_breakdown.py > $O/${tag}_e2e_breakdown.txt 2>&1; echo "e2e breakdown rc=$?"
B="python bench.py --workload uniform --clips 1 --frames 24 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-records"
S="python tools/sweep_configs.py"
run() { # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  "$@" > $O/${name}_plain.log 2>&1 || { echo "plain run failed: $name"; return; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o $O/$name "$@" > $O/${name}_ncu.log 2>&1
  ncu -i $O/$name.ncu-rep --page details > $O/${name}_details.txt 2>/dev/null
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
  echo "$name: $(tail -1 $O/${name}_ncu.log)"
}
$B > $O/${tag}_launches_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vos_ -s 40 -c 90 --csv --log-file $O/${tag}_launches_vos.csv $B > /dev/null 2>&1
run ${tag}_merge_writeback vos_merge_writeback 15 2 $B
run ${tag}_append vos_append 1 1 $B
run ${tag}_affinity_prob vos_affinity_prob 2 1 $S prob1
run ${tag}_affinity_idx_f16 vos_affinity_idx 15 1 $B
rm -f $O/*.ncu-rep   # details + raw pages are what profiles/ keeps
ls -la $O | tail -40

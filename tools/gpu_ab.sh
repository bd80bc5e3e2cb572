#!/bin/bash
# A/B of variant builds on one box: tools/gpu_ab.sh <tag> lib1.so lib2.so ...   (device-timed propagation stage only)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" >> gpurun_out/${tag}_ab.log
  VOS_LIB_NAME=$v python bench.py --no-e2e --no-cpu-baseline --steps 3 --warmup 2 >> gpurun_out/${tag}_ab.log 2>&1
done
python - <<PY
import json
for line in open('gpurun_out/${tag}_ab.log'):
    if line.startswith('=='): print(line.strip(), end='  ')
    elif line.startswith('{'):
        d = json.loads(line); r = d['roofline']
        print('value %.0f  affinity %.1f us  frac %.3f  clocks %s' % (d['value'], r['avg_launch_us'], r['frac'], d['clocks'].get('sm_mhz')))
PY

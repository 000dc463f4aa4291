#!/bin/bash
# usage: tools/quick_bench.sh "workload[:T] ..."   -- short device-resident bench lines (value, ms) of several workloads
for spec in $1; do
  w=${spec%%:*}; t=""; [ "$spec" != "$w" ] && t="--trajectories ${spec##*:}"
  python bench.py --workload $w $t --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'])"
done

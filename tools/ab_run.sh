#!/bin/bash
# usage: tools/ab_run.sh WORKLOAD T "tagA tagB ..." [bench args]   (run under gpurun; prints value per variant, restores base)
WL=$1; T=$2; TAGS=$3; shift 3
for t in $TAGS; do
  cp ab/libnqcb200_$t.so nqcdynamics.jl_b200/csrc/libnqcb200.so
  python bench.py --workload $WL --trajectories $T --steps 3 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/ab_$t.json').read().strip().splitlines()[-1]); print('$t', '%.4g'%d['value'], '%.2f ms'%d['ms_per_step'])" || tail -3 gpurun_out/ab_$t.err
done
cp ab/libnqcb200_base.so nqcdynamics.jl_b200/csrc/libnqcb200.so

#!/bin/bash
# IESH GPU visit: parity tests, compute-sanitizer on a small case, short benches.
set -u
mkdir -p gpurun_out
make -C oracle -s
echo "== pytest iesh"; timeout 500 python -m pytest tests/test_parity_gpu.py -m gpu -q -k iesh --timeout 300 2>&1 | tail -15
if [ "${SANITIZE:-1}" = 1 ]; then
echo "== memcheck"; timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_case.py iesh_anderson_holstein_m30 4 6 1 2>&1 | tail -4; timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_case.py iesh_anderson_holstein_m100 2 3 1 2>&1 | tail -6; timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_case.py iesh_anderson_holstein_m200 1 2 1 2>&1 | tail -6
echo "== racecheck"; timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 5 python tools/profile_case.py iesh_anderson_holstein_m30 2 3 1 2>&1 | tail -8
fi
for wl in ${WORKLOADS:-iesh_anderson_holstein_m100 iesh_anderson_holstein_m200}; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps ${STEPS:-2} --warmup ${WARMUP:-1} --trajectories ${TRAJ:-1480} ${BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/bench_$wl.json
done

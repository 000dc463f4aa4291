#!/bin/bash
# One GPU visit: parity tests (bounded), smoke, then short benches.  Everything is wrapped in `timeout`.
set -u
mkdir -p gpurun_out
make -C oracle -s
echo "== pytest -m gpu"; timeout ${PYTEST_TIMEOUT:-420} python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -${PYTEST_TAIL:-25}
echo "== smoke"; timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5
for wl in ${WORKLOADS:-tully1_fssh spinboson_debye100_fssh}; do
  echo "== bench $wl"; timeout 400 python bench.py --workload $wl --steps ${STEPS:-3} --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -3 | tee gpurun_out/bench_$wl.json
done

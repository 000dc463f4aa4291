#!/bin/bash
# A/B of the SpinBoson kernels on the bench default (run under gpurun): epoch kernels (E = 16, 8) vs step-by-step
mkdir -p gpurun_out
for k in 16 8 0; do
  NQCB200_SPINBOSON_EPOCH=$k python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_sb_e$k.json 2> gpurun_out/ab_sb_e$k.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_sb_e$k.json"))
print("EPOCH=$k value %.4g e2e %.4g ms/step %.2f kernel_ms %.2f checksum %s hops %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["kernel_ms_total"]/d["steps"], d["observable_checksum"], d["counters"]["hops"]))
PY
done

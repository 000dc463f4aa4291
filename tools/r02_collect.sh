#!/bin/bash
# After tools/r02_measure.sh (run under gpurun): copy the evidence from gpurun_out/ into profiles/r02/ and rebuild
# profiles/r02/executed_flops.json (the executed FP64 flops per trajectory-step that bench.py's roofline.frac uses).
set -e
cd "$(dirname "$0")/.."
P=profiles/r02
mkdir -p $P
for t in sb_fssh sb_ehr tully1 rpmd rpsh nrpmd langevin; do cp gpurun_out/jobmix_$t.csv $P/; done
python tools/executed_flops.py $P/executed_flops.json \
  "spinboson_debye100_fssh=$P/jobmix_sb_fssh.csv:262144:200:sb_bath|sb_elec|sb_prep" \
  "spinboson_debye100_ehrenfest=$P/jobmix_sb_ehr.csv:262144:200:sb_bath|sb_elec|sb_prep" \
  "tully1_fssh=$P/jobmix_tully1.csv:1048576:600:density_step_kernel" \
  "rpmd_harmonic32=$P/jobmix_rpmd.csv:262144:1000:classical_tpt" \
  "rpsh_morse3_16=$P/jobmix_rpsh.csv:113664:600:ring_tpt_step" \
  "nrpmd_morse3_16=$P/jobmix_nrpmd.csv:100000:600:nrpmd_step" \
  "langevin_harmonic32=$P/jobmix_langevin.csv:262144:1000:langevin"
for f in bench_default bench_tully1_fssh bench_spinboson_debye100_ehrenfest bench_rpmd_harmonic32 bench_rpsh_morse3_16 bench_nrpmd_morse3_16 \
         bench_langevin_harmonic32 bench_iesh_anderson_holstein_m100 bench_iesh_anderson_holstein_m200 bench_stream_spinboson_debye100_fssh bench_reference; do
  [ -s gpurun_out/$f.json ] && cp gpurun_out/$f.json $P/
done
cp gpurun_out/launches_r02_bench_default.csv $P/ 2>/dev/null || true
ls $P | wc -l

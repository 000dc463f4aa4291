#!/bin/bash
# Round-2 evidence pass on one B200 (run under gpurun): per-job instruction mixes, bench lines of every workload, the
# launch list of the default bench command.  Outputs under gpurun_out/ (copied to profiles/r02/ afterwards).
mkdir -p gpurun_out
tools/instmix_job.sh spinboson_debye100_fssh 262144 200 sb_fssh
tools/instmix_job.sh spinboson_debye100_ehrenfest 262144 200 sb_ehr
tools/instmix_job.sh tully1_fssh 1048576 600 tully1
tools/instmix_job.sh rpmd_harmonic32 262144 1000 rpmd
tools/instmix_job.sh rpsh_morse3_16 113664 600 rpsh
tools/instmix_job.sh nrpmd_morse3_16 100000 600 nrpmd
tools/instmix_job.sh langevin_harmonic32 262144 1000 langevin
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for w in tully1_fssh spinboson_debye100_ehrenfest rpmd_harmonic32 rpsh_morse3_16 nrpmd_morse3_16 langevin_harmonic32 iesh_anderson_holstein_m100; do
  python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python bench.py --workload iesh_anderson_holstein_m200 --trajectories 2960 --steps 2 --warmup 2 > gpurun_out/bench_iesh_anderson_holstein_m200.json 2> gpurun_out/bench_iesh_m200.err
python bench.py --stream --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_stream_spinboson_debye100_fssh.json 2> /dev/null
/usr/local/cuda/bin/ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r02_bench_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/launches_r02_bench_default.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> /dev/null
ls -la gpurun_out | tail -30

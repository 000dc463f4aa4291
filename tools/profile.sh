#!/bin/bash
# ncu evidence for one workload: launch list, full-set capture of the step kernel, FP64 instruction mix.
# usage: tools/profile.sh WORKLOAD KERNEL_REGEX T NSTEPS TAG
set -u
WL=$1; KRE=$2; T=$3; NS=$4; TAG=$5
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 300 $NCU --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/profile_case.py $WL $T $NS 3 > gpurun_out/launches_$TAG.log 2>&1
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/profile_case.py $WL $T $NS 2 > gpurun_out/prof_$TAG.log 2>&1
timeout 300 $NCU --clock-control none -k regex:$KRE -s 1 -c 1 --csv --log-file gpurun_out/instmix_$TAG.csv --metrics \
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    python tools/profile_case.py $WL $T $NS 2 > gpurun_out/instmix_$TAG.log 2>&1
# the .ncu-rep of these long kernels is ~60 MB (gpurun_out is capped at 64 MiB): export the pages we read
$NCU -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
$NCU -i gpurun_out/prof_$TAG.ncu-rep --page details --csv > gpurun_out/prof_${TAG}_details.csv 2>/dev/null
$NCU -i gpurun_out/prof_$TAG.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/prof_${TAG}_source.csv.gz
$NCU -i gpurun_out/prof_$TAG.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip -9 > gpurun_out/prof_${TAG}_cudasass.csv.gz
rm -f gpurun_out/prof_$TAG.ncu-rep
tail -2 gpurun_out/launches_$TAG.log
ls -la gpurun_out | head -20

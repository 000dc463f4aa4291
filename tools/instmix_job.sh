#!/bin/bash
# FP64 instruction mix + DRAM bytes + duration of EVERY kernel of one job of a workload (one nqcb200_run of NSTEPS steps).
# usage: tools/instmix_job.sh WORKLOAD T NSTEPS TAG      -> gpurun_out/jobmix_TAG.csv  (read by tools/executed_flops.py)
set -u
WL=$1; T=$2; NS=$3; TAG=$4
mkdir -p gpurun_out
timeout 900 /usr/local/cuda/bin/ncu --clock-control none -c 400 --csv --log-file gpurun_out/jobmix_$TAG.csv --metrics \
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,launch__registers_per_thread \
    python tools/profile_case.py $WL $T $NS 1 > gpurun_out/jobmix_$TAG.log 2>&1
tail -1 gpurun_out/jobmix_$TAG.log

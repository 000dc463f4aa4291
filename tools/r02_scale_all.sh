#!/bin/bash
# usage: tools/r02_scale_all.sh N   (under gpurun --gpus N): the default workload (weak) and BASELINE config 5 as written
# (10^5 trajectories in total, strong) on N GPUs
N=$1
tools/r02_scale.sh $N default --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs
tools/r02_scale.sh $N rpsh_strong --workload rpsh_morse3_16 --scaling strong --steps 3 --warmup 3 --no-cpu-baseline
tools/r02_scale.sh $N tully1 --workload tully1_fssh --steps 3 --warmup 3 --no-cpu-baseline

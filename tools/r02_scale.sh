#!/bin/bash
# usage: tools/r02_scale.sh N TAG [bench args...]  -- one torchrun launch of bench.py on N GPUs (run under gpurun --gpus N)
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
fi
tail -c 400 gpurun_out/scale_${TAG}_n$N.json; echo; tail -3 gpurun_out/scale_${TAG}_n$N.err

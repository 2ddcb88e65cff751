#!/bin/bash
# bench.py at N = 1, 2, 4, ... GPUs of one box (the driver's launch line); usage: tools/scale_run.sh TAG N1 N2 ...
TAG=$1; shift
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/scale_${TAG}_n1.json 2> gpurun_out/scale_${TAG}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
  fi
  tail -1 gpurun_out/scale_${TAG}_n$N.json | cut -c1-200
done

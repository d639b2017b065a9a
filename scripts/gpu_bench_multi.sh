#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out/benchN
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/benchN/n1.json 2> gpurun_out/benchN/n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/benchN/n$N.json 2> gpurun_out/benchN/n$N.err
fi
tail -n 1 gpurun_out/benchN/n$N.json | cut -c1-1500; tail -n 5 gpurun_out/benchN/n$N.err | cut -c1-300

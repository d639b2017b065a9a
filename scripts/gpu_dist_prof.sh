#!/bin/bash
W=${1:-2}
mkdir -p gpurun_out/dist2
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; w=$2; shift 2; timeout 600 $T --nproc-per-node $w scripts/dist_check.py "$@" > gpurun_out/dist2/$tag.log 2>&1; grep "^{\|pcg-dist" gpurun_out/dist2/$tag.log | cut -c1-330 | tail -n 5; grep -i "error\|assert\|timeout" gpurun_out/dist2/$tag.log | head -n 3; }
PPF_B200_PCG_TUNE=1 run lt_w$W $W large-twist 0 2

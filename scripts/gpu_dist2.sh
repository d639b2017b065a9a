#!/bin/bash
set -x
mkdir -p gpurun_out/dist2
nvidia-smi -L | head -n 8
W=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { # tag world scene n steps [env...]
  tag=$1; w=$2; shift 2
  timeout 600 $T --nproc-per-node $w scripts/dist_check.py "$@" > gpurun_out/dist2/$tag.log 2>&1; grep "^{" gpurun_out/dist2/$tag.log | tail -n 1; grep -i "error\|assert" gpurun_out/dist2/$tag.log | head -n 3
}
run ft_w1 1 fast-twist 48 40
run ft_w$W $W fast-twist 48 40
python - <<PY
import numpy as np
a=np.load('/tmp/dist_check_curr_w1.npy'); b=np.load('/tmp/dist_check_curr_w$W.npy')
d=np.abs(a-b).max(axis=1); print("fast-twist 40 steps, w1 vs w$W positions: max %.3e median %.3e" % (d.max(), np.median(d)))
PY
run lt_w1 1 large-twist 0 3
run lt_w$W $W large-twist 0 3
PPF_B200_DIST_NCCL=1 run lt_w${W}_nccl $W large-twist 0 3

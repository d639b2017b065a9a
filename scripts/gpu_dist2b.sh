#!/bin/bash
mkdir -p gpurun_out/dist2
W=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; w=$2; shift 2; timeout 600 $T --nproc-per-node $w scripts/dist_check.py "$@" > gpurun_out/dist2/$tag.log 2>&1; grep "^{" gpurun_out/dist2/$tag.log | tail -n 1; grep -i "error\|assert" gpurun_out/dist2/$tag.log | head -n 3; }
export PPF_B200_DIST_DEBUG=1
run ft_w1 1 fast-twist 48 40
run ft_w$W $W fast-twist 48 40
grep "dist 1/" gpurun_out/dist2/ft_w$W.log | awk 'NR%20==1' | head -n 5
python - <<PY
import numpy as np
a=np.load('/tmp/dist_check_curr_w1.npy'); b=np.load('/tmp/dist_check_curr_w$W.npy')
d=np.abs(a-b).max(axis=1); print("fast-twist 40 steps, w1 vs w$W positions: max %.3e median %.3e" % (d.max(), np.median(d)))
PY
run lt_w1 1 large-twist 0 3
run lt_w$W $W large-twist 0 3
grep "dist 1/" gpurun_out/dist2/lt_w$W.log | head -n 3

#!/bin/bash
# Round 2, closing call (8 GPUs): bench.py as the driver launches it at N=8 and N=4 (large-twist), after the last PCG / BVH changes
O=gpurun_out/r2n8
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
for n in 8 4; do
  t0=$(date +%s)
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 6 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "bench N=$n rc=$? in $(( $(date +%s) - t0 )) s"
  python - "$O/bench_n$n.json" "$n" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("N=%s: %.1f ms/step, %.1f us/PCG iter, iters %d, phases %s, value %.4f, e2e %.4f, clocks %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["value"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
done
tail -n 4 $O/bench_n8.err | cut -c1-300

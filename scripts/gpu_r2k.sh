#!/bin/bash
# Round 2 (N GPUs): large-twist sharded over the GPUs of one box as the driver launches it; $1 = N
N=${1:-8}
O=gpurun_out/r2k
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"
python - "$O/bench_n$N.json" "N=$N" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("%s: %.1f ms/step, %.1f us/PCG iter, iters/solve %.0f, solves %d, phases %s, value %.4f, roofline %.3f" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters_per_solve"], d["work"]["newton_solves"],
          {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["value"], d.get("roofline", {}).get("frac", 0)))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
tail -n 4 $O/bench_n$N.err | cut -c1-300

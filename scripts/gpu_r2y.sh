#!/bin/bash
# Round 2 (1 GPU): packed BVH traversal records -- the whole GPU suite, then the line-search / contact phases of the five configurations
O=gpurun_out/r2y
mkdir -p $O
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rfEs > $O/pytest_full.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_full.log | cut -c1-300 | tail -n 12
line() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-14s %8.2f ms/step  linsolve %6.1f us/PCG iter  iters %d  phases %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for scene in hang five-twist twist large-woven large-twist; do
  steps=20; [ $scene = hang ] && steps=60; [ $scene = large-twist ] && steps=6; [ $scene = large-woven ] && steps=6
  timeout 900 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_$scene.json 2> $O/bench_$scene.err
  line $O/bench_$scene.json "$scene"
done

#!/bin/bash
# Round 2 (1 GPU): does the evict_first hint on the matrix stream hurt when the whole system fits the 126 MB L2 (twist 47 MB, five-twist 24 MB, hang 3.5 MB)?
O=gpurun_out/r2l2
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
line() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-26s %8.2f ms/step  linsolve %6.1f us/PCG iter  iters %d  spmv %.2f us" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], d["roofline"]["us_per_launch"]))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for scene in twist five-twist hang; do
  steps=20; [ $scene = hang ] && steps=60
  for hint in 1 0 1 0; do
    PPF_B200_L2_STREAM=$hint timeout 300 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_${scene}_hint$hint.json 2> $O/bench_${scene}_hint$hint.err
    line $O/bench_${scene}_hint$hint.json "$scene evict_first=$hint"
  done
done

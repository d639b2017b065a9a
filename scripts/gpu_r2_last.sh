#!/bin/bash
# Round 2, last call (1 GPU): evict_first only for systems larger than the L2 -- smoke, parity subset, twist / large-twist lines
O=gpurun_out/r2last
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rods.py tests/test_smoke_entry.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest.log
line() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-14s %8.2f ms/step  linsolve %6.1f us/PCG iter  iters %d  spmv %.2f us  frac %.3f" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for scene in twist large-twist; do
  steps=20; [ $scene = large-twist ] && steps=6
  timeout 600 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_$scene.json 2> $O/bench_$scene.err
  line $O/bench_$scene.json $scene
done

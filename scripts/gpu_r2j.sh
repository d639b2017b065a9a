#!/bin/bash
# Round 2, tenth GPU call (1 GPU): Schwarz mode after the warp-reduced restriction fix
O=gpurun_out/r2j
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_schwarz.py -q -m gpu -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log | cut -c1-300
for scene in large-twist twist; do
  steps=4; [ $scene = twist ] && steps=20
  timeout 900 python bench.py --scene $scene --steps $steps --warmup 3 --precond 1 > $O/bench_${scene}_schwarz.json 2> $O/bench_${scene}_schwarz.err
  python - "$O/bench_${scene}_schwarz.json" "$scene schwarz" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s %9.2f ms/step  linsolve %.1f us/PCG iter  iters/solve %7.1f  phases %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters_per_solve"], {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done

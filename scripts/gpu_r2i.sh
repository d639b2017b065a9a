#!/bin/bash
# Round 2, ninth GPU call (1 GPU): regression of the multi-CTA scan and the warp-reduced restriction, Schwarz-mode timing again
O=gpurun_out/r2i
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_prims.py tests/test_gpu_schwarz.py tests/test_gpu_parity.py tests/test_smoke_entry.py -q -m gpu -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log | cut -c1-300
for pre in 1 0; do
  timeout 900 python bench.py --steps 4 --warmup 3 --precond $pre > $O/bench_lt_precond$pre.json 2> $O/bench_lt_precond$pre.err
  python - "$O/bench_lt_precond$pre.json" "large-twist precond=$pre" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s %9.2f ms/step  linsolve %.1f us/PCG iter  iters/solve %7.1f  phases %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters_per_solve"], {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done

#!/bin/bash
# Round 2 (1 GPU): control-block head read in one request by the iteration kernels -- parity subset and the small configurations
O=gpurun_out/r2t
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rods.py tests/test_smoke_entry.py tests/test_gpu_schwarz.py -q -m gpu -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest.log
line() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-26s %8.2f ms/step  linsolve %6.1f us/PCG iter  iters %d  spmv %.2f us  phases %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], d["roofline"]["us_per_launch"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for scene in hang five-twist twist; do
  steps=20; [ $scene = hang ] && steps=60
  for rep in 1 2; do
    timeout 600 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_${scene}_$rep.json 2> $O/bench_${scene}_$rep.err
    line $O/bench_${scene}_$rep.json "$scene #$rep"
  done
done

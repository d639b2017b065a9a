#!/bin/bash
# Round 2, fifth GPU call: the persistent cooperative PCG loop (k_pcg_loop) -- parity first, then A/B timing against the 3-kernel graph
O=gpurun_out/r2e
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_smoke_entry.py tests/test_gpu_rods.py -q -m gpu -x -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E   " $O/pytest_gpu.log | cut -c1-400 | tail -n 20
for v in loop graph; do
  for scene in large-twist twist; do
    if [ $v = graph ]; then export PPF_B200_PCG_GRAPH=1; else unset PPF_B200_PCG_GRAPH; fi
    steps=4; [ $scene = twist ] && steps=20
    timeout 600 python bench.py --scene $scene --steps $steps --warmup 3 > $O/bench_${v}_$scene.json 2> $O/bench_${v}_$scene.err; echo "bench $v $scene rc=$?"
    python - "$O/bench_${v}_$scene.json" "$v $scene" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%s: %.1f ms/step, %.1f us/PCG iter (linsolve), iters/solve %.0f, launches %s, value %.4f, clocks %s" % (
        sys.argv[2], d["ms_per_step"], 1e3 * d["phase_ms_per_step"]["linsolve"] / max(1, d["work"]["pcg_iters"]) * d["steps"],
        d["work"]["pcg_iters_per_solve"], d["gpu_launches"], d["value"], d["clocks"]))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
  done
done
tail -n 5 $O/bench_loop_large-twist.err

#!/bin/bash
# Round 2 (1 GPU): status poll one batch ahead -- parity subset, A/B on the same box; clean A/B of the chained launch at large-twist
O=gpurun_out/r2s
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rods.py tests/test_smoke_entry.py tests/test_gpu_schwarz.py -q -m gpu -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest.log
line() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-26s %8.2f ms/step  linsolve %6.1f us/PCG iter  iters %d  phases %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for scene in hang five-twist twist; do
  steps=20; [ $scene = hang ] && steps=60
  for pa in 1 0 1 0; do
    PPF_B200_POLL_AHEAD=$pa timeout 600 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_${scene}_ahead$pa.json 2> $O/bench_${scene}_ahead$pa.err
    line $O/bench_${scene}_ahead$pa.json "$scene ahead=$pa"
  done
done
for pdl in 0 1 0 1; do
  PPF_B200_PDL=$pdl timeout 900 python bench.py --steps 6 --warmup 3 --no-extra > $O/bench_large-twist_pdl$pdl.json 2> $O/bench_large-twist_pdl$pdl.err
  line $O/bench_large-twist_pdl$pdl.json "large-twist pdl=$pdl"
done

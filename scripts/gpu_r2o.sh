#!/bin/bash
# Round 2 (1 GPU): the cooperative loop kernel with SMALL grids on the small configurations (a barrier over few CTAs is cheap)
O=gpurun_out/r2o
mkdir -p $O
run() { # label scene steps env...
  label=$1; scene=$2; steps=$3; shift 3
  env "$@" timeout 600 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_${scene}_$label.json 2> $O/bench_${scene}_$label.err
  python - "$O/bench_${scene}_$label.json" "$scene $label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s %8.2f ms/step  linsolve %6.1f us/PCG iter  iters %d  phases %s" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items()}))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run graph hang 60 X=1
for g in 16 37 74 148; do run loop$g hang 60 PPF_B200_PCG_LOOP=1 PPF_B200_LOOP_GRID=$g; done
run graph five-twist 20 X=1
for g in 74 148 296; do run loop$g five-twist 20 PPF_B200_PCG_LOOP=1 PPF_B200_LOOP_GRID=$g; done
run graph twist 20 X=1
for g in 148 296; do run loop$g twist 20 PPF_B200_PCG_LOOP=1 PPF_B200_LOOP_GRID=$g; done

#!/bin/bash
# Round 2 (1 GPU): SpMV with cp.async-filled stages -- kernel timings, parity subset under the variant, bench
O=gpurun_out/r2m
mkdir -p $O
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
PPF_B200_PCG_TUNE=1 timeout 600 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-twist --steps 1 --out $O/lt --quiet > $O/lt_tune.log 2>&1; grep "pcg-tune\] spmv" $O/lt_tune.log
rm -rf $O/lt
PPF_B200_SPMV=cpasync2 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "fast_twist or woven_n32 or hang" > $O/pytest_cpasync2.log 2>&1; echo "pytest cpasync2 rc=$?"; tail -n 2 $O/pytest_cpasync2.log
for v in cpasync2 cpasync3; do
  PPF_B200_SPMV=$v timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_spmv_$v.json 2> $O/bench_spmv_$v.err
  python - "$O/bench_spmv_$v.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("SpMV %s: %.1f ms/step, %.1f us/PCG iter (linsolve), spmv %.1f us frac %.3f, iters/solve %.0f, clocks %s" % (
        sys.argv[2], d["ms_per_step"], 1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]),
        d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["work"]["pcg_iters_per_solve"], d["clocks"]))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
done

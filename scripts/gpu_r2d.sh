#!/bin/bash
# Round 2, fourth GPU call: pinned FMA contraction against the reference's device headers, the whole gpu suite (own radix sort /
# unique instead of CUB, pinned predicates, carried pattern before Dirichlet), SpMV variants on large-twist.
O=gpurun_out/r2d
mkdir -p $O
timeout 300 oracle/_ref/ref_device_check 400000 > $O/ref_device_check.log 2>&1; echo "ref_device_check rc=$?"; cat $O/ref_device_check.log
timeout 2400 python -m pytest tests -q -m gpu -rA -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed|^E   |reference only|this backend only|s call" $O/pytest_gpu.log | cut -c1-600 | tail -n 40
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
PPF_B200_PCG_TUNE=1 timeout 600 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-twist --steps 1 --out $O/lt --quiet > $O/lt_tune.log 2>&1; grep "pcg-tune" $O/lt_tune.log
rm -rf $O/lt
for v in staged direct1 direct2; do
  PPF_B200_SPMV=$v timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_spmv_$v.json 2> $O/bench_spmv_$v.err
  python - "$O/bench_spmv_$v.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("SpMV %s: %.1f ms/step, %.1f us/PCG iter (linsolve), spmv %.1f us frac %.3f, iters/solve %.0f, clocks %s" % (
        sys.argv[2], d["ms_per_step"], 1e3 * d["phase_ms_per_step"]["linsolve"] / (d["pcg_iters_per_newton_step"] * d["newton_steps_per_step"]),
        d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["work"]["pcg_iters_per_solve"], d["clocks"]))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
done

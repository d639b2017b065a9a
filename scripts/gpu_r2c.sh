#!/bin/bash
# Round 2, third GPU call: re-run of the parity tests that failed in r2b (with pair diagnostics), a contact-rich twist window on both
# arms (bench.py --state), and the PCG kernel timings on large-twist as the baseline for this round's PCG work.
O=gpurun_out/r2c
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -m gpu -rA -p no:cacheprovider -k "twist or woven_n32 or headless_ref or log_n16" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E   |reference only|this backend only" $O/pytest_gpu.log | cut -c1-700 | tail -n 40
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
timeout 900 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene twist --steps 300 --save-every 300 --out $O/twist_traj --quiet > $O/twist_traj.log 2>&1; tail -n 2 $O/twist_traj.log
if [ -f $O/twist_traj/state_00300.bin ]; then
  cp $O/twist_traj/state_00300.bin $O/twist_state_00300.bin
  timeout 600 python bench.py --scene twist --state $O/twist_state_00300.bin --steps 20 --warmup 3 > $O/twist_rich_ours.json 2> $O/twist_rich_ours.err; tail -c 2500 $O/twist_rich_ours.json
  timeout 900 python bench.py --impl reference --scene twist --state $O/twist_state_00300.bin --steps 20 --warmup 3 > $O/twist_rich_ref.json 2> $O/twist_rich_ref.err; tail -c 1500 $O/twist_rich_ref.json
fi
rm -rf $O/twist_traj
PPF_B200_PCG_TUNE=1 timeout 600 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-twist --steps 2 --out $O/lt --quiet > $O/lt_tune.log 2>&1; grep "pcg-tune" $O/lt_tune.log
rm -rf $O/lt
# A/B of the L2 policies on the headline scene (short runs: 3 warm-up + 3 timed steps each)
for cfg in "0 0" "0 1" "y 1" "yp 1"; do
  set -- $cfg
  PPF_B200_L2_PERSIST=$1 PPF_B200_L2_STREAM=$2 timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_l2_$1_$2.json 2> $O/bench_l2_$1_$2.err
  python - "$O/bench_l2_$1_$2.json" "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("L2 persist=%s stream_evict_first=%s: %.1f ms/step, %.1f us/PCG iter (linsolve), spmv %.1f us frac %.3f, clocks %s" % (
        sys.argv[2], sys.argv[3], d["ms_per_step"], 1e3 * d["phase_ms_per_step"]["linsolve"] / (d["pcg_iters_per_newton_step"] * d["newton_steps_per_step"]),
        d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["clocks"]))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
done

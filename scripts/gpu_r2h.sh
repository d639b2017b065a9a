#!/bin/bash
# Round 2, eighth GPU call: ParamSet.precond = Schwarz (multilevel additive preconditioner) -- tests, then both arms with their own
# Schwarz implementations; a contact-rich window of the headline scene (state after 120 steps) on both arms.
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_schwarz.py tests/test_gpu_parity.py -q -m gpu -rA -p no:cacheprovider -k "schwarz or fast_twist_n16" > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E   " $O/pytest.log | cut -c1-500 | tail -n 20
line() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-34s %9.2f ms/step  %8.1f us/PCG iter  iters/solve %7.1f  solves %3d  contacts(last) %s  value %.4f" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["ms_per_step"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters_per_solve"], d["work"]["newton_solves"], d["work"]["num_contact_last"], d["value"]))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for scene in twist large-twist; do
  steps=20; [ $scene = large-twist ] && steps=4
  timeout 900 python bench.py --scene $scene --steps $steps --warmup 3 --precond 1 > $O/bench_ours_schwarz_$scene.json 2> $O/bench_ours_schwarz_$scene.err; line $O/bench_ours_schwarz_$scene.json "ours schwarz $scene"
  timeout 900 python bench.py --scene $scene --steps $steps --warmup 3 > $O/bench_ours_bj_$scene.json 2> $O/bench_ours_bj_$scene.err; line $O/bench_ours_bj_$scene.json "ours block-jacobi $scene"
  timeout 1500 python bench.py --impl reference --scene $scene --steps $steps --warmup 3 --precond 1 > $O/bench_ref_schwarz_$scene.json 2> $O/bench_ref_schwarz_$scene.err; line $O/bench_ref_schwarz_$scene.json "reference schwarz $scene"
done
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
timeout 900 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-twist --steps 120 --save-every 120 --out $O/lt_traj --quiet > $O/lt_traj.log 2>&1; tail -n 1 $O/lt_traj.log
if [ -f $O/lt_traj/state_00120.bin ]; then
  mv $O/lt_traj/state_00120.bin /tmp/lt_state_00120.bin
  timeout 900 python bench.py --state /tmp/lt_state_00120.bin --steps 4 --warmup 3 > $O/bench_ours_lt_rich.json 2> $O/bench_ours_lt_rich.err; line $O/bench_ours_lt_rich.json "ours large-twist t=2s"
  timeout 1500 python bench.py --impl reference --state /tmp/lt_state_00120.bin --steps 4 --warmup 3 > $O/bench_ref_lt_rich.json 2> $O/bench_ref_lt_rich.err; line $O/bench_ref_lt_rich.json "reference large-twist t=2s"
fi
rm -rf $O/lt_traj

#!/bin/bash
# First GPU call of the next round (prepared at the end of round 1, when the GPU budget was spent):
#   1. the whole gpu suite, incl. the tests that have never run on a GPU (tests/test_zz_gpu_headless.py, 21 tests);
#   2. reference golden fixtures for the features added after the budget ran out (kinematic colliders, headless, ARAP / StVK /
#      SNHk shells, damping, stitches, plasticity, fix_xz, static collision mesh, collision windows, quad / log barriers): copy gpurun_out/golden/* into tests/golden/ and add the names to tests/conftest.py::golden_cases;
#   3. a contact-rich timing window: twist run to t = 5 s on this backend, state saved, then bench.py --state on both arms;
#   4. ncu launch list of one large-woven step (rod / collider kernels are not in any profile yet).
set -x
O=gpurun_out/r2a
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 30 $O/pytest_gpu.log | cut -c1-300
timeout 600 bash tests/golden/make_golden.sh next-round > $O/golden.log 2>&1; echo "golden rc=$?"; tail -n 8 $O/golden.log
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
timeout 600 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene twist --steps 300 --save-every 300 --out $O/twist_traj --quiet > $O/twist_traj.log 2>&1; tail -n 2 $O/twist_traj.log
if [ -f $O/twist_traj/state_00300.bin ]; then
  timeout 600 python bench.py --scene twist --state $O/twist_traj/state_00300.bin --steps 20 --warmup 3 > $O/twist_rich_ours.json 2> $O/twist_rich_ours.err; tail -c 1500 $O/twist_rich_ours.json
  timeout 900 python bench.py --impl reference --scene twist --state $O/twist_traj/state_00300.bin --steps 20 --warmup 3 > $O/twist_rich_ref.json 2> $O/twist_rich_ref.err; tail -c 600 $O/twist_rich_ref.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file $O/launches_large_woven.csv python bench.py --scene large-woven --steps 1 --warmup 3 > $O/ncu_lw.log 2>&1
python scripts/summarize_launches.py $O/launches_large_woven.csv > $O/launches_large_woven_summary.txt 2>/dev/null; head -n 25 $O/launches_large_woven_summary.txt
rm -rf $O/twist_traj/data $O/*/final_*.bin

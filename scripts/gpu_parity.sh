#!/bin/bash
# GPU parity run: reference vs B200 backend, one step from identical reference states.
set -x
H=./ppf-contact-solver_b200/bin/ppf_host
REF=oracle/_ref/libsimbackend_cuda.so
OUR=ppf-contact-solver_b200/lib/libsimbackend_cuda.so
O=gpurun_out/parity
mkdir -p $O
export PPF_HOST_QUIET=1 PPF_B200_QUIET=1
# 0. sanity + memcheck on a tiny scene
timeout 300 $H --lib $OUR --scene fast-twist --n 16 --steps 3 --out $O/our_tiny > $O/our_tiny.log 2>&1; echo "exit=$?" >> $O/our_tiny.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $H --lib $OUR --scene fast-twist --n 16 --steps 2 --out $O/our_tiny_mc > $O/our_tiny_mc.log 2>&1; echo "exit=$?" >> $O/our_tiny_mc.log
# 1. reference trajectory with saved states
timeout 300 $H --lib $REF --scene fast-twist --steps 120 --save-every 20 --out $O/ref_ft > $O/ref_ft.log 2>&1; echo "exit=$?" >> $O/ref_ft.log
# 2. one step from each saved reference state, both backends
for st in 00020 00040 00060 00080 00100 00120; do
  timeout 120 $H --lib $REF --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/ref_1_$st > $O/ref_1_$st.log 2>&1
  timeout 120 $H --lib $OUR --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --pairs-every 1 --out $O/our_1_$st > $O/our_1_$st.log 2>&1; echo "exit=$?" >> $O/our_1_$st.log
  python scripts/compare_runs.py $O/ref_1_$st $O/our_1_$st > $O/cmp_$st.json 2>&1
done
# 3. full trajectory with our backend
timeout 600 $H --lib $OUR --scene fast-twist --steps 120 --out $O/our_ft > $O/our_ft.log 2>&1; echo "exit=$?" >> $O/our_ft.log
python scripts/compare_runs.py $O/ref_ft $O/our_ft 1e-2 > $O/cmp_traj.json 2>&1
# 4. twist (200k verts): 30 steps each, then 1-step from the reference's state
timeout 600 $H --lib $REF --scene twist --steps 30 --out $O/ref_twist > $O/ref_twist.log 2>&1
timeout 600 $H --lib $OUR --scene twist --steps 30 --out $O/our_twist > $O/our_twist.log 2>&1; echo "exit=$?" >> $O/our_twist.log
timeout 300 $H --lib $REF --scene twist --steps 1 --load $O/ref_twist/state_final.bin --out $O/ref_twist_1 > $O/ref_twist_1.log 2>&1
timeout 300 $H --lib $OUR --scene twist --steps 1 --load $O/ref_twist/state_final.bin --out $O/our_twist_1 > $O/our_twist_1.log 2>&1
python scripts/compare_runs.py $O/ref_twist_1 $O/our_twist_1 > $O/cmp_twist_1.json 2>&1
# 5. large-twist timing, 4 steps
timeout 900 $H --lib $OUR --scene large-twist --steps 4 --out $O/our_large > $O/our_large.log 2>&1; echo "exit=$?" >> $O/our_large.log
rm -f $O/*/final_*.bin $O/*twist*/state_*.bin $O/our_large/*.bin $O/*twist*/dyn_*.bin
cat $O/cmp_*.json
tail -n 5 $O/our_tiny.log $O/our_tiny_mc.log $O/our_ft.log $O/our_twist.log $O/our_large.log
cat $O/our_ft/summary.json $O/our_twist/summary.json $O/ref_twist/summary.json $O/our_large/summary.json
du -sh gpurun_out

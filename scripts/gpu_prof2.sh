#!/bin/bash
set -x
H=./ppf-contact-solver_b200/bin/ppf_host
OUR=ppf-contact-solver_b200/lib/libsimbackend_cuda.so
REF=oracle/_ref/libsimbackend_cuda.so
O=gpurun_out/prof2
mkdir -p $O
export PPF_HOST_QUIET=1 PPF_B200_QUIET=1
# correctness of the new PCG: one step from a reference state
timeout 300 $H --lib $REF --scene fast-twist --steps 60 --save-every 30 --out $O/ref_ft > $O/ref_ft.log 2>&1
for st in 00030 00060; do
  timeout 120 $H --lib $REF --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/ref_1_$st > $O/ref_1_$st.log 2>&1
  timeout 120 $H --lib $OUR --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/our_1_$st > $O/our_1_$st.log 2>&1
  python scripts/compare_runs.py $O/ref_1_$st $O/our_1_$st
done
timeout 300 $H --lib $OUR --scene twist --steps 30 --out $O/our_twist > $O/our_twist.log 2>&1; tail -2 $O/our_twist.log; cat $O/our_twist/summary.json
# tune + timing on large-twist
PPF_B200_PCG_TUNE=1 timeout 900 $H --lib $OUR --scene large-twist --steps 4 --out $O/our_large > $O/our_large.log 2>&1
grep -E "pcg-tune|step " $O/our_large.log; cat $O/our_large/summary.json
# launch list of one large-twist step, PCG kernels excluded
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(?!.*(k_spmv|k_update|k_direction)).*$" -c 600 --csv --log-file $O/launches_large.csv $H --lib $OUR --scene large-twist --steps 1 --out $O/ncu_large > $O/ncu_large.log 2>&1
python scripts/summarize_launches.py $O/launches_large.csv > $O/launches_large_summary.txt; head -40 $O/launches_large_summary.txt
rm -rf $O/*/state_* $O/*/final_* $O/*/dyn_*

#!/bin/bash
set -x
H=./ppf-contact-solver_b200/bin/ppf_host
OUR=ppf-contact-solver_b200/lib/libsimbackend_cuda.so
REF=oracle/_ref/libsimbackend_cuda.so
O=gpurun_out/prof1
mkdir -p $O
export PPF_HOST_QUIET=1 PPF_B200_QUIET=1
# kernel-variant timings on twist and large-twist
PPF_B200_PCG_TUNE=1 timeout 300 $H --lib $OUR --scene twist --steps 2 --out $O/tune_twist > $O/tune_twist.log 2>&1
PPF_B200_PCG_TUNE=1 timeout 600 $H --lib $OUR --scene large-twist --steps 1 --out $O/tune_large > $O/tune_large.log 2>&1
PPF_B200_PCG_TUNE=1 PPF_B200_NO_PERMUTE=1 timeout 600 $H --lib $OUR --scene large-twist --steps 1 --out $O/tune_large_noperm > $O/tune_large_noperm.log 2>&1
grep pcg-tune $O/*.log
# launch list of one twist step (ours)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_twist.csv $H --lib $OUR --scene twist --steps 1 --out $O/ncu_twist > $O/ncu_twist.log 2>&1
python scripts/summarize_launches.py $O/launches_twist.csv > $O/launches_twist_summary.txt
cat $O/launches_twist_summary.txt
rm -rf $O/*/state_* $O/*/final_* $O/*/dyn_*

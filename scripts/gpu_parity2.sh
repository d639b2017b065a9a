#!/bin/bash
set -x
H=./ppf-contact-solver_b200/bin/ppf_host
REF=oracle/_ref/libsimbackend_cuda.so
OUR=ppf-contact-solver_b200/lib/libsimbackend_cuda.so
O=gpurun_out/parity2
mkdir -p $O
export PPF_HOST_QUIET=1 PPF_B200_QUIET=1
oracle/_ref/ref_device_check 400000 > $O/device_check.log 2>&1; echo "exit=$?" >> $O/device_check.log
timeout 300 $H --lib $REF --scene fast-twist --steps 120 --save-every 10 --out $O/ref_ft > $O/ref_ft.log 2>&1
for st in 00020 00040 00050 00060 00070 00080 00090 00100 00110 00120; do
  timeout 120 $H --lib $REF --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/ref_1_$st > $O/ref_1_$st.log 2>&1
  timeout 120 $H --lib $OUR --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/our_1_$st > $O/our_1_$st.log 2>&1
  python scripts/compare_runs.py $O/ref_1_$st $O/our_1_$st > $O/cmp_$st.json 2>&1
done
cat $O/device_check.log
cat $O/cmp_*.json | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['pattern_equal'], r['dyn_nnz'], r['num_contact'], r['iter'], r['toi'], r['max_abs_pos_diff'], r.get('pattern_only_ref'), r.get('pattern_only_ours'))
"
rm -rf $O/ref_ft/state_* $O/*/final_* $O/*/state_*

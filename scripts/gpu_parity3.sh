#!/bin/bash
set -x
H=./ppf-contact-solver_b200/bin/ppf_host
REF=oracle/_ref/libsimbackend_cuda.so
OUR=ppf-contact-solver_b200/lib/libsimbackend_cuda.so
O=gpurun_out/parity3
mkdir -p $O
export PPF_HOST_QUIET=1 PPF_B200_QUIET=1
timeout 300 $H --lib $REF --scene fast-twist --steps 100 --save-every 20 --out $O/ref_ft > $O/ref_ft.log 2>&1
for st in 00020 00060 00100; do
  for k in 0 1; do
    mkdir -p $O/ref_d${k}_$st $O/our_d${k}_$st $O/ref2_d${k}_$st
    PPF_DUMP_LINSYS=$k PPF_DUMP_DIR=$O/ref_d${k}_$st timeout 120 $H --lib $REF --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/ref_d${k}_$st > $O/ref_d${k}_$st.log 2>&1
    PPF_DUMP_LINSYS=$k PPF_DUMP_DIR=$O/ref2_d${k}_$st timeout 120 $H --lib $REF --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/ref2_d${k}_$st > $O/ref2_d${k}_$st.log 2>&1
    PPF_DUMP_LINSYS=$k PPF_DUMP_DIR=$O/our_d${k}_$st timeout 120 $H --lib $OUR --scene fast-twist --steps 1 --load $O/ref_ft/state_$st.bin --out $O/our_d${k}_$st > $O/our_d${k}_$st.log 2>&1
    echo "--- state $st solve $k: ref vs ours" >> $O/linsys_cmp.txt
    python scripts/compare_linsys.py $O/ref_d${k}_$st/linsys_$k.bin $O/our_d${k}_$st/linsys_$k.bin >> $O/linsys_cmp.txt 2>&1
    echo "--- state $st solve $k: ref vs ref (run-to-run)" >> $O/linsys_cmp.txt
    python scripts/compare_linsys.py $O/ref_d${k}_$st/linsys_$k.bin $O/ref2_d${k}_$st/linsys_$k.bin >> $O/linsys_cmp.txt 2>&1
  done
done
cat $O/linsys_cmp.txt
mkdir -p $O/keep; cp $O/ref_ft/state_00060.bin $O/keep/; cp $O/ref_d0_00060/linsys_0.bin $O/keep/ref_linsys_0_00060.bin; cp $O/our_d0_00060/linsys_0.bin $O/keep/our_linsys_0_00060.bin
rm -rf $O/ref_ft $O/ref_d* $O/ref2_d* $O/our_d*/

#!/bin/bash
O=gpurun_out/r2n
mkdir -p $O
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
for cfg in "cpasync2 0" "cpasync2 1" "cpasync3 0"; do
  set -- $cfg
  PPF_B200_SPMV=$1 PPF_B200_L2_STREAM=$2 timeout 300 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene twist --steps 2 --out $O/tw --quiet > $O/tw_$1_$2.log 2>&1; echo "twist $cfg rc=$?"; grep -E "FATAL|done:" $O/tw_$1_$2.log | cut -c1-200
done
rm -rf $O/tw

#!/bin/bash
# rods on the GPU: reference golden for the woven scene, rod parity tests, first timings of woven / large-woven
set -x
O=gpurun_out/r1i
mkdir -p $O
timeout 120 bash tests/golden/make_golden.sh woven-only > $O/golden.log 2>&1; echo "golden rc=$?"
tail -n 3 $O/golden.log
if [ -f gpurun_out/golden/woven_n32/meta.json ]; then rm -rf tests/golden/woven_n32; cp -r gpurun_out/golden/woven_n32 tests/golden/; fi
timeout 400 python -m pytest tests -q -x -m gpu -k "woven or rod" > $O/pytest_rods.log 2>&1; echo "pytest rc=$?"
tail -n 15 $O/pytest_rods.log | cut -c1-300
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
timeout 120 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene woven --steps 10 --out $O/woven > $O/woven.log 2>&1; tail -n 4 $O/woven.log
timeout 120 $H --lib oracle/_ref/libsimbackend_cuda.so --scene woven --steps 10 --out $O/woven_ref > $O/woven_ref.log 2>&1; tail -n 4 $O/woven_ref.log
timeout 150 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-woven --steps 3 --out $O/lw > $O/lw.log 2>&1; tail -n 5 $O/lw.log
rm -rf $O/*/final_*.bin $O/*/state_final.bin $O/*/dyn_*.bin

#!/bin/bash
# final GPU pass of the round: reference golden for the strain-limited rod scene, the whole gpu test suite, smoke,
# large-woven on this backend and on the reference's CUDA build (same host, same DataSet, same 6 steps)
set -x
O=gpurun_out/r1j
mkdir -p $O
timeout 90 bash tests/golden/make_golden.sh woven-sl-only > $O/golden.log 2>&1; echo "golden rc=$?"
tail -n 3 $O/golden.log
if [ -f gpurun_out/golden/woven_sl_n32/meta.json ]; then rm -rf tests/golden/woven_sl_n32; cp -r gpurun_out/golden/woven_sl_n32 tests/golden/; fi
timeout 600 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 15 $O/pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
timeout 120 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-woven --steps 6 --out $O/lw > $O/lw.log 2>&1; tail -n 3 $O/lw.log
timeout 200 $H --lib oracle/_ref/libsimbackend_cuda.so --scene large-woven --steps 6 --out $O/lw_ref > $O/lw_ref.log 2>&1; tail -n 3 $O/lw_ref.log
rm -rf $O/*/final_*.bin $O/*/state_final.bin $O/*/dyn_*.bin

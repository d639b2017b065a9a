#!/bin/bash
# Round 2, second GPU call: the new parity tests (full-size live reference with pair lists, intersection positive path, late-feature goldens)
O=gpurun_out/r2b
mkdir -p $O
timeout 2400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_intersection.py tests/test_gpu_parity.py tests/test_zz_gpu_headless.py -q -m gpu -rA -p no:cacheprovider --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed|^E   |s call" $O/pytest_gpu.log | cut -c1-260 | tail -n 120
# why does the reference refuse the closed-window step? keep its log this time
H=./ppf-contact-solver_b200/bin/ppf_host
mkdir -p $O/cw
PPF_HOST_QUIET=1 timeout 300 $H --lib oracle/_ref/libsimbackend_cuda.so --scene headless-cw-closed --n 24 --steps 70 --out $O/cw > $O/cw.log 2>&1; echo "cw rc=$?"; tail -n 30 $O/cw.log | cut -c1-200
rm -rf $O/cw/data

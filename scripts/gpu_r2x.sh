#!/bin/bash
# Round 2 (1 GPU): compute-sanitizer beyond the smoke step: initcheck on smoke, memcheck over the feature tests (worker subprocesses included)
O=gpurun_out/r2x
mkdir -p $O
export PPF_B200_QUIET=1 PPF_HOST_QUIET=1
timeout 600 compute-sanitizer --tool initcheck --print-limit 30 --log-file $O/sanitizer_initcheck.log python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_initcheck.out 2>&1; echo "initcheck rc=$?"
tail -n 1 $O/smoke_initcheck.out | cut -c1-200
grep -E "ERROR SUMMARY|Uninitialized" $O/sanitizer_initcheck.log | sort | uniq -c | head
grep -A12 "Uninitialized" $O/sanitizer_initcheck.log | grep -E "at |by thread" | sort | uniq -c | sort -rn | head -12
timeout 1500 compute-sanitizer --tool memcheck --target-processes all --print-limit 30 --log-file $O/sanitizer_memcheck_%p.log python -m pytest tests/test_gpu_schwarz.py tests/test_gpu_rods.py tests/test_gpu_statistics.py tests/test_gpu_intersection.py tests/test_gpu_prims.py tests/test_gpu_abi_calls.py tests/test_zz_gpu_headless.py -q -m gpu -p no:cacheprovider -x > $O/pytest_memcheck.out 2>&1; echo "memcheck pytest rc=$?"
tail -n 3 $O/pytest_memcheck.out | cut -c1-300
cat $O/sanitizer_memcheck_*.log | grep -E "ERROR SUMMARY" | sort | uniq -c
cat $O/sanitizer_memcheck_*.log | grep -E "Invalid|out of bounds|misaligned" | sort | uniq -c | head

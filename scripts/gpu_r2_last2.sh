#!/bin/bash
# Round 2, last call: the GPU tests the previous call did not run, on the final library
O=gpurun_out/r2last
mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_schwarz.py tests/test_gpu_statistics.py tests/test_gpu_abi_calls.py tests/test_gpu_intersection.py tests/test_gpu_prims.py tests/test_zz_gpu_headless.py -q -m gpu -p no:cacheprovider > $O/pytest2.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest2.log

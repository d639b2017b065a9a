#!/bin/bash
# Round 2 (1 GPU): per-step ABI entry points and contact statistics against the live reference, plus a parity subset
O=gpurun_out/r2l
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_abi_calls.py tests/test_gpu_statistics.py tests/test_gpu_parity.py tests/test_gpu_intersection.py tests/test_zz_gpu_headless.py -q -m gpu -rA -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E   " $O/pytest.log | cut -c1-500 | tail -n 25

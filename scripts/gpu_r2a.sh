#!/bin/bash
# Round 2, first GPU call: the whole gpu suite without -x (every test reports), then reference golden for the late features.
O=gpurun_out/r2a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -rA -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed" $O/pytest_gpu.log | cut -c1-220 | tail -n 70
timeout 900 bash tests/golden/make_golden.sh next-round > $O/golden.log 2>&1; echo "golden rc=$?"; tail -n 25 $O/golden.log

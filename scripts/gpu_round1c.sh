#!/bin/bash
set -x
mkdir -p gpurun_out/r1c
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r1c/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r1c/pytest_gpu.log
PPF_B200_PCG_TUNE=1 PPF_HOST_QUIET=1 timeout 600 ./ppf-contact-solver_b200/bin/ppf_host --scene large-twist --steps 1 --out gpurun_out/r1c/tune_large > gpurun_out/r1c/tune_large.log 2>&1
grep "pcg-tune" gpurun_out/r1c/tune_large.log
rm -rf gpurun_out/r1c/tune_large
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/r1c/ours.json 2> gpurun_out/r1c/ours.err
cat gpurun_out/r1c/ours.json

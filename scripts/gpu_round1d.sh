#!/bin/bash
set -x
mkdir -p gpurun_out/r1d
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r1d/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 15 gpurun_out/r1d/pytest_gpu.log
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/r1d/ours.json 2> gpurun_out/r1d/ours.err
cat gpurun_out/r1d/ours.json
PPF_B200_SWEPT_HULL=1 PPF_B200_NO_CAND_REUSE=1 timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/r1d/ours_hull.json 2> gpurun_out/r1d/ours_hull.err
cat gpurun_out/r1d/ours_hull.json

#!/bin/bash
set -x
mkdir -p gpurun_out/final
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 6 gpurun_out/final/pytest_gpu.log | cut -c1-250
timeout 600 python bench.py > gpurun_out/final/ours.json 2> gpurun_out/final/ours.err; tail -n 1 gpurun_out/final/ours.json | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/final/ref.json 2> gpurun_out/final/ref.err; tail -n 1 gpurun_out/final/ref.json | cut -c1-300

#!/bin/bash
set -x
mkdir -p gpurun_out/r1h
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r1h/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/r1h/pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --impl reference > gpurun_out/r1h/ref.json 2> gpurun_out/r1h/ref.err; tail -n 1 gpurun_out/r1h/ref.json | cut -c1-400
timeout 900 python bench.py > gpurun_out/r1h/ours.json 2> gpurun_out/r1h/ours.err; tail -n 1 gpurun_out/r1h/ours.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2

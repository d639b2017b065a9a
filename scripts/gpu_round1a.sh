#!/bin/bash
# golden fixtures from the reference build + first full bench lines (both arms)
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
nproc
bash tests/golden/make_golden.sh > gpurun_out/golden.log 2>&1
tail -n 5 gpurun_out/golden.log
mkdir -p gpurun_out/bench
timeout 900 python bench.py --impl reference --steps 6 --warmup 3 > gpurun_out/bench/ref.json 2> gpurun_out/bench/ref.err
cat gpurun_out/bench/ref.json
timeout 1200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench/ours.json 2> gpurun_out/bench/ours.err
cat gpurun_out/bench/ours.json
tail -n 5 gpurun_out/bench/ours.err

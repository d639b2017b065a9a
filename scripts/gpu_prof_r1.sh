#!/bin/bash
# round-1 profiles of the bench command: launch list (per-kernel shares) + full-set capture of the PCG kernels
set -x
mkdir -p gpurun_out/prof_r1
timeout 2400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/prof_r1/launches_bench.csv \
   python bench.py --steps 2 --warmup 3 > gpurun_out/prof_r1/ncu_bench.log 2>&1
python scripts/summarize_launches.py gpurun_out/prof_r1/launches_bench.csv > gpurun_out/prof_r1/launches_bench_summary.txt 2>&1; head -n 30 gpurun_out/prof_r1/launches_bench_summary.txt
rm -f gpurun_out/prof_r1/launches_bench.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_staged|k_update|k_direction" -s 6000 -c 6 -o gpurun_out/prof_r1/pcg_full \
   python bench.py --steps 1 --warmup 3 > gpurun_out/prof_r1/ncu_pcg.log 2>&1
ls -la gpurun_out/prof_r1

#!/bin/bash
# Round 2, closing: ncu launch list of the headline bench command with the final kernels (shares, not absolutes)
O=gpurun_out/r2launch
mkdir -p $O
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-extra > $O/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/summarize_launches.py $O/launches_bench.csv > $O/launches_bench_summary.txt 2>&1; head -24 $O/launches_bench_summary.txt

#!/bin/bash
# GPU tests + smoke + ncu launch list (twist scene, short) + one --set full capture of k_spmv
set -x
mkdir -p gpurun_out/r1b
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r1b/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/r1b/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r1b/smoke.log
# launch list of the bench command on a reduced scene so that the list stays small (same code path):
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r1b/launches_twist.csv \
   python bench.py --scene twist --steps 1 --warmup 3 > gpurun_out/r1b/ncu_bench_twist.log 2>&1
python scripts/summarize_launches.py gpurun_out/r1b/launches_twist.csv > gpurun_out/r1b/launches_twist_summary.txt 2>&1; head -n 40 gpurun_out/r1b/launches_twist_summary.txt
# full-set capture of the SpMV on the headline scene (3 launches after 2000 warm launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 2000 -c 3 -o gpurun_out/r1b/spmv_full \
   python bench.py --steps 1 --warmup 3 > gpurun_out/r1b/ncu_spmv.log 2>&1
ls -la gpurun_out/r1b

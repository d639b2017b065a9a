#!/bin/bash
set -x
mkdir -p gpurun_out/r1g
bash tests/golden/make_golden.sh hang-only > gpurun_out/r1g/golden.log 2>&1; tail -n 2 gpurun_out/r1g/golden.log
cp -r gpurun_out/golden/hang_n32 tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r1g/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 30 gpurun_out/r1g/pytest_gpu.log | cut -c1-250
PPF_B200_PCG_TUNE=1 PPF_HOST_QUIET=1 timeout 600 ./ppf-contact-solver_b200/bin/ppf_host --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene large-twist --steps 1 --out gpurun_out/r1g/tune > gpurun_out/r1g/tune.log 2>&1
grep "pcg-tune" gpurun_out/r1g/tune.log; rm -rf gpurun_out/r1g/tune

#!/bin/bash
set -x
mkdir -p gpurun_out/r1e
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r1e/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/r1e/pytest_gpu.log
# candidate-reuse verification on a bigger contact-rich run (fast-twist n=48, 100 steps)
PPF_B200_VERIFY_CAND=1 PPF_HOST_QUIET=1 PPF_B200_QUIET=1 timeout 600 ./ppf-contact-solver_b200/bin/ppf_host --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene fast-twist --steps 100 --out gpurun_out/r1e/verify48 > gpurun_out/r1e/verify48.log 2>&1; echo "verify rc=$?"
grep -c "verify-cand" gpurun_out/r1e/verify48.log; tail -n 3 gpurun_out/r1e/verify48.log; rm -rf gpurun_out/r1e/verify48

#!/bin/bash
set -x
mkdir -p gpurun_out/r1f
bash tests/golden/make_golden.sh five-only > gpurun_out/r1f/golden.log 2>&1; tail -n 3 gpurun_out/r1f/golden.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r1f/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/r1f/pytest_gpu.log
H=./ppf-contact-solver_b200/bin/ppf_host
export PPF_HOST_QUIET=1 PPF_B200_QUIET=1
for scene in twist five-twist; do
  for lib in oracle/_ref/libsimbackend_cuda.so ppf-contact-solver_b200/lib/libsimbackend_cuda.so; do
    tag=$(echo $lib | cut -c1-3)
    timeout 900 $H --lib $lib --scene $scene --steps 60 --out gpurun_out/r1f/${scene}_$tag > gpurun_out/r1f/${scene}_$tag.log 2>&1
    tail -n 1 gpurun_out/r1f/${scene}_$tag.log
    cp gpurun_out/r1f/${scene}_$tag/summary.json gpurun_out/r1f/${scene}_${tag}_summary.json; rm -rf gpurun_out/r1f/${scene}_$tag
  done
done

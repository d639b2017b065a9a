#!/bin/bash
# First GPU probe: run the reference build through our host on a few scenes.
set -x
H=./ppf-contact-solver_b200/bin/ppf_host
REF=oracle/_ref/libsimbackend_cuda.so
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
export PPF_HOST_QUIET=1
timeout 600 $H --lib $REF --scene fast-twist --steps 120 --save-every 10 --out gpurun_out/ref_fast_twist > gpurun_out/ref_fast_twist.log 2>&1
echo "exit=$?" >> gpurun_out/ref_fast_twist.log
timeout 600 $H --lib $REF --scene twist --steps 30 --save-every 10 --out gpurun_out/ref_twist > gpurun_out/ref_twist.log 2>&1
echo "exit=$?" >> gpurun_out/ref_twist.log
timeout 900 $H --lib $REF --scene large-twist --steps 4 --out gpurun_out/ref_large_twist > gpurun_out/ref_large_twist.log 2>&1
echo "exit=$?" >> gpurun_out/ref_large_twist.log
# linear-system dumps (assembled A, b) for PCG parity: first solve of fast-twist and one late solve
PPF_DUMP_LINSYS=0 PPF_DUMP_DIR=gpurun_out timeout 300 $H --lib $REF --scene fast-twist --steps 2 --out gpurun_out/ref_dump0 > gpurun_out/ref_dump0.log 2>&1
rm -f gpurun_out/ref_large_twist/*.bin gpurun_out/ref_twist/state_000*.bin
tail -3 gpurun_out/*.log
ls -la gpurun_out gpurun_out/*
du -sh gpurun_out

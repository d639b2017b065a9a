#!/bin/bash
# analytic colliders on the GPU: reference golden for the drape scene, then the whole gpu suite
set -x
O=gpurun_out/r1l
mkdir -p $O
timeout 45 bash tests/golden/make_golden.sh drape-only > $O/golden.log 2>&1; echo "golden rc=$?"
tail -n 2 $O/golden.log
if [ -f gpurun_out/golden/drape_n32/meta.json ]; then rm -rf tests/golden/drape_n32; cp -r gpurun_out/golden/drape_n32 tests/golden/; fi
timeout 110 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 $O/pytest_gpu.log | cut -c1-400

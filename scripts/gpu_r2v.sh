#!/bin/bash
# Round 2 (1 GPU): the whole GPU suite twice (flake check), as the driver runs it
O=gpurun_out/r2v
mkdir -p $O
for rep in 1 2; do
  timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rfEs > $O/pytest_full_$rep.log 2>&1; echo "pytest #$rep rc=$?"
  grep -E "^(FAILED|ERROR|SKIPPED)|passed|failed" $O/pytest_full_$rep.log | cut -c1-300 | tail -n 12
done

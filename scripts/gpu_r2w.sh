#!/bin/bash
# Round 2 (1 GPU): compute-sanitizer over the smoke step (contact-rich fast-twist state through the C ABI): memcheck, racecheck, synccheck
O=gpurun_out/r2w
mkdir -p $O
export PPF_B200_QUIET=1 PPF_HOST_QUIET=1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file $O/sanitizer_$tool.log python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$tool.out 2>&1; echo "$tool rc=$?"
  tail -n 1 $O/smoke_$tool.out | cut -c1-200
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" $O/sanitizer_$tool.log | sort | uniq -c | head -12
done

#!/bin/bash
# Round 2, seventh GPU call (1 GPU): bench lines of the five BASELINE.json configurations on both arms, then the ncu evidence for
# the headline: launch list of the bench command and one --set full capture of the three PCG kernels.
O=gpurun_out/r2g
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_intersection.py tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "scope or fast_twist_n16 or drape" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest.log
for scene in hang twist five-twist large-woven large-twist; do
  steps=20; [ $scene = large-twist ] && steps=6; [ $scene = large-woven ] && steps=6; [ $scene = hang ] && steps=60
  timeout 900 python bench.py --scene $scene --steps $steps --warmup 3 > $O/bench_ours_$scene.json 2> $O/bench_ours_$scene.err; echo "ours $scene rc=$?"
  timeout 1200 python bench.py --impl reference --scene $scene --steps $steps --warmup 3 > $O/bench_ref_$scene.json 2> $O/bench_ref_$scene.err; echo "ref $scene rc=$?"
  python - $O $scene <<'PY'
import json, sys
O, sc = sys.argv[1:3]
try:
    a = json.loads(open("%s/bench_ours_%s.json" % (O, sc)).read().strip().splitlines()[-1])
    b = json.loads(open("%s/bench_ref_%s.json" % (O, sc)).read().strip().splitlines()[-1])
    print("%-12s ours %9.2f ms/step (e2e %.4f f/s)  ref %9.2f ms/step   ratio %.2fx   newton solves %d / %d  pcg iters %d / %d  contacts(last) %s / %s  roofline %.3f" % (
        sc, a["ms_per_step"], a["e2e"]["value"], b["ms_per_step"], b["ms_per_step"] / a["ms_per_step"], a["work"]["newton_solves"], b["work"]["newton_solves"],
        a["work"]["pcg_iters"], b["work"]["pcg_iters"], a["work"]["num_contact_last"], b["work"]["num_contact_last"], a.get("roofline", {}).get("frac", 0)))
except Exception as e:
    print(sc, "failed", e)
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 3 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/summarize_launches.py $O/launches_bench.csv > $O/launches_bench_summary.txt 2>/dev/null; head -n 16 $O/launches_bench_summary.txt
rm -f $O/launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_spmv_staged|k_update|k_direction' -s 6000 -c 6 -o $O/pcg_full python bench.py --steps 1 --warmup 3 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/*.ncu-rep

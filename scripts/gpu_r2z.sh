#!/bin/bash
# Round 2, closing call (1 GPU): smoke(), both bench arms as the driver runs them (reference first), the small configurations on both arms
O=gpurun_out/r2z
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
t0=$(date +%s)
timeout 1500 python bench.py --impl reference > $O/bench_reference_large-twist.json 2> $O/bench_reference_large-twist.err; echo "reference arm rc=$? in $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 1500 python bench.py > $O/bench_ours_large-twist.json 2> $O/bench_ours_large-twist.err; echo "our arm rc=$? in $(( $(date +%s) - t0 )) s"
for scene in hang five-twist twist; do
  steps=20; [ $scene = hang ] && steps=60
  timeout 600 python bench.py --impl reference --scene $scene --steps $steps --warmup 3 > $O/bench_reference_$scene.json 2> $O/bench_reference_$scene.err
  timeout 600 python bench.py --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_ours_$scene.json 2> $O/bench_ours_$scene.err
done
python - <<'PY'
import json
O = "gpurun_out/r2z"
def L(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
for sc in ("hang", "five-twist", "twist", "large-twist"):
    a, b = L("%s/bench_ours_%s.json" % (O, sc)), L("%s/bench_reference_%s.json" % (O, sc))
    if not a or not b: print(sc, "missing", bool(a), bool(b)); continue
    print("%-12s ours %8.2f ms/step (value %.4f, e2e %.4f)  ref %8.2f ms/step (value %.4f)  ratio %.2fx  e2e ratio %.2fx  solves %s / %s  iters %s / %s  frac %.3f  launches %s  clocks %s" % (
        sc, a["ms_per_step"], a["value"], a["e2e"]["value"], b["ms_per_step"], b["value"], b["ms_per_step"] / a["ms_per_step"], a["e2e"]["value"] / b["e2e"]["value"],
        a["work"]["newton_solves"], b["work"]["newton_solves"], a["work"]["pcg_iters"], b["work"]["pcg_iters"], a["roofline"]["frac"], a.get("gpu_launches"), a["clocks"]))
a = L(O + "/bench_ours_large-twist.json")
if a: print("cpu_baseline", a.get("cpu_baseline")); print("schwarz_mode", a.get("schwarz_mode")); print("keys", sorted(a.keys()))
PY

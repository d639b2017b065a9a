#!/bin/bash
# Round 2 (2 GPUs): the 2-GPU tests on the reworked iteration chain, large-woven sharded over 2 GPUs, large-twist at N=2 with / without chained launches
O=gpurun_out/r2u
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -rA -p no:cacheprovider > $O/pytest_multi.log 2>&1; echo "pytest rc=$?"
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed|^E   " $O/pytest_multi.log | cut -c1-400 | tail -n 20
run() { # name, scene, steps, env...
  name=$1; scene=$2; steps=$3; shift 3
  env "$@" timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --scene $scene --steps $steps --warmup 3 --no-extra > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name rc=$?"
  python - "$O/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("%s: %.1f ms/step, %.1f us/PCG iter, iters %d, contacts %s, phases %s, value %.4f" % (sys.argv[2], d["ms_per_step"],
          1e3 * d["phase_ms_per_step"]["linsolve"] * d["steps"] / max(1, d["work"]["pcg_iters"]), d["work"]["pcg_iters"], d["work"].get("num_contact_last"), {k: round(v, 1) for k, v in d["phase_ms_per_step"].items()}, d["value"]))
except Exception as e:
    print("bench failed", sys.argv[1:], e)
PY
}
run n2_large-woven large-woven 6 PPF_B200_DIST_CONTACT=1
tail -n 3 $O/bench_n2_large-woven.err | cut -c1-300
run n2_large-twist large-twist 4 PPF_B200_DIST_CONTACT=1
run n2_large-twist_pdl large-twist 4 PPF_B200_PDL=1

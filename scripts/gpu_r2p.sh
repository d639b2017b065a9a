#!/bin/bash
# where does the erratic contact-phase time of twist come from? per-step allocations and phase times
O=gpurun_out/r2p
mkdir -p $O
export PPF_HOST_QUIET=1
H=./ppf-contact-solver_b200/bin/ppf_host
for rep in 1 2; do
  PPF_B200_CONTACT_PROF=1 timeout 300 $H --lib ppf-contact-solver_b200/lib/libsimbackend_cuda.so --scene twist --steps 23 --out $O/tw$rep > $O/tw$rep.log 2>&1
  python - $O/tw$rep <<'PY'
import sys, os
d = sys.argv[1] + "/data"
def rd(ch):
    p = os.path.join(d, "advance.%s.out" % ch)
    return [float(l.split()[1]) for l in open(p)] if os.path.exists(p) else []
print("device-alloc per step:", [int(v) for v in rd("device-alloc")])
print("num_contact:", [int(v) for v in rd("num_contact")][::2])
PY
  grep "contact-prof" $O/tw$rep.log | awk '{print $3, $4, $5, $(NF-5), $(NF-4), $(NF-3), $(NF-1), $NF}' | tr '\n' ';' | cut -c1-1500; echo
  rm -rf $O/tw$rep/data
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_smoke_entry.py -q -m gpu -x -p no:cacheprovider -k "twist or woven_n32 or drape_n32 or smoke" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 $O/pytest.log
for rep in 1 2; do
  timeout 600 python bench.py --scene twist --steps 20 --warmup 3 --no-extra > $O/bench_twist_$rep.json 2> $O/bench_twist_$rep.err
  python - "$O/bench_twist_$rep.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("twist: %.2f ms/step phases %s" % (d["ms_per_step"], {k: round(v, 2) for k, v in d["phase_ms_per_step"].items()}))
PY
done

"""bisect helper: one golden case under several env switches; prints logs and deviation from the reference"""
import json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
name, idx = sys.argv[1], int(sys.argv[2])
g = os.path.join(REPO, "tests", "golden", name); meta = json.load(open(g + "/meta.json")); case = meta["cases"][idx]
ref = np.fromfile(os.path.join(g, case["dir"], "final_curr.bin"), dtype=np.float32).reshape(-1, 3)
logs = json.load(open(os.path.join(g, case["dir"], "logs.json")))
print("reference: num_contact", logs["num_contact"], "iter", logs["iter"], "toi", logs["toi"])
combos = [{}, {"PPF_B200_NO_CAND_REUSE": "1"}, {"PPF_B200_SWEPT_HULL": "1"}, {"PPF_B200_SWEPT_HULL": "1", "PPF_B200_NO_CAND_REUSE": "1"}, {"PPF_B200_SPMV_LEGACY": "1"},
          {"PPF_B200_SWEPT_HULL": "1", "PPF_B200_NO_CAND_REUSE": "1", "PPF_B200_SPMV_LEGACY": "1"}]
for env in combos:
    for rep in range(2):
        job = {"lib": "ours", "scene": meta["scene"], "n": meta["n"], "state": os.path.join(g, case["state"]), "steps": 1, "env": env, "out": "/tmp/bisect_job"}
        json.dump(job, open("/tmp/bisect_job.json", "w"))
        r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "_worker.py"), "/tmp/bisect_job.json"], capture_output=True, text=True)
        if r.returncode: print(env, "FAILED", r.stderr[-500:]); continue
        d = np.load("/tmp/bisect_job/result.npz")
        print(sorted(env), "pos diff %.3e" % np.abs(d["curr"] - ref).max(), "num_contact", d["log_num_contact"], "iter", d["log_iter"], "toi", d["log_toi"])

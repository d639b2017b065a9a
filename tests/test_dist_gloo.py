"""The N>1 path on CPU: world_size 2 over gloo (127.0.0.1).  Each rank steps its own replica of a
small scene through the ABI (the CPU oracle library stands in for the GPU backend) and the
measurements are combined exactly as bench.py does (replicas.Ranks): max over ranks of the times,
sum of the frames."""
import json
import os
import subprocess
import sys

from conftest import PKG, REPO

WORKER = r'''
import json, os, sys, time
sys.path.insert(0, %(pkg)r)
import ppf_b200, replicas
os.environ["PPF_HOST_QUIET"] = "1"
ranks = replicas.Ranks("gloo")
s = ppf_b200.Session("fast-twist", n=8 + 4 * ranks.rank, lib=ppf_b200.LIB_ORACLE, log_dir="/tmp/ppf_gloo_%%d" %% ranks.rank)
ranks.barrier()
t0 = time.perf_counter(); t_sim0 = s.time
for _ in range(2 + ranks.rank):
    ok, ms = s.step(); assert ok
wall = time.perf_counter() - t0
frames = (s.time - t_sim0) * 60.0
ranks.barrier()
t_dev, t_e2e, tot = ranks.reduce(wall, wall * 2, frames)
print(json.dumps({"rank": ranks.rank, "world": ranks.world, "wall": wall, "frames": frames, "t_dev": t_dev, "t_e2e": t_e2e, "tot": tot}))
ranks.close()
'''


def test_two_rank_replicas_over_gloo(built, tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"pkg": PKG})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-3000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    a, b = sorted(outs, key=lambda d: d["rank"])
    assert a["world"] == b["world"] == 2
    # both ranks hold the same reduced numbers: max of the times, sum of the frames
    assert abs(a["t_dev"] - max(a["wall"], b["wall"])) < 1e-9 and a["t_dev"] == b["t_dev"]
    assert abs(a["t_e2e"] - 2 * max(a["wall"], b["wall"])) < 1e-9
    assert abs(a["tot"] - (a["frames"] + b["frames"])) < 1e-9 and a["tot"] == b["tot"]
    assert abs(a["frames"] - 2 * 0.9999) < 1e-3 and abs(b["frames"] - 3 * 0.9999) < 1e-3


def test_single_rank_is_a_no_op(built):
    sys.path.insert(0, PKG)
    import replicas

    env = {k: os.environ.pop(k) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK") if k in os.environ}
    try:
        r = replicas.Ranks()
        assert r.world == 1 and r.reduce(1.0, 2.0, 3.0) == (1.0, 2.0, 3.0)
        r.barrier()
    finally:
        os.environ.update(env)

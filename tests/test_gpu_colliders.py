"""Analytic colliders (invisible sphere + floor with friction) on a real B200: the `drape` scene.

Against the reference: tests/golden/drape_n32 (when present) is picked up by the parametrised golden tests of
test_gpu_parity.py.  Here: CUDA backend vs CPU oracle on the same steps, and no-tunnelling invariants at a larger size.
"""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_drape_first_contacts_match_the_oracle(run_job):
    job = {"scene": "drape", "n": 32, "state": None, "steps": 5}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert json.loads(str(g["stats_json"]))["n_kernel_launches"] > 0
    assert len(g["log_num_contact"]) == len(o["log_num_contact"]) and o["log_num_contact"].max() > 20
    first = int(np.argmax(o["log_num_contact"] > 0))
    assert list(g["log_num_contact"][:first + 1]) == list(o["log_num_contact"][:first + 1])   # free fall, then the same first touch
    for mine, theirs in zip(g["log_num_contact"], o["log_num_contact"]):
        assert abs(mine - theirs) <= 3 + 0.02 * theirs
    for mine, theirs in zip(g["log_toi"], o["log_toi"]):
        assert abs(mine - theirs) < 2e-2
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))


def test_drape_never_tunnels(run_job):
    r = run_job({"lib": "ours", "scene": "drape", "n": 96, "state": None, "steps": 60}, force_subprocess=True)
    assert bool(r["ok"].all()) and np.isfinite(r["curr"]).all()
    x = r["curr"].astype(np.float64)
    assert np.linalg.norm(x, axis=1).min() > 0.4 and x[:, 1].min() > -0.08
    assert r["log_num_contact"].max() > 500 and (r["log_toi"] > 0).all() and (r["log_toi"] <= 1).all()

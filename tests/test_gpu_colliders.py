"""Analytic colliders (invisible sphere + floor with friction) on a real B200: the `drape` scene.

Against the reference: tests/golden/drape_n32 (when present) is picked up by the parametrised golden tests of
test_gpu_parity.py.  Here: CUDA backend vs CPU oracle on the same steps, and no-tunnelling invariants at a larger size.
"""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_drape_first_contacts_match_the_oracle(run_job):
    """Five chained steps from rest: free fall, first touch of the sphere, then ~15 Newton iterations that the collider
    line search cuts to a few percent each.  Whether a vertex sits inside the 1e-3 contact gap at the first touch is an
    fp32 tie between the two builds (observed on a B200: 1 vs 2 vertices, after which this backend needs 21 Newton
    iterations and the oracle 20), so the iteration-by-iteration logs are compared in aggregate; exact agreement with the
    reference is the job of the golden cases drape_n32 in test_gpu_parity.py."""
    job = {"scene": "drape", "n": 32, "state": None, "steps": 5}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert json.loads(str(g["stats_json"]))["n_kernel_launches"] > 0
    assert abs(len(g["log_num_contact"]) - len(o["log_num_contact"])) <= 4 and o["log_num_contact"].max() > 20
    first_g, first_o = int(np.argmax(g["log_num_contact"] > 0)), int(np.argmax(o["log_num_contact"] > 0))
    assert abs(first_g - first_o) <= 1 and first_o > 0            # the same free fall before the first touch
    assert abs(g["log_num_contact"][-1] - o["log_num_contact"][-1]) <= 3 + 0.25 * o["log_num_contact"][-1]
    assert (g["log_toi"] > 0).all() and (g["log_toi"] <= 1).all() and g["log_toi"].min() < 0.5   # the collider line search cuts steps
    # a cut step advances time by dt * toi_advanced, so the two runs are at slightly different times (and the falling sheet at
    # slightly different heights): compare the clocks, and the positions only against the colliders
    assert abs(float(g["time"][0]) - float(o["time"][0])) < 0.3 * float(o["time"][0])
    x = g["curr"].astype(np.float64)
    assert np.isfinite(x).all() and np.linalg.norm(x, axis=1).min() > 0.4 and x[:, 1].min() > -0.08   # outside the sphere, above the floor


def test_drape_never_tunnels(run_job):
    r = run_job({"lib": "ours", "scene": "drape", "n": 96, "state": None, "steps": 60}, force_subprocess=True)
    assert bool(r["ok"].all()) and np.isfinite(r["curr"]).all()
    x = r["curr"].astype(np.float64)
    assert np.linalg.norm(x, axis=1).min() > 0.4 and x[:, 1].min() > -0.08
    assert r["log_num_contact"].max() > 500 and (r["log_toi"] > 0).all() and (r["log_toi"] <= 1).all()

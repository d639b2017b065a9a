"""Analytic colliders (invisible sphere + floor with friction) on a real B200: the `drape` scene.

Against the reference: tests/golden/drape_n32 (when present) is picked up by the parametrised golden tests of
test_gpu_parity.py.  Here: no-tunnelling invariants at a larger size; the comparison of chained steps with the CPU oracle lives in
test_zz_gpu_headless.py (rewritten after its last GPU run, so it sorts with the tests that have not run on a GPU yet).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_drape_never_tunnels(run_job):
    r = run_job({"lib": "ours", "scene": "drape", "n": 96, "state": None, "steps": 60}, force_subprocess=True)
    assert bool(r["ok"].all()) and np.isfinite(r["curr"]).all()
    x = r["curr"].astype(np.float64)
    assert np.linalg.norm(x, axis=1).min() > 0.4 and x[:, 1].min() > -0.08
    assert r["log_num_contact"].max() > 500 and (r["log_toi"] > 0).all() and (r["log_toi"] <= 1).all()

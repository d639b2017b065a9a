"""The per-step entry points of the ABI that a host calls between steps (main.cu:2079-2228) on a real B200, against the LIVE
reference build and the CPU oracle: override_velocity, override_angular_velocity, gather_current_positions (velocity
keyframes of S/backend.rs) and update_rest_shape (streamed rest shapes + the per-face exclusion mask).
Bars: the gathered positions and the rewritten `prev` right after the calls are bit-exact (they are copies and single fp32
expressions; 1 ulp is allowed for `prev`, ptxas may fuse `p - vw*dt` differently per build); the step that follows lands within
the usual position bars (90 % 5e-5, worst 5e-4) of the reference and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
IDX = list(range(40, 120))


def _libs(ppf):
    return ["ours", "oracle"] + (["reference"] if os.path.exists(ppf.LIB_REF) else [])


def _close(a, b):
    d = np.abs(a - b).max(axis=1)
    return float(np.percentile(d, 90)) < 5e-5 and float(d.max()) < 5e-4


def test_velocity_overrides_match_the_reference(run_job, ppf):
    dt = 0.01
    calls = [["gather", IDX], ["override_velocity", IDX, [0.3, -0.2, 0.1], dt], ["override_angular_velocity", IDX, [0.0, 0.0, 2.0], [0.1, 0.2, 0.0], dt]]
    r = {lib: run_job({"lib": lib, "scene": "sheet", "n": 16, "state": None, "steps": 2, "abi_calls": calls}, force_subprocess=True) for lib in _libs(ppf)}
    g = r["ours"]
    assert float(np.abs(g["after_calls_prev"][IDX] - g["after_calls_curr"][IDX]).max()) > 1e-3      # the overrides did something
    for lib, o in r.items():
        assert bool(o["ok"].all())
        assert np.array_equal(g["gather0"], o["gather0"]) and np.array_equal(g["after_calls_curr"], o["after_calls_curr"])
        assert float(np.abs(g["after_calls_prev"] - o["after_calls_prev"]).max()) <= 2.4e-7, lib   # 1 ulp at |x| <= 2
        assert _close(g["curr"], o["curr"]), lib


def test_update_rest_shape_matches_the_reference(run_job, ppf):
    calls = [["update_rest_shape", 1.25, list(range(0, 450, 3))]]
    r = {lib: run_job({"lib": lib, "scene": "sheet", "n": 16, "state": None, "steps": 3, "abi_calls": calls}, force_subprocess=True) for lib in _libs(ppf)}
    plain = run_job({"lib": "ours", "scene": "sheet", "n": 16, "state": None, "steps": 3}, force_subprocess=True)
    g = r["ours"]
    assert float(np.abs(g["curr"] - plain["curr"]).max()) > 1e-3                                      # the new rest shape is felt
    for lib, o in r.items():
        assert bool(o["ok"].all()) and len(o["log_iter"]) == len(g["log_iter"])
        assert _close(g["curr"], o["curr"]), lib

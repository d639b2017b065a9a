"""Positive path of check_intersection on a real B200 (contact.cu:3056-3175, main.cu:247-263, 2341):
initialize() on a deliberately intersecting start state must return false with fatal code 1 and the SAME intersection
records -- element types, indices and the elements' positions verbatim -- as the CPU oracle and as the live reference build.
Records are compared as sorted sets: the slot order in the 256-entry buffer is the order of the atomics."""
import os
import sys

import numpy as np
import pytest

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "tests"))
import pierce  # noqa: E402

pytestmark = pytest.mark.gpu


def _records(run_job, lib, scene, n, state):
    r = run_job({"lib": lib, "scene": scene, "n": n, "state": state, "steps": 0, "allow_init_failure": True}, force_subprocess=True)
    return bool(r["init_ok"][0]), int(r["fatal_code"][0]), r["isect"]


@pytest.mark.parametrize("case", ["pierced-curtain", "crowded-rods"])
def test_intersecting_start_state_is_refused_with_the_same_records(run_job, ppf, tmp_path, case):
    state = str(tmp_path / "bad.bin")
    if case == "pierced-curtain":
        pierce.pierced_curtain(ppf, ppf.LIB_ORACLE, state)
        scene, n, kinds = "headless", 24, {0}
    else:
        pierce.crowded_rods(ppf, ppf.LIB_ORACLE, state)
        scene, n, kinds = "woven", 32, {1, 3}
    ok_g, code_g, rec_g = _records(run_job, "ours", scene, n, state)
    ok_o, code_o, rec_o = _records(run_job, "oracle", scene, n, state)
    assert not ok_g and code_g == 1 and not ok_o and code_o == 1
    assert len(rec_g) >= len(kinds) and set(rec_g[:, 0].astype(int)) == kinds
    assert rec_g.shape == rec_o.shape and np.array_equal(rec_g, rec_o)          # indices and positions: exact
    if os.path.exists(ppf.LIB_REF):
        ok_r, code_r, rec_r = _records(run_job, "reference", scene, n, state)
        assert not ok_r and code_r == 1
        assert rec_g.shape == rec_r.shape and np.array_equal(rec_g, rec_r)
    # the log carries the marker the frontend classifies failures by (main.cu:248)
    # ... and the untouched start state is accepted
    r = run_job({"lib": "ours", "scene": scene, "n": n, "state": None, "steps": 0, "allow_init_failure": True}, force_subprocess=True)
    assert bool(r["init_ok"][0]) and len(r["isect"]) == 0


def test_scope_gate_refuses_without_killing_the_process(run_job):
    """A dataset with something this backend does not implement (here: barrier-held fix pins, ParamSet.disable_pin_dof_removal)
    is refused: initialize() returns false with fatal code 4 and the reason in ppf_fatal_detail -- the worker process lives on
    to report it (round 1 called exit(1) here)."""
    r = run_job({"lib": "ours", "scene": "fast-twist", "n": 16, "state": None, "steps": 0, "allow_init_failure": True,
                 "param_flags": {"disable_pin_dof_removal": 1}}, force_subprocess=True)
    assert not bool(r["init_ok"][0]) and int(r["fatal_code"][0]) == 4
    assert "disable_pin_dof_removal" in str(r["fatal_detail"])

"""Pins the WHOLE-STEP oracle (oracle/ppf_oracle.cpp, CPU) against outputs of the real reference.

tests/golden/fast_twist_n*/ were produced by the unmodified reference backend on a B200
(tests/golden/make_golden.sh): a contact-rich state, the first assembled Newton system of the next
step, and the result of that step.  The oracle must reproduce them from the same state:
  * block pattern of the assembled system: exact, except for blocks touched by tie-decided pairs
    (fp32 ties that FMA contraction on the GPU resolves differently; <= 0.2 % of the blocks);
  * block values / right-hand side: 1e-3 relative (fp32 atomics ordering noise is ~1e-5) on all but
    the <= 0.2 % of blocks / 1 % of rows a tie-decided pair touches;
  * per-Newton-iteration contact counts: |delta| <= 3 + 1e-3 N;  PCG iteration counts: 5 %;
  * positions after the step: 5e-4 absolute (scene size ~1; the reference itself varies by ~1e-5
    run to run because of fp32 atomics).
"""
import json
import os
import sys

import numpy as np
import pytest

from conftest import REPO, golden_cases

sys.path.insert(0, os.path.join(REPO, "scripts"))
from compare_linsys import compare  # noqa: E402

POS_TOL = 5e-4


@pytest.mark.parametrize("name,scene,n,state,refdir", golden_cases(), ids=lambda v: os.path.basename(v) if isinstance(v, str) and os.sep in v else None)
def test_oracle_reproduces_reference_step(run_job, name, scene, n, state, refdir):
    r = run_job({"lib": "oracle", "scene": scene, "n": n, "state": state, "steps": 1, "dump_linsys": True})
    assert bool(r["ok"].all())
    logs = json.load(open(os.path.join(refdir, "logs.json")))
    ref = np.fromfile(os.path.join(refdir, "final_curr.bin"), dtype=np.float32).reshape(-1, 3)
    assert r["curr"].shape == ref.shape
    assert float(np.abs(r["curr"] - ref).max()) < POS_TOL
    assert len(r["log_num_contact"]) == len(logs["num_contact"])
    for mine, theirs in zip(r["log_num_contact"], logs["num_contact"]):
        assert abs(mine - theirs) <= 3 + 1e-3 * theirs
    for mine, theirs in zip(r["log_iter"], logs["iter"]):
        assert abs(mine - theirs) <= 2 + 0.05 * theirs
    for mine, theirs in zip(r["log_toi"], logs["toi"]):
        assert abs(mine - theirs) < 1e-2
    # carried contact pattern (fetch_dyn) of the last assembly: tie-decided pairs aside, the reference's
    import ppf_b200
    ri = np.fromfile(os.path.join(refdir, "dyn_index.bin"), dtype=np.uint32)
    ro = np.fromfile(os.path.join(refdir, "dyn_offset.bin"), dtype=np.uint32)
    a = set(ppf_b200.pattern_keys(ri, ro).tolist()); b = set(ppf_b200.pattern_keys(r["dyn_index"], r["dyn_offset"]).tolist())
    assert len(a ^ b) <= 12 + 0.01 * len(a), (len(a), len(b), len(a ^ b))
    c = compare(os.path.join(refdir, "linsys_0.bin"), os.path.join(r["out_dir"], "linsys_0.bin"), verbose=False)
    nb = c["n_blocks"][0]
    assert c["only_ref"] + c["only_ours"] <= max(2, 0.002 * nb), c
    assert c["frac_blocks_off"] * nb <= 12 + 0.002 * nb, c      # a tie-decided contact touches <= 6 blocks
    # ... and <= 4 rows.  A system in equilibrium (balloon_n24 at step 20: pressure against membrane stress, |rhs| ~ 1e-3 of the
    # terms that cancel in it) carries fp32 cancellation noise of ~0.5 % of its tiny rhs in every row; the rhs bar is relative to
    # max |rhs|, so it is waived when the step itself lands within 1e-6 of the reference (500x tighter than POS_TOL)
    if float(np.abs(r["curr"] - ref).max()) >= 1e-6:
        assert c["frac_rhs_rows_off"] * c["nrow"][0] <= 8 + 0.005 * c["nrow"][0], c


def test_oracle_pair_lists_are_sorted_unique_and_geometric(run_job):
    """Pair lists of the oracle: sorted, unique, every accepted pair really is within the gap (float64 recheck)."""
    name, scene, n, state, refdir = golden_cases()[1]
    r = run_job({"lib": "oracle", "scene": scene, "n": n, "state": state, "steps": 0, "detect": "state"})
    total = int(r["state_total"][0])
    assert total == sum(len(r["state_pairs%d" % k]) for k in range(4)) and total > 1000
    for k in range(4):
        p = r["state_pairs%d" % k].astype(np.int64)
        key = p[:, 0] << 32 | p[:, 1]
        assert (np.diff(key) > 0).all()
    pp = r["state_pairs2"].astype(np.int64)      # point-point: plain distance < ghat
    x = r["curr"].astype(np.float64)              # steps = 0: curr is the state itself
    d = np.linalg.norm(x[pp[:, 0]] - x[pp[:, 1]], axis=1)
    assert (pp[:, 1] < pp[:, 0]).all() and d.max() < 2.0 / n * 0.35 * 1.0001

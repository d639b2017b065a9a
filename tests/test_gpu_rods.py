"""Rod (yarn) scenes on a real B200: Hookean rod edges, interior-vertex bending, rod-rod contact with contact
offsets and friction (configs `woven` / `large-woven` of BASELINE.json).

Against the reference: tests/golden/woven_n32 (when present) is picked up by the parametrised golden tests of
test_gpu_parity.py.  Here the CUDA backend is compared with the CPU oracle on the same inputs, and checked
through size-independent properties at the full size of examples/woven.ipynb.

Bars: first assembled system -- block pattern exact up to tie-decided pairs (<= 0.2 % of the blocks), values
1e-3 relative; contact count of the first Newton iteration exact; PCG iterations 5 %; positions after three
chained steps: median 2e-5, worst vertex 2e-3 (scene size ~1; a tie-decided contact moves a vertex by ~1e-4).
"""
import json
import os
import sys

import numpy as np
import pytest

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "scripts"))
from compare_linsys import compare  # noqa: E402

pytestmark = pytest.mark.gpu


def test_woven_first_steps_match_the_oracle(run_job):
    job = {"scene": "woven", "n": 32, "state": None, "steps": 3, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"))
    o = run_job(dict(job, lib="oracle"))
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    stats = json.loads(str(g["stats_json"]))
    assert stats["n_kernel_launches"] > 0 and stats["n_cg_iter"] > 0
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    nb = c["n_blocks"][0]
    assert c["only_ref"] + c["only_ours"] <= max(2, 0.002 * nb), c
    assert c["frac_blocks_off"] * nb <= 12 + 0.002 * nb, c
    assert c["frac_rhs_rows_off"] * c["nrow"][0] <= 8 + 0.005 * c["nrow"][0], c
    assert g["log_num_contact"][0] == o["log_num_contact"][0] and g["log_num_contact"][0] > 100
    assert len(g["log_iter"]) == len(o["log_iter"])
    for mine, theirs in zip(g["log_iter"], o["log_iter"]):
        assert abs(mine - theirs) <= 2 + 0.05 * theirs
    for mine, theirs in zip(g["log_toi"], o["log_toi"]):
        assert abs(mine - theirs) < 1e-2
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))


def test_woven_contact_pairs_match_the_oracle(run_job):
    """rod-rod contacts at the rest pose (every ring passes every strand within the contact gap): pair sets are
    exact outside the fp32 ties the oracle itself reports"""
    g = run_job({"lib": "ours", "scene": "woven", "n": 48, "state": None, "steps": 0, "detect": "state", "detect_twice": True})
    o = run_job({"lib": "oracle", "scene": "woven", "n": 48, "state": None, "steps": 0, "detect": "state"})
    assert bool(g["state_repeat_equal"][0])
    n_pairs = n_tie = 0
    for k in range(4):
        key = lambda p: (p[:, 0].astype(np.int64) << 32 | p[:, 1].astype(np.int64))
        gk, ok, fk = key(g["state_pairs%d" % k]), key(o["state_pairs%d" % k]), key(o["state_fragile%d" % k])
        assert (np.diff(gk) > 0).all()
        missing = np.setdiff1d(np.setdiff1d(ok, fk), gk)
        extra = np.setdiff1d(np.setdiff1d(gk, ok), fk)
        assert missing.size == 0 and extra.size == 0, (k, missing[:4], extra[:4])
        n_pairs += ok.size; n_tie += fk.size
    assert n_pairs > 500 and n_tie <= 5 + 5e-3 * n_pairs
    assert g["state_pairs0"].size == 0                              # no faces: no point-triangle pairs


def test_woven_at_full_size_keeps_its_invariants(run_job):
    """examples/woven.ipynb at its own size (115 588 vertices, 115 440 rods, ~1.2e5 contacts from the first step):
    too big for the oracle in seconds, so check size-independent properties."""
    r = run_job({"lib": "ours", "scene": "woven", "n": 0, "state": None, "steps": 3, "detect": "final", "detect_twice": True, "fix_positions": True})
    assert bool(r["ok"].all()) and np.isfinite(r["curr"]).all()
    assert int(r["counts"][0]) == 115588 and int(r["counts"][2]) == 115440 and int(r["counts"][1]) == 0
    assert bool(r["final_repeat_equal"][0])
    assert r["log_num_contact"].min() > 50000
    assert (r["log_toi"] > 0).all() and (r["log_toi"] <= 1).all()
    assert (r["log_reresid"] < 1e-3 * 1.0001).all()
    if (r["log_toi"][-2:] == 1.0).all():                            # last step not cut: exact Dirichlet rows
        idx = r["fix_index"].astype(np.int64)
        assert np.array_equal(r["curr"][idx], r["fix_target"])
    off = r["dyn_offset"].astype(np.int64)
    assert off[0] == 0 and (np.diff(off) >= 0).all() and off[-1] == r["dyn_index"].size and r["dyn_index"].size > 10000
    for k in range(4):
        p = r["final_pairs%d" % k].astype(np.int64)
        if p.size:
            assert (np.diff(p[:, 0] << 32 | p[:, 1]) > 0).all()
    # no yarn segment collapsed or tore: lengths stay within the pre-tension range (rest = 0.8 x drawn)
    import ppf_b200
    h = ppf_b200.host()
    import ctypes
    p = ctypes.c_void_p(h.ppfh_create(b"woven", ctypes.c_uint(0)))
    edge = np.empty((int(r["counts"][2]), 2), np.uint32)
    h.ppfh_topology(p, None, edge.ctypes.data_as(ctypes.c_void_p), None)
    h.ppfh_destroy(p)
    x = r["curr"].astype(np.float64)
    ln = np.linalg.norm(x[edge[:, 0]] - x[edge[:, 1]], axis=1)
    assert ln.min() > 3e-3 and ln.max() < 2.5e-2


def test_rod_strain_limit_matches_the_oracle(run_job):
    """woven-sl: the woven cylinder stretched along its axis, rods with a 2 % strain limit (strainlimiting.cu:122-245).
    The strain-limit line search cuts the very first Newton step; compare the cut and the result with the oracle."""
    job = {"scene": "woven-sl", "n": 32, "state": None, "steps": 3, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert len(g["log_SL_rod_toi"]) == len(o["log_SL_rod_toi"]) > 0 and o["log_SL_rod_toi"].min() < 1.0
    assert abs(g["log_SL_rod_toi"][0] - o["log_SL_rod_toi"][0]) < 2e-3      # first cut: same positions up to one PCG solve
    for mine, theirs in zip(g["log_SL_rod_toi"], o["log_SL_rod_toi"]):
        assert abs(mine - theirs) < 2e-2
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    nb = c["n_blocks"][0]
    assert c["only_ref"] + c["only_ours"] <= max(2, 0.002 * nb), c
    assert c["frac_blocks_off"] * nb <= 12 + 0.002 * nb, c
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))

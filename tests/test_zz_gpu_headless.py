"""examples/headless.py (BASELINE.json's first config names it) on a real B200, at res 24: a fully pinned sphere mesh pushed
through five pinned curtains -- moving prescribed object vs free cloth, strain limiting, min-newton-steps 8.

There is no reference golden for this scene: the starting state comes from the CPU oracle
(tests/golden/make_headless_state.py) and the CUDA backend is compared with the oracle on the step that follows.
Bars as in test_gpu_rods.py: first assembled system exact in pattern up to tie-decided pairs, values 1e-3; contact count
of the first Newton iteration exact; PCG iterations 5 %; positions after the step: median 2e-5, worst vertex 2e-3.
(The file sorts last on purpose: it widens coverage, it is not what the parity claims of DESIGN.md rest on.)
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
STATE = os.path.join(REPO, "tests", "golden", "headless_n24", "state_00068.bin")


def test_headless_step_matches_the_oracle(run_job):
    job = {"scene": "headless", "n": 24, "state": STATE, "steps": 1, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert json.loads(str(g["stats_json"]))["n_kernel_launches"] > 0
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    nb = c["n_blocks"][0]
    assert c["only_ref"] + c["only_ours"] <= max(2, 0.002 * nb), c
    assert c["frac_blocks_off"] * nb <= 12 + 0.002 * nb, c
    assert c["frac_rhs_rows_off"] * c["nrow"][0] <= 8 + 0.005 * c["nrow"][0], c
    assert abs(g["log_num_contact"][0] - o["log_num_contact"][0]) <= 1 and o["log_num_contact"][0] > 100
    assert abs(g["log_iter"][0] - o["log_iter"][0]) <= 2 + 0.05 * o["log_iter"][0]
    assert abs(g["log_toi"][0] - o["log_toi"][0]) < 1e-2
    assert abs(len(g["log_iter"]) - len(o["log_iter"])) <= 1 and len(g["log_iter"]) >= 9
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))
    # the prescribed sphere (last 162 vertices) is where the oracle has it: Dirichlet rows copy the pin targets (a step whose
    # accumulated line-search fraction stays below 1 rewinds them by that fraction of the 1.6 cm the sphere moves per step)
    assert float(np.abs(g["curr"][-162:] - o["curr"][-162:]).max()) < 1e-4


def test_kinematic_collider_lift_matches_the_oracle(run_job):
    """`lift`: a sheet pinned at two corners lifted by a keyframed invisible sphere (kinematic collider branch of
    contact.cu:1230-1263).  Chained steps through the first contacts, compared in aggregate like the drape test."""
    job = {"scene": "lift", "n": 24, "state": None, "steps": 30}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert o["log_num_contact"].max() > 30
    assert abs(len(g["log_num_contact"]) - len(o["log_num_contact"])) <= 4
    assert abs(g["log_num_contact"].max() - o["log_num_contact"].max()) <= 5 + 0.25 * o["log_num_contact"].max()
    assert abs(float(g["time"][0]) - float(o["time"][0])) < 0.1 * float(o["time"][0])
    x = g["curr"].astype(np.float64)
    t = min(float(g["time"][0]), 0.6) / 0.6
    c = np.array([0.1, -0.9 + 0.7 * t, 0.2])
    # the kinematic sphere is a soft penalty (stiffness m / ghat^2, no CCD: contact.cu:1230-1263), so millimetre penetration is
    # the reference's behaviour; the bar is the oracle's own minimum distance (0.49707 in round 1 on both), not the radius
    dmin_g = np.linalg.norm(x - c, axis=1).min()
    dmin_o = np.linalg.norm(o["curr"].astype(np.float64) - c, axis=1).min()
    assert np.isfinite(x).all() and abs(dmin_g - dmin_o) < 5e-4 and dmin_g > 0.5 - 5e-3, (dmin_g, dmin_o)
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 1e-3, float(np.median(d))


@pytest.mark.parametrize("scene", ["sheet-arap", "sheet-stvk", "sheet-snhk"])
def test_svd_based_shell_models_match_the_oracle(run_job, scene):
    """ARAP / StVK / SNHk shell faces (k_face_iso): first assembled system and five steps of a sheet hanging from two corners
    against the CPU oracle.  The derivative tables are bit-exact against the reference's headers and the eigen-projection is
    the one the strain limiter shares (tests/test_oracle_cpu.py); no reference golden exists for these models yet."""
    job = {"scene": scene, "n": 24, "state": None, "steps": 5, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    assert c["only_ref"] == 0 and c["only_ours"] == 0, c
    assert c["offdiag_max_rel_block_diff"] < 1e-3 and c["diag_max_rel_diff"] < 1e-3 and c["rhs_max_abs_diff_over_scale"] < 1e-3, c
    assert len(g["log_iter"]) == len(o["log_iter"])
    for mine, theirs in zip(g["log_iter"], o["log_iter"]):
        assert abs(mine - theirs) <= 2 + 0.05 * theirs
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))


@pytest.mark.parametrize("scene,n", [("sheet-damped", 24), ("sheet-arap-damped", 24), ("woven-damped", 32)])
def test_rayleigh_damping_matches_the_oracle(run_job, scene, n):
    """Rayleigh damping of membranes / hinges / rods (k_damp_*: the damped elements' delta on top of the undamped kernels)
    against the CPU oracle, which applies it in place like the reference (energy.cu:88-180).  The first assembled system
    already carries the damping: u = x - x^n is still zero there, but every damped element Hessian is scaled by 1 + beta/dt
    (= 3 for the membranes, 6 for the bending terms of these scenes), which the undamped system does not show."""
    job = {"scene": scene, "n": n, "state": None, "steps": 3, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    nb = c["n_blocks"][0]
    assert c["only_ref"] + c["only_ours"] <= max(2, 0.002 * nb), c
    assert c["frac_blocks_off"] * nb <= 12 + 0.002 * nb, c
    base = {"sheet-damped": "sheet", "sheet-arap-damped": "sheet-arap", "woven-damped": "woven"}[scene]
    u = run_job({"lib": "oracle", "scene": base, "n": n, "state": None, "steps": 1, "dump_linsys": True}, force_subprocess=True)
    cu = compare(os.path.join(u["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    assert cu["frac_blocks_off"] > 0.5, cu                      # ... and the damping is really on: the undamped matrix is a different one
    assert len(g["log_iter"]) == len(o["log_iter"])
    for mine, theirs in zip(g["log_iter"], o["log_iter"]):
        assert abs(mine - theirs) <= 2 + 0.05 * theirs
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))


def test_stitches_match_the_oracle(run_job):
    """`seam`: two sheets sewn by a row of stitches (k_stitch; barycentric and plain-vertex targets) against the CPU oracle"""
    job = {"scene": "seam", "n": 16, "state": None, "steps": 4, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    assert c["only_ref"] == 0 and c["only_ours"] == 0, c
    assert c["offdiag_max_rel_block_diff"] < 1e-3 and c["diag_max_rel_diff"] < 1e-3 and c["rhs_max_abs_diff_over_scale"] < 1e-3, c
    assert len(g["log_iter"]) == len(o["log_iter"])
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    # stiff, length-capped seam springs: the 1e-3 stopping residual of two different PCGs shows more than elsewhere
    # (the CPU emulation of this backend against the same oracle: median 3e-5, worst vertex 1.2e-3)
    assert float(np.median(d)) < 1e-4 and float(d.max()) < 5e-3, (float(np.median(d)), float(d.max()))


@pytest.mark.parametrize("scene,n", [("sheet-plastic", 24), ("woven-plastic", 32)])
def test_plasticity_matches_the_oracle(run_job, scene, n):
    """plastic creep of inverse rest matrices / rest angles after every step (k_plastic_*) against the CPU oracle: chained
    steps, so the crept rest shapes of the earlier steps shape the later ones"""
    steps = 6 if scene == "sheet-plastic" else 2     # chained woven steps separate by their line-search cuts after ~3 steps, plastic or not
    job = {"scene": scene, "n": n, "state": None, "steps": steps}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert abs(len(g["log_iter"]) - len(o["log_iter"])) <= 1
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 5e-5 and float(d.max()) < 5e-3, (float(np.median(d)), float(d.max()))
    base = {"sheet-plastic": "sheet", "woven-plastic": "woven"}[scene]
    u = run_job({"lib": "oracle", "scene": base, "n": n, "state": None, "steps": steps}, force_subprocess=True)
    if abs(float(u["time"][0]) - float(o["time"][0])) < 1e-9:           # same clock: the plastic run must be the closer one
        assert float(np.median(np.abs(g["curr"] - u["curr"]).max(axis=1))) > float(np.median(d))


_BARRIER = [("fast-twist", 16, os.path.join(REPO, "tests", "golden", "fast_twist_n16", "state_00060.bin")),
            ("woven-sl", 32, os.path.join(REPO, "tests", "golden", "woven_sl_n32", "state_00080.bin"))]


@pytest.mark.parametrize("kind", ["quad", "log"])
@pytest.mark.parametrize("scene,n,state", _BARRIER, ids=[c[0] for c in _BARRIER])
def test_quadratic_and_logarithmic_barriers_match_the_oracle(run_job, scene, n, state, kind):
    """ParamSet.barrier = quad / log (barrier/quadratic.hpp, barrier/logarithm.hpp; dispatch barrier/barrier.cu): one step
    from a contact-rich (twist) / strain-limited (woven-sl) saved state against the CPU oracle -- first assembled system,
    contact counts, positions (the device takes __logf where the host takes logf: inside the 2e-4 block bar)."""
    job = {"scene": "%s@%s" % (scene, kind), "n": n, "state": state, "steps": 1, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    nb = c["n_blocks"][0]
    assert c["only_ref"] + c["only_ours"] <= max(2, 0.002 * nb), c
    # __logf's absolute error (2^-21.4 near 1) shows in blocks made of nothing but a contact at the very edge of the gap
    assert c["frac_blocks_off"] * nb <= 12 + (0.01 if kind == "log" else 0.002) * nb, c
    assert abs(len(g["log_iter"]) - len(o["log_iter"])) <= 1
    assert abs(g["log_num_contact"][0] - o["log_num_contact"][0]) <= 2
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 1e-4 and float(d.max()) < 3e-3, (float(np.median(d)), float(d.max()))


def test_fix_xz_matches_the_oracle(run_job):
    """ParamSet.fix_xz (horizontal drag above a height; k_momentum + k_fix_xz) on the swinging sheet against the CPU oracle"""
    job = {"scene": "sheet-fixxz", "n": 24, "state": None, "steps": 8, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    assert c["only_ref"] == 0 and c["only_ours"] == 0, c
    assert c["offdiag_max_rel_block_diff"] < 1e-3 and c["diag_max_rel_diff"] < 1e-3 and c["rhs_max_abs_diff_over_scale"] < 1e-3, c
    assert len(g["log_iter"]) == len(o["log_iter"])
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))


def test_collision_windows_match_the_oracle(run_job):
    """collision windows (k_cw_refresh + the per-vertex flags in the broad phase and the intersection check): `headless` with
    the sphere's window closed, from the saved contact-rich state, against the CPU oracle; and the open window is a no-op"""
    job = {"scene": "headless-cw-closed", "n": 24, "state": STATE, "steps": 2}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert len(g["log_num_contact"]) == len(o["log_num_contact"])
    assert abs(g["log_num_contact"][0] - o["log_num_contact"][0]) <= 1
    for mine, theirs in zip(g["log_num_contact"], o["log_num_contact"]):
        assert abs(mine - theirs) <= 3 + 0.05 * theirs
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))
    p = run_job(dict(job, scene="headless", lib="ours"), force_subprocess=True)
    q = run_job(dict(job, scene="headless-cw-open", lib="ours"), force_subprocess=True)
    assert p["log_num_contact"][0] == q["log_num_contact"][0] > g["log_num_contact"][0] + 20
    assert float(np.abs(p["curr"] - q["curr"]).max()) < 2e-3      # atomics order only (same bar as against the oracle)


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


def test_static_collision_mesh_matches_the_oracle(run_job):
    """`drape-mesh`: contacts with the static collision mesh (k_cm_vertex / k_cm_edge / k_cm_face, their line-search kernels
    and the static part of the intersection check) against the CPU oracle: same contact counts and line-search cuts through
    the first touch, same positions"""
    job = {"scene": "drape-mesh", "n": 24, "state": None, "steps": 7, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    assert abs(len(g["log_num_contact"]) - len(o["log_num_contact"])) <= 2 and o["log_num_contact"].max() > 15
    first_g, first_o = int(np.argmax(g["log_num_contact"] > 0)), int(np.argmax(o["log_num_contact"] > 0))
    assert abs(first_g - first_o) <= 1 and first_o > 0
    assert abs(g["log_num_contact"][-1] - o["log_num_contact"][-1]) <= 3 + 0.25 * o["log_num_contact"][-1]
    assert abs(float(g["time"][0]) - float(o["time"][0])) < 0.1 * float(o["time"][0])
    x = g["curr"].astype(np.float64)
    assert np.isfinite(x).all() and np.linalg.norm(x, axis=1).min() > 0.37 and x[:, 1].min() > -0.42
    if len(g["log_num_contact"]) == len(o["log_num_contact"]) and abs(float(g["time"][0]) - float(o["time"][0])) < 1e-5:
        d = np.abs(g["curr"] - o["curr"]).max(axis=1)
        assert float(np.median(d)) < 1e-4 and float(d.max()) < 3e-3, (float(np.median(d)), float(d.max()))


def test_torque_pins_match_the_oracle(run_job):
    """`sheet-torque`: a torque pin spins up a free sheet (k_torque_groups / k_torque_vertices) -- first system, positions and
    the resulting angular momentum against the CPU oracle"""
    job = {"scene": "sheet-torque", "n": 24, "state": None, "steps": 5, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    assert c["only_ref"] == 0 and c["only_ours"] == 0, c
    # rhs: 5e-3.  The group's axis comes from a Cardano eigen-solve (energy.cu:20-86) that is ill-conditioned for a strip (two
    # close small eigenvalues: acos near its branch point), evaluated with the fast __cosf on the device like the reference's
    # float_math.hpp and with cosf on the host; round 2 measured 1.6e-3 on a B200.  The tight bar for this feature is the
    # reference's own golden (tests/golden/sheet_torque_n24 in test_gpu_parity.py), where both sides run on the device.
    assert c["offdiag_max_rel_block_diff"] < 1e-3 and c["diag_max_rel_diff"] < 1e-3 and c["rhs_max_abs_diff_over_scale"] < 5e-3, c
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))
    x, v = g["curr"].astype(np.float64), (g["curr"] - g["prev"]).astype(np.float64)
    L = np.cross(x - x.mean(0), v).sum(0)
    assert abs(L[2]) > 0.9 * np.linalg.norm(L) > 0                 # spun about the strip's long axis (z)


def test_inflation_pressure_matches_the_oracle(run_job):
    """`balloon`: FaceParam.pressure (k_face_pressure: conditioned gradient + SVD-projected Hessian) against the CPU oracle"""
    job = {"scene": "balloon", "n": 24, "state": None, "steps": 6, "dump_linsys": True}
    g = run_job(dict(job, lib="ours"), force_subprocess=True)
    o = run_job(dict(job, lib="oracle"), force_subprocess=True)
    assert bool(g["ok"].all()) and bool(o["ok"].all())
    c = compare(os.path.join(o["out_dir"], "linsys_0.bin"), os.path.join(g["out_dir"], "linsys_0.bin"), verbose=False)
    assert c["only_ref"] == 0 and c["only_ours"] == 0, c
    assert c["offdiag_max_rel_block_diff"] < 1e-3 and c["diag_max_rel_diff"] < 1e-3 and c["rhs_max_abs_diff_over_scale"] < 1e-3, c
    assert len(g["log_iter"]) == len(o["log_iter"])
    d = np.abs(g["curr"] - o["curr"]).max(axis=1)
    assert float(np.median(d)) < 2e-5 and float(d.max()) < 2e-3, (float(np.median(d)), float(d.max()))

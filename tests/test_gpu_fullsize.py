"""The BASELINE.json configurations AT THEIR OWN SIZE against the live reference on a real B200.

For twist (n = 360, V = 203 965), five-twist (n = 180, V = 107 695), hang (res 128, V = 16 384), woven (256 strands) and
large-twist (n = 1410, V = 3 214 258) this backend first runs some steps from t = 0 and saves the state; then the UNMODIFIED
reference build (oracle/_ref, compiled from /root/reference by oracle/Makefile.ref; it travels to the GPU box) and this backend
both load that state and take ONE step.  The reference library used here carries one extra hook (oracle/ref_pair_dump.hpp) that
writes its four detect-once contact-pair caches (contact.cu:364-377) of the first assembly to a file.

Bars (BASELINE.json north_star: "identical sorted contact-pair sets (bit-exact) and per-frame vertex positions within a stated
fp32 tolerance"):
  contact-pair sets    the four sorted (query, candidate) lists of the first assembly: bit-exact, every pair
  contact count        first Newton iteration: exact
  dynamic pattern      fetch_dyn after the step: a pair decided at positions that carry fp32-atomics noise (second Newton
                       iteration on) may flip in either library from run to run -> symmetric difference <= 4 + 0.4 %
  PCG iterations       first solve: 2 %
  positions            after the step: 90 % of the vertices within 5e-5, worst vertex 5e-4 (scene size ~1)
"""
import os

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
REF_PAIRS = os.path.join(REPO, "oracle", "_ref", "libsimbackend_cuda_pairs.so")

#        scene          warm-up steps of this backend before the compared step
CONFIGS = [("twist", 75), ("five-twist", 60), ("hang", 40), ("woven", 30), ("large-twist", 1)]


def _describe_difference(ppf, scene, state, kind, a, b):
    """the pairs only one side has, with the bit patterns of the positions of their vertices (enough to replay the predicate)"""
    key = lambda p: p[:, 0].astype(np.int64) << 32 | p[:, 1].astype(np.int64)
    ka, kb = key(a), key(b)
    x = ppf.load_state_file(state)["curr"]
    sess = ppf.Session(scene, n=0, lib=ppf.LIB_ORACLE) if os.path.exists(ppf.LIB_ORACLE) else None
    face, edge, _ = sess.topology() if sess else (None, None, None)
    out = []
    for name, only in (("reference only", np.setdiff1d(ka, kb)), ("this backend only", np.setdiff1d(kb, ka))):
        for k in only[:6]:
            q, c = int(k >> 32), int(k & 0xFFFFFFFF)
            verts = []
            if edge is not None:
                verts = ([q] if kind < 3 else list(edge[q])) + (list(face[c]) if kind == 0 else list(edge[c]) if kind in (1, 3) else [c])
            out.append("%s: (%d, %d) vertices %s positions %s" % (name, q, c, [int(v) for v in verts],
                                                                 [x[int(v)].view(np.uint32).tolist() for v in verts]))
    return "\n".join(out)


@pytest.mark.parametrize("scene,warm", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_full_size_step_matches_the_live_reference(run_job, tmp_path, scene, warm):
    if not os.path.exists(REF_PAIRS):
        pytest.skip("oracle/_ref/libsimbackend_cuda_pairs.so is not built (needs /root/reference at build time)")
    state = str(tmp_path / "state.bin")
    w = run_job({"lib": "ours", "scene": scene, "n": 0, "state": None, "steps": warm, "save_state": state}, force_subprocess=True)
    assert bool(w["ok"].all()) and os.path.exists(state)
    job = {"scene": scene, "n": 0, "state": state, "steps": 1}
    r = run_job(dict(job, lib="reference_pairs", dump_pairs=True), force_subprocess=True)
    g = run_job(dict(job, lib="ours", detect="state"), force_subprocess=True)
    assert bool(r["ok"].all()) and bool(g["ok"].all())
    import ppf_b200
    # --- contact-pair sets of the first assembly: bit-exact ---
    n_pairs = 0
    for k, kind in enumerate(("point-face", "point-edge", "point-point", "edge-edge")):
        a, b = r["ref_pairs%d" % k], g["state_pairs%d" % k]
        if not (a.shape == b.shape and np.array_equal(a, b)):
            pytest.fail("%s lists differ: %d (reference) vs %d pairs\n%s" % (kind, len(a), len(b), _describe_difference(ppf_b200, scene, state, k, a, b)))
        n_pairs += len(a)
    assert n_pairs == int(r["log_num_contact"][0]) == int(g["log_num_contact"][0])
    if scene != "hang":
        assert n_pairs > 1000                                        # the compared state is in contact
    # --- solve and step ---
    assert abs(g["log_iter"][0] - r["log_iter"][0]) <= max(4.0, 1 + 0.02 * r["log_iter"][0])
    assert len(g["log_num_contact"]) == len(r["log_num_contact"])     # same number of Newton iterations
    for mine, theirs in zip(g["log_toi"], r["log_toi"]):
        assert abs(mine - theirs) < 1e-2
    d = np.abs(g["curr"] - r["curr"]).max(axis=1)
    assert float(np.percentile(d, 90)) < 5e-5 and float(d.max()) < 5e-4, (float(np.percentile(d, 90)), float(d.max()))
    pa = set(ppf_b200.pattern_keys(r["dyn_index"], r["dyn_offset"]).tolist())
    pb = set(ppf_b200.pattern_keys(g["dyn_index"], g["dyn_offset"]).tolist())
    assert len(pa ^ pb) <= 4 + 0.004 * len(pa), (len(pa), len(pb), len(pa ^ pb))

"""Host-side logic (C++ host, reference: crates/ppf-cts-solver/src/{builder,scene,backend}.rs):
scene generation, dataset tables, state round trip, stepping a backend through the ABI on CPU
(the oracle library stands in for a backend here; the CUDA backend is exercised by the gpu tests)."""
import os

import numpy as np

from conftest import REPO, golden_cases


def test_scene_counts_follow_euler(ppf):
    for scene, n in (("fast-twist", 16), ("fast-twist", 24), ("sheet", 8)):
        s = ppf.Session(scene, n=n, lib=ppf.LIB_ORACLE) if scene == "sheet" and False else None
    # counts through the builder only (no backend needed): V - E + F = chi (0 for an open cylinder, 1 for a sheet)
    import ctypes

    h = ppf.host()
    for scene, n, chi in (("fast-twist", 16, 0), ("twist", 40, 0), ("sheet", 8, 1)):
        p = ctypes.c_void_p(h.ppfh_create(scene.encode(), ctypes.c_uint(n)))
        c = (ctypes.c_uint * 6)()
        h.ppfh_counts(p, c)
        v, f, e, hg = int(c[0]), int(c[1]), int(c[2]), int(c[3])
        assert v - e + f == chi
        boundary = 3 * f - 2 * hg  # every interior edge is a hinge
        assert e == hg + boundary
        h.ppfh_destroy(p)


def test_topology_tables_are_consistent(ppf):
    import ctypes

    h = ppf.host()
    p = ctypes.c_void_p(h.ppfh_create(b"fast-twist", ctypes.c_uint(16)))
    c = (ctypes.c_uint * 6)()
    h.ppfh_counts(p, c)
    nv, nf, ne = int(c[0]), int(c[1]), int(c[2])
    face = np.empty((nf, 3), np.uint32); edge = np.empty((ne, 2), np.uint32); vp = np.empty((nv, 2), np.float32)
    h.ppfh_topology(p, face.ctypes.data_as(ctypes.c_void_p), edge.ctypes.data_as(ctypes.c_void_p), vp.ctypes.data_as(ctypes.c_void_p))
    assert face.max() == nv - 1 and edge.max() == nv - 1
    # every face edge is in the edge list exactly once (undirected)
    fe = np.concatenate([face[:, [0, 1]], face[:, [1, 2]], face[:, [2, 0]]])
    key = lambda a: np.minimum(a[:, 0], a[:, 1]).astype(np.int64) << 32 | np.maximum(a[:, 0], a[:, 1])
    assert np.array_equal(np.unique(key(fe)), np.sort(key(edge)))
    assert len(np.unique(key(edge))) == ne
    assert (vp[:, 0] > 0).all()  # contact gap resolved for every vertex
    h.ppfh_destroy(p)


def test_state_round_trip_and_stepping_through_the_abi(ppf, run_job, tmp_path):
    name, scene, n, state, refdir = golden_cases()[0]
    st = ppf.load_state_file(state)
    r = run_job({"lib": "oracle", "scene": scene, "n": n, "state": state, "steps": 0})
    assert np.array_equal(r["curr"], st["curr"]) and np.array_equal(r["prev"], st["prev"])  # upload -> fetch is lossless
    r2 = run_job({"lib": "oracle", "scene": scene, "n": n, "state": state, "steps": 2}, force_subprocess=True)
    assert bool(r2["ok"].all()) and float(r2["time"][0]) > st["time"]
    assert abs((float(r2["time"][0]) - st["time"]) - 2 * 0.016665) < 1e-4  # dt = min(dt, 0.9999 / fps)
    assert np.isfinite(r2["curr"]).all()


def _rods(ppf, scene, n):
    """built rod scene through the C API (no backend): counts, edges, rod props, vertex props, tables"""
    import ctypes

    h = ppf.host()
    h.ppfh_rods.restype = ctypes.c_uint
    p = ctypes.c_void_p(h.ppfh_create(scene.encode(), ctypes.c_uint(n)))
    c = (ctypes.c_uint * 6)()
    h.ppfh_counts(p, c)
    nv, nf, ne, nh, nfix, nnz = [int(v) for v in c]
    vp = ctypes.c_void_p
    n_rod = int(h.ppfh_rods(p, None, None, None, None, None))
    rod = np.empty((n_rod, 5), np.float32); vert = np.empty((nv, 2), np.float32); slots = np.empty((nv, 9), np.uint32)
    off = np.empty(nv + 1, np.uint32); col = np.empty(nnz, np.uint32)
    h.ppfh_rods(p, rod.ctypes.data_as(vp), vert.ctypes.data_as(vp), slots.ctypes.data_as(vp), off.ctypes.data_as(vp), col.ctypes.data_as(vp))
    edge = np.empty((ne, 2), np.uint32); par = np.empty((nv, 2), np.float32)
    h.ppfh_topology(p, None, edge.ctypes.data_as(vp), par.ctypes.data_as(vp))
    h.ppfh_destroy(p)
    return dict(nv=nv, nf=nf, ne=ne, nh=nh, nfix=nfix, nnz=nnz, n_rod=n_rod, rod=rod, vert=vert, slots=slots, off=off, col=col, edge=edge, vparam=par)


def test_woven_scene_sizes_match_the_reference_examples(ppf):
    """examples/woven.ipynb: 115.6 K vertices / 115.4 K rods (README.md:599); strands 148 x 256, rings 175 x 444"""
    r = _rods(ppf, "woven", 0)
    assert (r["nv"], r["n_rod"], r["ne"], r["nf"], r["nh"]) == (148 * 256 + 175 * 444, 148 * 255 + 175 * 444, 148 * 255 + 175 * 444, 0, 0)
    assert r["nfix"] == 2 * 148                                   # both ends of every axial strand
    # diagonal + one block per rod edge + one (j, k) block per interior vertex (every vertex but the 2 x 148 strand ends)
    assert r["nnz"] == r["nv"] + r["n_rod"] + (r["nv"] - 2 * 148)


def test_rod_tables_are_consistent(ppf):
    r = _rods(ppf, "woven", 32)
    nv, n_rod, edge, rod, vert, slots, off, col = r["nv"], r["n_rod"], r["edge"], r["rod"], r["vert"], r["slots"], r["off"], r["col"]
    assert n_rod == r["ne"] and n_rod > 0
    # rest length = drawn length x length-factor 0.8; mass = density x drawn length; mass lumped half / half
    assert np.allclose(rod[:, 0], 0.8 * rod[:, 1], rtol=1e-6) and np.allclose(rod[:, 2], rod[:, 1], rtol=1e-6)
    m = np.zeros(nv); np.add.at(m, edge[:, 0], 0.5 * rod[:, 2]); np.add.at(m, edge[:, 1], 0.5 * rod[:, 2])
    assert np.allclose(vert[:, 0], m, rtol=1e-5)
    assert (rod[:, 3] == np.float32(1e5)).all()
    # bend = 1e-3 (seg / 1 cm)^2 with seg = mean drawn spacing x 0.8, per strand
    assert np.allclose(rod[:, 4], 1e-3 * (rod[:, 0] / 1e-2) ** 2, rtol=2e-2)
    assert (vert[:, 1] == np.float32(np.pi)).all()                # straight rest shape unless bend-rest-from-geometry
    # length-weighted average of identical rod parameters: equal up to the rounding of (w0 g + w1 g) / (w0 + w1)
    assert np.allclose(r["vparam"][:, 0], 1.5e-3, rtol=1e-6) and np.allclose(r["vparam"][:, 1], 2e-3, rtol=1e-6)
    # fixed index table: sorted rows starting at the diagonal, upper triangle only
    deg = np.zeros(nv, np.int64); np.add.at(deg, edge[:, 0], 1); np.add.at(deg, edge[:, 1], 1)
    for i in range(nv):
        row = col[off[i]:off[i + 1]]
        assert row[0] == i and (np.diff(row.astype(np.int64)) > 0).all()
    # bending stencil (j, i, k) of every interior vertex: slot(a, b) addresses column max(a, b)'s entry of row min(a, b)
    nbr = [[] for _ in range(nv)]
    for e0, e1 in edge:
        nbr[e0].append(int(e1)); nbr[e1].append(int(e0))
    n_interior = 0
    for i in range(nv):
        if deg[i] != 2:
            assert (slots[i] == 0xFFFFFFFF).all()
            continue
        n_interior += 1
        st = [nbr[i][0], i, nbr[i][1]]
        for a in range(3):
            for b in range(3):
                s = slots[i, 3 * a + b]
                if st[a] > st[b]:
                    assert s == 0xFFFFFFFF
                else:
                    assert off[st[a]] <= s < off[st[a] + 1] and col[s] == st[b]
    assert n_interior == nv - r["nfix"]


def test_oracle_steps_a_woven_rod_scene(run_job):
    """CPU restatement on a small woven cylinder: contacts from the first step, rods keep their length, pins are exact"""
    o = run_job({"lib": "oracle", "scene": "woven", "n": 32, "state": None, "steps": 6, "fix_positions": True})
    assert bool(o["ok"].all()) and np.isfinite(o["curr"]).all()
    assert o["log_num_contact"].min() > 100 and (o["log_toi"] > 0).all() and (o["log_toi"] <= 1).all()
    idx = o["fix_index"].astype(np.int64)
    # a step that the line search cuts short rewinds its kinematic pins by (1 - toi_advanced) of their per-step motion
    # (main.cu:1418-1440), so the pins sit between the previous and the current targets
    assert idx.size == 36 and float(np.abs(o["curr"][idx] - o["fix_target"]).max()) < 0.03
    first = run_job({"lib": "oracle", "scene": "woven", "n": 32, "state": None, "steps": 1, "fix_positions": True})
    assert (first["log_toi"] == 1.0).all() and np.array_equal(first["curr"][idx], first["fix_target"])   # full step: exact Dirichlet rows


def test_headless_scene_matches_the_reference_example(ppf):
    """examples/headless.py: five 64 x 64 curtains + a fully pinned icosphere(r = 0.5, subdiv 4): 23 042 vertices"""
    import ctypes

    h = ppf.host()
    p = ctypes.c_void_p(h.ppfh_create(b"headless", ctypes.c_uint(0)))
    c = (ctypes.c_uint * 6)()
    h.ppfh_counts(p, c)
    nv, nf, ne, nh, nfix = [int(v) for v in c[:5]]
    assert nv == 5 * 64 * 64 + 2562 and nf == 5 * 2 * 63 * 63 + 5120
    assert nfix == 5 * 64 + 2562                      # top edge of every curtain + the whole sphere
    face = np.empty((nf, 3), np.uint32)
    h.ppfh_topology(p, face.ctypes.data_as(ctypes.c_void_p), None, None)
    h.ppfh_destroy(p)
    # the sphere is the last object: a closed surface (V - E + F = 2 with E = 3F / 2)
    sph = face[5 * 2 * 63 * 63:]
    vs = np.unique(sph)
    assert vs.size == 2562 and vs.size - 3 * len(sph) // 2 + len(sph) == 2


def test_oracle_steps_headless_from_a_contact_state(run_job):
    """the pinned sphere pushing into the curtains (state made by tests/golden/make_headless_state.py): moving fully pinned
    object vs free cloth, strain limiting and min-newton-steps 8 in one step"""
    state = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "headless_n24", "state_00068.bin")
    o = run_job({"lib": "oracle", "scene": "headless", "n": 24, "state": state, "steps": 1, "fix_positions": True})
    assert bool(o["ok"].all()) and len(o["log_iter"]) >= 9          # min-newton-steps 8 + the error-reduction iteration
    assert o["log_num_contact"].min() > 100
    idx = o["fix_index"].astype(np.int64)
    assert idx.size == 5 * 24 + 162
    # the sphere (last 162 vertices) stays rigid: it is fully prescribed
    x = o["curr"].astype(np.float64)[-162:]
    r = np.linalg.norm(x - x.mean(axis=0), axis=1)
    assert np.allclose(r, 0.5, atol=1e-5)


def test_example_scene_sizes(ppf):
    """further reference examples the host restates; vertex counts at the notebooks' own resolutions"""
    import ctypes

    h = ppf.host()
    want = {"curtain": 15 * 64 * 64 + 2562, "drape-sheets": 5 * 128 * 128 + 2562, "yarn": 53 * 520 + 2 * 25 * 20, "belt": None}
    for scene, nv_want in want.items():
        p = ctypes.c_void_p(h.ppfh_create(scene.encode(), ctypes.c_uint(0)))
        assert p
        c = (ctypes.c_uint * 6)()
        h.ppfh_counts(p, c)
        if nv_want is not None:
            assert int(c[0]) == nv_want, (scene, int(c[0]))
        assert int(c[4]) > 0                                  # every one of them has pins
        h.ppfh_destroy(p)


def test_oracle_steps_reduced_example_scenes(run_job):
    """reduced-resolution variants step on the CPU oracle: sheets settling on a fully pinned sphere mesh (contacts between free
    cloth and prescribed faces), and the five-cylinder twist with orbiting pins"""
    o = run_job({"lib": "oracle", "scene": "drape-sheets", "n": 24, "state": None, "steps": 25})
    assert bool(o["ok"].all()) and o["log_num_contact"].max() > 50
    o = run_job({"lib": "oracle", "scene": "large-five-twist", "n": 24, "state": None, "steps": 8})
    assert bool(o["ok"].all()) and o["log_num_contact"].max() > 500


def test_oracle_steps_sheets_with_the_svd_based_shell_models(run_job):
    """ARAP / StVK / SNHk shells (energy.cu:337-352) on a sheet hanging from two corners: each model steps and sags differently"""
    res = {}
    for m in ("sheet", "sheet-arap", "sheet-stvk", "sheet-snhk"):
        o = run_job({"lib": "oracle", "scene": m, "n": 16, "state": None, "steps": 20})
        assert bool(o["ok"].all()) and np.isfinite(o["curr"]).all()
        res[m] = o["curr"]
        assert o["curr"][:, 1].min() < -0.05                       # gravity pulled the free part down
    for a in res:
        for b in res:
            if a < b:
                assert float(np.abs(res[a] - res[b]).max()) > 1e-5  # four different constitutive models


def test_oracle_rayleigh_damping_dissipates(run_job):
    """deformation / bending damping (energy.cu:88-180) on a sheet hanging from two corners and on woven rods: the damped run
    steps, differs from the undamped one, and carries less internal (non-rigid) motion"""
    a = run_job({"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 40})
    d = run_job({"lib": "oracle", "scene": "sheet-damped", "n": 16, "state": None, "steps": 40})
    assert bool(a["ok"].all()) and bool(d["ok"].all())
    assert float(np.abs(a["curr"] - d["curr"]).max()) > 1e-4

    def wobble(r):            # velocity field minus its mean: what membrane / bending damping acts on
        v = (r["curr"] - r["prev"]).astype(np.float64)
        return float(np.linalg.norm(v - v.mean(axis=0)))
    assert wobble(d) < wobble(a)
    w = run_job({"lib": "oracle", "scene": "woven-damped", "n": 32, "state": None, "steps": 6})
    assert bool(w["ok"].all()) and w["log_num_contact"].min() > 100


def test_oracle_sews_two_sheets_with_stitches(run_job):
    """`seam`: a sheet hanging from two corners carries a second sheet through a row of stitches (energy.cu:796-886), every
    other stitch landing on an edge midpoint of the second sheet (barycentric target)"""
    o = run_job({"lib": "oracle", "scene": "seam", "n": 12, "state": None, "steps": 50})
    assert bool(o["ok"].all()) and np.isfinite(o["curr"]).all()
    x = o["curr"].astype(np.float64)
    A, B = x[:144], x[144:]
    near = np.sort(np.linalg.norm(A[:, None, :] - B[None, :, :], axis=2).min(axis=0))[:12]
    assert near.max() < 0.05                      # the twelve seam vertices of the second sheet stay sewn to the first
    assert B[:, 1].min() < A[:, 1].min() - 0.2    # and it hangs below


def _pins_at(ppf, scene, n, time, prev_time):
    import ctypes

    h = ppf.host()
    h.ppfh_constraint_at.restype = ctypes.c_uint
    p = ctypes.c_void_p(h.ppfh_create(scene.encode(), ctypes.c_uint(n)))
    c = (ctypes.c_uint * 6)()
    h.ppfh_counts(p, c)
    nv = int(c[0])
    x0 = np.empty((nv, 3), np.float32)
    h.ppfh_initial_positions(p, x0.ctypes.data_as(ctypes.c_void_p))
    vp = ctypes.c_void_p
    k = int(h.ppfh_constraint_at(p, ctypes.c_double(time), ctypes.c_double(prev_time), None, None, None, None))
    idx = np.empty(k, np.uint32); pos = np.empty((k, 3), np.float32); dlt = np.empty((k, 3), np.float32); kin = np.empty(k, np.uint8)
    h.ppfh_constraint_at(p, ctypes.c_double(time), ctypes.c_double(prev_time), idx.ctypes.data_as(vp), pos.ctypes.data_as(vp), dlt.ctypes.data_as(vp), kin.ctypes.data_as(vp))
    h.ppfh_destroy(p)
    return x0.astype(np.float64), idx.astype(np.int64), pos.astype(np.float64), dlt.astype(np.float64), kin


def test_twist_pin_operations_follow_the_notebook(ppf):
    """examples/twist.ipynb: each end ring spins about +-x at 360 deg/s, scales to 0.9 over 10 s (about the origin) and moves
    +-0.2 along x over 10 s, applied in that order (scene.rs:2747-2870, pin_apply.rs:27-160); FixPair.step_delta is the
    displacement since the previous step's time; a pin that moves is kinematic"""
    for t, tp in ((0.25, 0.25 - 1 / 60), (3.0, 3.0 - 1 / 60), (12.0, 12.0 - 1 / 60)):
        x0, idx, pos, dlt, kin = _pins_at(ppf, "twist", 24, t, tp)
        assert idx.size > 0 and kin.all()

        def where(tt):
            out = np.empty((idx.size, 3))
            for k, v in enumerate(idx):
                p = x0[v]
                sg = -1.0 if p[0] < 0 else 1.0
                ang = np.deg2rad(360.0) * tt * (-sg)                  # left ring: axis +x, right ring: axis -x
                c, s = np.cos(ang), np.sin(ang)
                q = np.array([p[0], c * p[1] - s * p[2], s * p[1] + c * p[2]])       # rotation about the x axis
                q = q * (1.0 + (0.9 - 1.0) * min(tt, 10.0) / 10.0)                   # scale about the origin
                q = q + np.array([-sg * 0.2 * min(tt, 10.0) / 10.0, 0, 0])           # move_by
                out[k] = q
            return out
        want, want_prev = where(t), where(tp)
        assert np.abs(pos - want).max() < 2e-6, np.abs(pos - want).max()
        assert np.abs(dlt - (want - want_prev)).max() < 2e-6
    # at t = 0 nothing has moved yet: targets are the rest positions, and the ring is static for the solver
    x0, idx, pos, dlt, kin = _pins_at(ppf, "twist", 24, 0.0, 0.0)
    assert np.abs(pos - x0[idx]).max() < 1e-7 and not dlt.any()


def test_static_pins_are_not_kinematic(ppf):
    x0, idx, pos, dlt, kin = _pins_at(ppf, "hang", 16, 0.7, 0.699)
    assert idx.size > 0 and not kin.any() and not dlt.any() and np.abs(pos - x0[idx]).max() < 1e-7


def test_five_twist_pin_operations_follow_the_notebook(ppf):
    """examples/five-twist.ipynb: spin (5 s) -> move_by (25 s) -> scale about the ring's own centre (5 s) -> orbit about the common
    axis (5..15 s), all in the object's local frame, then the object's displacement"""
    half = 2.6
    for t in (2.0, 7.0, 20.0):
        x0, idx, pos, dlt, kin = _pins_at(ppf, "five-twist", 16, t, t - 1 / 60)
        assert idx.size > 0 and kin.all()
        # the object of a pinned vertex: recover its displacement from the cylinder axis (rings sit at radius 1 about x)
        worst = 0.0
        for k, v in enumerate(idx):
            g = x0[v]                                                # rest position incl. the object's displacement
            cands = [np.array([0.0, np.cos(2 * i / 5 * np.pi), np.sin(2 * i / 5 * np.pi)]) for i in range(5)]
            d = min(cands, key=lambda c: abs(np.linalg.norm(g[1:] - c[1:]) - 0.55))
            p = g - d                                                # local rest position
            sg = -1.0 if p[0] < 0 else 1.0

            def rot_x(q, deg, sign, centre):
                a = np.deg2rad(deg) * sign
                c, s = np.cos(a), np.sin(a)
                r = q - centre
                return np.array([r[0], c * r[1] - s * r[2], s * r[1] + c * r[2]]) + centre
            q = rot_x(p, 360.0 * min(t, 5.0), -sg, np.zeros(3))
            q = q + np.array([-sg * 1.0 * min(t, 25.0) / 25.0, 0, 0])
            f = 1.0 + (0.75 - 1.0) * min(t, 5.0) / 5.0
            centre = np.array([-d[0] + sg * half, -d[1], -d[2]])
            q = (q - centre) * f + centre
            if t > 5.0:
                q = rot_x(q, 180.0 * (min(t, 15.0) - 5.0), -sg, -d)
            worst = max(worst, float(np.abs(pos[k] - (q + d)).max()))
        assert worst < 5e-6, worst


def test_shell_slot_and_transpose_tables_address_the_right_blocks(ppf):
    """builder.rs:1244-1496 for a shell scene: the fixed index table holds the upper triangle of (edges + hinge stencils + diagonal);
    face / hinge slot tables address block (row min, col max) or carry the lower-triangle sentinel; the transpose table lists,
    for every column, the (row, slot) of each strictly-upper block in it"""
    import ctypes

    h = ppf.host()
    p = ctypes.c_void_p(h.ppfh_create(b"fast-twist", ctypes.c_uint(12)))
    c = (ctypes.c_uint * 6)()
    h.ppfh_counts(p, c)
    nv, nf, ne, nh, nfix, nnz = [int(v) for v in c]
    vp = ctypes.c_void_p
    face = np.empty((nf, 3), np.uint32); edge = np.empty((ne, 2), np.uint32)
    h.ppfh_topology(p, face.ctypes.data_as(vp), edge.ctypes.data_as(vp), None)
    fs = np.empty((nf, 9), np.uint32); hs = np.empty((nh, 16), np.uint32); hg = np.empty((nh, 4), np.uint32)
    off = np.empty(nv + 1, np.uint32); col = np.empty(nnz, np.uint32); toff = np.empty(nv + 1, np.uint32); trs = np.empty((nnz - nv, 2), np.uint32)
    h.ppfh_slot_tables(p, fs.ctypes.data_as(vp), hs.ctypes.data_as(vp), hg.ctypes.data_as(vp), off.ctypes.data_as(vp), col.ctypes.data_as(vp),
                       toff.ctypes.data_as(vp), trs.ctypes.data_as(vp))
    h.ppfh_destroy(p)
    # the pattern: diagonal + edges + the two opposite vertices of every hinge, upper triangle
    want = {(i, i) for i in range(nv)}
    for a, b in edge:
        want.add((min(a, b), max(a, b)))
    for q in hg:
        for a in q:
            for b in q:
                want.add((min(a, b), max(a, b)))
    got = {(i, int(col[k])) for i in range(nv) for k in range(off[i], off[i + 1])}
    assert got == {(int(a), int(b)) for a, b in want}
    assert off[nv] == nnz and all((np.diff(col[off[i]:off[i + 1]].astype(np.int64)) > 0).all() for i in range(nv))

    def check(slot, a, b):
        if a > b:
            assert slot == 0xFFFFFFFF
        else:
            assert off[a] <= slot < off[a + 1] and col[slot] == b
    for i in range(nf):
        for a in range(3):
            for b in range(3):
                check(fs[i, 3 * a + b], face[i, a], face[i, b])
    for i in range(nh):
        r = [hg[i, 2], hg[i, 1], hg[i, 0], hg[i, 3]]
        for a in range(4):
            for b in range(4):
                check(hs[i, 4 * a + b], r[a], r[b])
    # transpose table: column j lists (row i, slot of (i, j)) for every strictly-upper block (i, j)
    n_t = 0
    for j in range(nv):
        for k in range(toff[j], toff[j + 1]):
            i, slot = int(trs[k, 0]), int(trs[k, 1])
            assert i < j and off[i] <= slot < off[i + 1] and col[slot] == j
            n_t += 1
    assert n_t == nnz - nv


def test_oracle_plastic_sheet_keeps_its_deformation(run_job):
    """membrane + bend plasticity (plasticity.cu): the run steps, and ends elsewhere than the elastic sheet"""
    a = run_job({"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 30})
    p = run_job({"lib": "oracle", "scene": "sheet-plastic", "n": 16, "state": None, "steps": 30})
    assert bool(a["ok"].all()) and bool(p["ok"].all()) and np.isfinite(p["curr"]).all()
    assert float(np.abs(a["curr"] - p["curr"]).max()) > 1e-3
    w = run_job({"lib": "oracle", "scene": "woven-plastic", "n": 32, "state": None, "steps": 6})
    assert bool(w["ok"].all())


def test_fix_xz_drags_the_horizontal_motion_above_the_threshold(run_job):
    """ParamSet.fix_xz (energy.cu:260-266, main.cu:1190-1208): the swinging sheet with a horizontal drag above y = -0.4 --
    vertices that stay above the threshold move less in x/z than in the free run, the ones far below are not held"""
    f = run_job({"lib": "oracle", "scene": "sheet-fixxz", "n": 16, "state": None, "steps": 25}, force_subprocess=True)
    u = run_job({"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 25}, force_subprocess=True)
    assert bool(f["ok"].all()) and bool(u["ok"].all())
    s0 = run_job({"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 0}, force_subprocess=True)["curr"]
    hf = np.hypot(f["curr"][:, 0] - s0[:, 0], f["curr"][:, 2] - s0[:, 2])
    hu = np.hypot(u["curr"][:, 0] - s0[:, 0], u["curr"][:, 2] - s0[:, 2])
    assert hu.max() > 0.02                                       # the free sheet swings
    high = (f["curr"][:, 1] > -0.2) & (u["curr"][:, 1] > -0.2) & (hu > 1e-3)
    assert high.sum() > 10 and float(np.median(hf[high] / hu[high])) < 0.7
    assert not np.array_equal(f["curr"], u["curr"])


def test_collision_windows_switch_an_object_out_of_contact(run_job):
    """init_collision_windows / refresh_collision_active (main.cu:48-110, 2230-2330; scene.rs:2471-2542): `headless` from a
    saved state at t = 0.68 where the moving sphere presses into the curtains.  A window that is open ([0, 100)) changes
    nothing, bit for bit; with the sphere's window closed ([0, 0.3)) its contacts are gone (the curtain-curtain ones stay),
    the step still passes its intersection check (the sphere is ignored there too) and the curtains are released."""
    state = os.path.join(REPO, "tests", "golden", "headless_n24", "state_00068.bin")
    job = {"lib": "oracle", "n": 24, "state": state, "steps": 2}
    plain = run_job(dict(job, scene="headless"), force_subprocess=True)
    opened = run_job(dict(job, scene="headless-cw-open"), force_subprocess=True)
    closed = run_job(dict(job, scene="headless-cw-closed"), force_subprocess=True)
    assert bool(plain["ok"].all()) and bool(opened["ok"].all()) and bool(closed["ok"].all())
    assert np.array_equal(plain["curr"], opened["curr"]) and np.array_equal(plain["log_num_contact"], opened["log_num_contact"])
    assert closed["log_num_contact"][0] < plain["log_num_contact"][0] - 20
    assert closed["log_num_contact"][-1] < 0.6 * plain["log_num_contact"][-1]
    assert float(np.abs(closed["curr"] - plain["curr"]).max()) > 1e-3


def test_torque_pin_spins_a_free_sheet_without_moving_its_centre(run_job):
    """PinOperation::Torque (scene.rs:2061-2198): the pinned strip stays free (no fix / pull pairs), the sheet gains angular
    momentum about the strip's longest principal axis while the torque is on (t <= 0.2 s) and keeps it afterwards; no net force"""
    r0 = run_job({"lib": "oracle", "scene": "sheet-torque", "n": 16, "state": None, "steps": 0}, force_subprocess=True)
    r = run_job({"lib": "oracle", "scene": "sheet-torque", "n": 16, "state": None, "steps": 30}, force_subprocess=True)
    assert bool(r["ok"].all()) and "fix_index" not in r or len(r.get("fix_index", [])) == 0
    x0, x, v = r0["curr"].astype(np.float64), r["curr"].astype(np.float64), (r["curr"] - r["prev"]).astype(np.float64)
    assert np.abs(x.mean(0) - x0.mean(0)).max() < 1e-5
    L = np.cross(x - x.mean(0), v).sum(0)
    assert abs(L[2]) > 0.99 * np.linalg.norm(L) and np.linalg.norm(L) > 1e-3
    assert np.abs(x - x0).max() > 5e-3


def test_inflation_pressure_blows_up_a_closed_shell_without_moving_it(run_job):
    """FaceParam.pressure on the `balloon` scene (a closed icosphere away from the origin, no gravity): the shell grows
    (mean radius +2..8 % after 0.2 s), stays round and its centre stays where it was -- the per-face pressure terms depend
    on the distance from the origin, their sum over the closed surface does not."""
    r0 = run_job({"lib": "oracle", "scene": "balloon", "n": 16, "state": None, "steps": 0}, force_subprocess=True)
    r = run_job({"lib": "oracle", "scene": "balloon", "n": 16, "state": None, "steps": 20}, force_subprocess=True)
    assert bool(r["ok"].all())
    x0, x = r0["curr"].astype(np.float64), r["curr"].astype(np.float64)
    rad0, rad = np.linalg.norm(x0 - x0.mean(0), axis=1), np.linalg.norm(x - x.mean(0), axis=1)
    assert 1.02 < rad.mean() / rad0.mean() < 1.08
    assert rad.std() < 0.01 * rad.mean()
    assert np.abs(x.mean(0) - x0.mean(0)).max() < 5e-4


def test_static_mesh_tables_of_the_drape_mesh_scene(ppf):
    """builder.rs:1823-1935 make_collision_mesh restated in host/builder.cpp: the two static objects merged, every undirected edge
    once (closed icosphere: E = 3F/2; open ground square: more), the scene's contact parameters on every element"""
    s = ppf.Session("drape-mesh", n=16, lib=ppf.LIB_ORACLE)
    vert, face, edge, vpar, epar, fpar = s.static_mesh()
    n_ball_f = 20 * 4 ** 2                                         # icosphere, two subdivisions at n = 16
    n_ground_f = 2 * 7 * 7                                        # gen_square(8, ...): 8 vertices per side
    assert len(face) == n_ball_f + n_ground_f and face.max() == len(vert) - 1
    assert len(vert) == (10 * 4 ** 2 + 2) + 8 * 8
    und = {tuple(sorted(e)) for e in edge.tolist()}
    assert len(und) == len(edge)                                   # unique
    want = {tuple(sorted((int(f[k]), int(f[(k + 1) % 3])))) for f in face for k in range(3)}
    assert und == want and len(edge) == 3 * n_ball_f // 2 + (3 * n_ground_f + 4 * 7) // 2
    for par in (vpar, epar, fpar):
        assert np.allclose(par[:, 0], 1e-3, rtol=1e-6) and np.allclose(par[:, 1], 0.0) and np.allclose(par[:, 2], 0.3, rtol=1e-6)
    ball = vert[:10 * 4 ** 2 + 2]
    assert np.allclose(np.linalg.norm(ball, axis=1), 0.4, atol=1e-5) and np.allclose(vert[len(ball):, 1], -0.42, atol=1e-6)

"""Analytic colliders (invisible spheres / bowls / walls): the oracle's plain-C restatement and the CUDA backend's
header (csrc/math_collider.cuh compiled for the host) against each other and against first principles.

The reference keeps this code inside contact.cu (embed_vertex_constraint_force_hessian 1154-1383,
vertex_constraint_line_search 2318-2470), not in a header, so it cannot be compiled for the host: leaf parity with
the reference is pinned through the whole-step golden fixtures of tests/golden/drape_n32 instead (produced by the
reference on a B200).  Here: (1) oracle == backend header, bit for bit, over random configurations of every branch;
(2) the push force is stiffness x the gradient of the cubic push energy -d^3 / (3 ghat) (push.hpp:15-50);
(3) the line search returns the first crossing and flags prescribed vertices driven through a collider.
"""
import ctypes as C
import os

import numpy as np
import pytest

from test_oracle_cpu import _build_csrc_shim, fp, same_bits

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libs(built):
    o = C.CDLL(os.path.join(REPO, "oracle", "build", "liboracle_math.so"))
    b = C.CDLL(_build_csrc_shim())
    for l in (o, b):
        l.orc_collider_vertex.restype = C.c_uint
        l.orc_collider_toi.restype = C.c_float
    return o, b


def _scene(rng):
    """two spheres (one random flavour each) and two floors, 10 floats per record"""
    sp = np.zeros((2, 10), np.float32); fl = np.zeros((2, 10), np.float32)
    for j in range(2):
        sp[j, :3] = rng.uniform(-0.2, 0.2, 3); sp[j, 3] = rng.choice([1e-3, 5e-3]); sp[j, 4] = rng.choice([0.0, 0.3])
        sp[j, 5] = rng.uniform(0.3, 0.6); sp[j, 6] = rng.choice([1.0, 0.05]); sp[j, 7] = rng.integers(0, 2); sp[j, 8] = rng.integers(0, 2); sp[j, 9] = rng.integers(0, 2)
        up = rng.normal(size=3); up /= np.linalg.norm(up)
        fl[j, :3] = rng.uniform(-0.5, 0.5, 3); fl[j, 3] = rng.choice([1e-3, 5e-3]); fl[j, 4] = rng.choice([0.0, 0.5]); fl[j, 5] = rng.choice([1.0, 0.02])
        fl[j, 6:9] = up; fl[j, 9] = rng.integers(0, 2)
    return sp, fl


def _call_vertex(l, x, dx, mass, vf, H, sp, fl, mode):
    f = np.zeros(3, np.float32); Hh = np.zeros(9, np.float32)
    n = l.orc_collider_vertex(fp(x), fp(dx), C.c_float(mass), C.c_float(vf), fp(H), fp(sp), C.c_uint(len(sp)), fp(fl), C.c_uint(len(fl)), C.c_int(mode),
                              C.c_float(1e-5), C.c_float(0.01), fp(f), fp(Hh))
    return int(n), f, Hh


def test_oracle_and_backend_header_agree_bit_for_bit(libs):
    o, b = libs
    rng = np.random.default_rng(3)
    hits = 0
    for _ in range(1500):
        sp, fl = _scene(rng)
        x = rng.uniform(-0.7, 0.7, 3).astype(np.float32)
        if rng.random() < 0.5:                                      # near a sphere surface (inside the gap)
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            x = (sp[0, :3] + d * (sp[0, 5] + rng.uniform(-2, 2) * sp[0, 3])).astype(np.float32)
        dx = (rng.normal(size=3) * 1e-3).astype(np.float32)
        A = rng.normal(size=(3, 3)); H = (A @ A.T).astype(np.float32).reshape(9)
        mode = int(rng.integers(0, 3))
        n0, f0, H0 = _call_vertex(o, x, dx, 0.01, 0.4, H, sp, fl, mode)
        n1, f1, H1 = _call_vertex(b, x, dx, 0.01, 0.4, H, sp, fl, mode)
        assert n0 == n1 and same_bits(f0, f1).all() and same_bits(H0, H1).all()
        hits += n0
        y1 = (x + rng.normal(size=3) * 0.2).astype(np.float32)
        for pinned in (0, 1):
            i0 = C.c_int(0); i1 = C.c_int(0)
            t0 = o.orc_collider_toi(fp(x), fp(y1), pinned, fp(sp), 2, fp(fl), 2, C.c_float(1.25), C.byref(i0))
            t1 = b.orc_collider_toi(fp(x), fp(y1), pinned, fp(sp), 2, fp(fl), 2, C.c_float(1.25), C.byref(i1))
            assert same_bits(np.float32(t0), np.float32(t1)) and i0.value == i1.value
    assert hits > 300


@pytest.mark.parametrize("which", [0, 1])
def test_push_force_is_stiffness_times_the_energy_gradient(libs, which):
    l = libs[which]
    rng = np.random.default_rng(9)
    none = np.zeros((0, 10), np.float32)
    worst = 0.0
    for case in range(200):
        ghat, mass = 2e-3, 0.02
        A = rng.normal(size=(3, 3)); Hm = A @ A.T
        H = Hm.astype(np.float32).T.reshape(9)
        if case % 2 == 0:   # static floor without friction: E = -d^3 / (3 ghat), d = (x - ground - ghat up) . up < 0
            up = rng.normal(size=3); up /= np.linalg.norm(up)
            ground = rng.uniform(-0.3, 0.3, 3)
            gap = rng.uniform(0.1, 0.9) * ghat
            x = (ground + up * gap + np.cross(up, rng.normal(size=3)) * 0.1)
            fl = np.array([[*ground, ghat, 0.0, 1.0, *up, 0]], np.float32)
            n, f, Hh = _call_vertex(l, x.astype(np.float32), np.zeros(3, np.float32), mass, 0.0, H, none, fl, 0)
            x32 = x.astype(np.float32).astype(np.float64); g32 = fl[0, :3].astype(np.float64); u32 = fl[0, 6:9].astype(np.float64)
            d = (x32 - g32 - np.float64(np.float32(ghat)) * u32) @ u32
            gap_now = (x32 - g32) @ u32
            k = u32 @ (Hm.astype(np.float32).astype(np.float64) @ u32) + mass / gap_now ** 2
            want = k * (-(d * d) / ghat) * u32
            curv = k * (-2.0 * d / ghat)
            want_H = curv * np.outer(u32, u32)
        else:               # static solid sphere without friction
            c = rng.uniform(-0.2, 0.2, 3); r = 0.4
            nrm = rng.normal(size=3); nrm /= np.linalg.norm(nrm)
            gap = rng.uniform(0.1, 0.9) * ghat
            x = c + nrm * (r + gap)
            sp = np.array([[*c, ghat, 0.0, r, 1.0, 0, 0, 0]], np.float32)
            n, f, Hh = _call_vertex(l, x.astype(np.float32), np.zeros(3, np.float32), mass, 0.0, H, sp, none, 0)
            x32 = x.astype(np.float32).astype(np.float64); c32 = sp[0, :3].astype(np.float64)
            dist = np.linalg.norm(x32 - c32); n32 = (x32 - c32) / dist
            d = dist - (np.float64(np.float32(r)) + np.float64(np.float32(ghat)))
            k = n32 @ (Hm.astype(np.float32).astype(np.float64) @ n32) + mass / (dist - np.float64(np.float32(r))) ** 2
            want = k * (-(d * d) / ghat) * n32
            want_H = k * (-2.0 * d / ghat) * np.outer(n32, n32)
        assert n == 1
        worst = max(worst, np.abs(f - want).max() / np.abs(want).max(), np.abs(Hh.reshape(3, 3).T - want_H).max() / np.abs(want_H).max())
    assert worst < 2e-3, worst
    # friction only adds a tangential force opposing nothing when the vertex has not moved, and never changes the count
    fl = np.array([[0, 0, 0, ghat, 0.5, 1.0, 0, 1, 0, 0]], np.float32)
    x = np.array([0.1, 0.5 * ghat, 0.2], np.float32)
    n0, f0, _ = _call_vertex(l, x, np.zeros(3, np.float32), 0.02, 0.5, np.zeros(9, np.float32), none, fl, 0)
    n1, f1, _ = _call_vertex(l, x, np.array([1e-3, 0, 0], np.float32), 0.02, 0.5, np.zeros(9, np.float32), none, fl, 0)
    assert n0 == n1 == 1 and f0[0] == 0 and f1[0] != 0 and f0[1] == f1[1]
    # the gradient points against the wall normal (sign convention: "force" arrays hold dE/dx) and friction along the slip
    assert f0[1] < 0 and f1[0] > 0
    # zero thickness cutoff: a vertex deeper than `thickness` below the wall passes through
    fl[0, 5] = 1e-4
    n2, f2, _ = _call_vertex(l, np.array([0.1, -5e-4, 0.2], np.float32), np.zeros(3, np.float32), 0.02, 0.5, np.zeros(9, np.float32), none, fl, 0)
    assert n2 == 0 and not f2.any()


@pytest.mark.parametrize("which", [0, 1])
def test_collider_line_search(libs, which):
    l = libs[which]
    none = np.zeros((0, 10), np.float32)
    fl = np.array([[0, 0, 0, 1e-3, 0, 1.0, 0, 1, 0, 0]], np.float32)
    inf = C.c_int(0)
    y0 = np.array([0, 0.1, 0], np.float32)
    # free vertex heading through the wall: stops exactly at the wall (t = h0 / (h0 - h1), in units of the 1.25 x extrapolated step)
    t = l.orc_collider_toi(fp(y0), fp(np.array([0, -0.1, 0], np.float32)), 0, fp(none), 0, fp(fl), 1, C.c_float(1.25), C.byref(inf))
    assert abs(t - 1.25 * (0.1 / (0.1 + 0.15))) < 1e-6 and inf.value == 0
    # staying above: no cut
    t = l.orc_collider_toi(fp(y0), fp(np.array([0.3, 0.05, 0], np.float32)), 0, fp(none), 0, fp(fl), 1, C.c_float(1.25), C.byref(inf))
    assert t == np.float32(1.25)
    # a pinned vertex crossing from outside to inside is flagged, not clamped
    t = l.orc_collider_toi(fp(y0), fp(np.array([0, -0.1, 0], np.float32)), 1, fp(none), 0, fp(fl), 1, C.c_float(1.25), C.byref(inf))
    assert t == np.float32(1.25) and inf.value == 1
    # sphere: radial crossing at r = radius; inverted sphere (container) keeps the vertex inside
    sp = np.array([[0, 0, 0, 1e-3, 0, 0.5, 1.0, 0, 0, 0]], np.float32)
    t = l.orc_collider_toi(fp(np.array([0.8, 0, 0], np.float32)), fp(np.array([0.4, 0, 0], np.float32)), 0, fp(sp), 1, fp(none), 0, C.c_float(1.25), C.byref(inf))
    assert abs(t - 1.25 * (0.3 / 0.5)) < 1e-6
    sp[0, 8] = 1
    t = l.orc_collider_toi(fp(np.array([0.2, 0, 0], np.float32)), fp(np.array([0.6, 0, 0], np.float32)), 0, fp(sp), 1, fp(none), 0, C.c_float(1.25), C.byref(inf))
    assert abs(t - 1.25 * (0.3 / 0.5)) < 1e-6
    # kinematic colliders take no part in the line search
    sp[0, 9] = 1
    t = l.orc_collider_toi(fp(np.array([0.2, 0, 0], np.float32)), fp(np.array([0.6, 0, 0], np.float32)), 0, fp(sp), 1, fp(none), 0, C.c_float(1.25), C.byref(inf))
    assert t == np.float32(1.25)


def test_oracle_drapes_a_sheet_over_sphere_and_floor(run_job):
    """whole steps on the CPU: the sheet lands on the invisible sphere, its corners on the invisible floor; nothing tunnels"""
    o = run_job({"lib": "oracle", "scene": "drape", "n": 24, "state": None, "steps": 140})
    assert bool(o["ok"].all()) and np.isfinite(o["curr"]).all()
    x = o["curr"].astype(np.float64)
    assert np.linalg.norm(x, axis=1).min() > 0.4 and x[:, 1].min() > -0.08      # outside the sphere (r = 0.4), above the floor (y = -0.08)
    assert o["log_num_contact"].max() > 100
    assert np.linalg.norm(x, axis=1).min() < 0.4 + 1.01e-3                       # resting inside the contact gap


def test_oracle_lifts_a_sheet_with_a_kinematic_sphere(run_job):
    """keyframed (kinematic) collider: penalty with the stiffness m / ghat^2 and no line search (contact.cu:1230-1263)"""
    o = run_job({"lib": "oracle", "scene": "lift", "n": 24, "state": None, "steps": 70})
    assert bool(o["ok"].all()) and np.isfinite(o["curr"]).all()
    assert o["log_num_contact"].max() > 100
    x = o["curr"].astype(np.float64)
    t = min(float(o["time"][0]), 0.6) / 0.6
    c = np.array([0.1, -0.9 + 0.7 * t, 0.2])
    assert np.linalg.norm(x - c, axis=1).min() > 0.5 - 2e-3           # a penalty, not a hard constraint: at most a fraction of the gap inside
    assert x[:, 1].max() > 0.2                                          # the sheet has been lifted above its pinned corners


def test_isotropic_shell_face_oracle_equals_backend_header(libs):
    """ARAP / StVK / SNHk shell faces: the oracle's plain C and the backend's header perform the same fp32 operations"""
    o, b = libs
    rng = np.random.default_rng(17)
    for _ in range(300):
        rest = rng.uniform(-1, 1, size=(3, 2)) * 0.05
        Dm = np.stack([rest[1] - rest[0], rest[2] - rest[0]], axis=1)
        if abs(np.linalg.det(Dm)) < 1e-5:
            continue
        ir = np.linalg.inv(Dm)
        ir32 = np.array([ir[0, 0], ir[1, 0], ir[0, 1], ir[1, 1]], np.float32)
        x = (rng.uniform(-1, 1, size=(3, 3)) * 0.05).astype(np.float32)
        model = int(rng.choice([0, 1, 3]))
        out = []
        for l in (o, b):
            g9 = np.zeros(9, np.float32); H81 = np.zeros(81, np.float32)
            l.orc_iso_face(fp(x), fp(ir32), C.c_int(model), C.c_float(30.0), C.c_float(20.0), C.c_float(0.01), C.c_float(1e-2), fp(g9), fp(H81))
            out.append((g9, H81))
        assert same_bits(out[0][0], out[1][0]).all() and same_bits(out[0][1], out[1][1]).all()


def test_rayleigh_damping_oracle_equals_backend_header_and_the_definition(libs):
    """add_stiffness_damping / add_stiffness_damping_lagged (energy.cu:88-180): gradient += (beta/dt) K u, Hessian *= 1 + beta/dt
    (lagged: gradient += (beta/dt) K_lag u, Hessian += (beta/dt) K_lag); oracle and backend header bit for bit, both against numpy"""
    o, b = libs
    rng = np.random.default_rng(4)
    for n in (2, 3, 4):
        for lagged in (0, 1):
            for _ in range(40):
                A = rng.normal(size=(3 * n, 3 * n)); K = (A @ A.T).astype(np.float32)
                B = rng.normal(size=(3 * n, 3 * n)); Kl = (B @ B.T).astype(np.float32)
                u = (rng.normal(size=3 * n) * 1e-2).astype(np.float32); f = rng.normal(size=3 * n).astype(np.float32)
                beta, dt = 0.03, 0.01
                outs = []
                for l in (o, b):
                    Kc = np.ascontiguousarray(K.T).reshape(-1).copy(); Klc = np.ascontiguousarray(Kl.T).reshape(-1).copy(); fc = f.copy()
                    l.orc_damping(C.c_int(n), fp(Kc), fp(Klc), fp(u), C.c_float(beta), C.c_float(dt), fp(fc), C.c_int(lagged))
                    outs.append((Kc, fc))
                assert same_bits(outs[0][0], outs[1][0]).all() and same_bits(outs[0][1], outs[1][1]).all()
                s = np.float64(np.float32(beta)) / np.float64(np.float32(dt))
                Kout = outs[0][0].reshape(3 * n, 3 * n).T
                op = Kl if lagged else K
                assert np.allclose(outs[0][1], f + s * (op.astype(np.float64) @ u), rtol=2e-5, atol=1e-5)
                want = K + s * Kl if lagged else K * (1 + s)
                assert np.allclose(Kout, want, rtol=2e-6, atol=2e-5 * float(np.abs(want).max()))
    # beta = 0 or dt = 0: nothing changes
    K = np.eye(6, dtype=np.float32).reshape(-1); f = np.ones(6, np.float32); u = np.ones(6, np.float32)
    for beta, dt in ((0.0, 0.01), (0.1, 0.0)):
        Kc, fc = K.copy(), f.copy()
        o.orc_damping(C.c_int(2), fp(Kc), fp(K), fp(u), C.c_float(beta), C.c_float(dt), fp(fc), C.c_int(0))
        assert np.array_equal(Kc, K) and np.array_equal(fc, f)


def test_stitch_oracle_equals_backend_header_and_the_spring_energy(libs):
    """embed_stitch_force_hessian (energy.cu:796-886): oracle and backend header bit for bit; below the length cap the
    gradient is that of the spring energy (l0 / 2)(l / l0 - 1)^2 of the distance between the two barycentric points; the
    slot gradients sum to zero (source and target weights both sum to one); the Hessian is symmetric PSD."""
    o, b = libs
    rng = np.random.default_rng(12)
    worst = 0.0
    for case in range(200):
        x = rng.uniform(-0.05, 0.05, size=(6, 3)).astype(np.float32)
        ws = rng.dirichlet([1, 1, 1]); wd = rng.dirichlet([1, 1, 1])
        if case % 3 == 0:
            ws = np.array([1.0, 0, 0]); x[1] = x[0]; x[2] = x[0]         # plain vertex as the source: {s, s, s}, {1, 0, 0}
        w = np.concatenate([ws, wd]).astype(np.float32)
        ghat = rng.choice([1e-3, 4e-3], size=6).astype(np.float32); off = rng.choice([0.0, 1e-3], size=6).astype(np.float32)
        lf = 10.0 if case % 2 else 1e6                                    # capped / uncapped
        outs = []
        for l in (o, b):
            f = np.zeros(18, np.float32); H = np.zeros(324, np.float32)
            l.orc_stitch(fp(x), fp(ghat), fp(off), fp(w), C.c_float(2.5), C.c_float(lf), fp(f), fp(H))
            outs.append((f, H))
        assert same_bits(outs[0][0], outs[1][0]).all() and same_bits(outs[0][1], outs[1][1]).all()
        f, H = outs[0]
        assert np.abs(f.reshape(6, 3).sum(axis=0)).max() <= 2e-5 * max(np.abs(f).max(), 1e-9)
        Hm = H.reshape(18, 18).T.astype(np.float64)
        assert np.abs(Hm - Hm.T).max() <= 1e-6 * max(np.abs(Hm).max(), 1e-12)
        assert np.linalg.eigvalsh(0.5 * (Hm + Hm.T)).min() >= -1e-5 * np.abs(Hm).max()
        if lf > 1e5:
            l0 = float((w.astype(np.float64) * ghat).sum() / 2.0)

            def energy(y):
                t = (w[:3, None] * y[:3]).sum(axis=0) - (w[3:, None] * y[3:]).sum(axis=0)
                return 2.5 * 0.5 * l0 * (np.linalg.norm(t) / l0 - 1.0) ** 2
            x64 = x.astype(np.float64)
            num = np.zeros(18)
            for k in range(18):
                d = np.zeros(18); d[k] = 1e-7
                num[k] = (energy(x64 + d.reshape(6, 3)) - energy(x64 - d.reshape(6, 3))) / 2e-7
            worst = max(worst, np.abs(f - num).max() / max(np.abs(num).max(), 1e-9))
    assert worst < 2e-3, worst


def test_plastic_creep_oracle_equals_backend_header_and_relaxes_toward_the_dead_zone(libs):
    """plasticity.cu:21-86, 147-245: oracle and backend header bit for bit; with alpha -> 1 (a huge rate) the rebuilt inverse rest
    matrix puts the singular values of the same deformed face on the dead-zone edge 1 +- threshold; rest angles move toward the
    current angle by alpha x (deviation - threshold) and stay put inside the dead zone"""
    o, b = libs
    for l in libs:
        l.orc_face_plasticity.restype = C.c_int
        l.orc_hinge_plasticity.restype = C.c_float
        l.orc_rod_bend_plasticity.restype = C.c_float
    rng = np.random.default_rng(31)
    n_changed = 0
    for case in range(200):
        rest = np.array([[0.0, 0.0], [1.0, 0.1], [0.2, 0.9]]) * 0.05
        Dm = np.stack([rest[1] - rest[0], rest[2] - rest[0]], axis=1)
        ir = np.linalg.inv(Dm)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        s1, s2 = rng.uniform(0.9, 1.2), rng.uniform(0.9, 1.2)
        Fm = q[:, :2] @ np.diag([s1, s2])
        x0 = rng.uniform(-1, 1, 3)
        x = np.stack([x0, x0 + Fm @ Dm[:, 0], x0 + Fm @ Dm[:, 1]]).astype(np.float32)
        ir32 = np.array([ir[0, 0], ir[1, 0], ir[0, 1], ir[1, 1]], np.float32)
        thr = 0.05
        outs = []
        for l in (o, b):
            irc = ir32.copy()
            ch = l.orc_face_plasticity(fp(x), fp(irc), C.c_float(thr), C.c_float(1e4 if case % 2 else 3.0), C.c_float(0.01))
            outs.append((ch, irc))
        assert outs[0][0] == outs[1][0] and same_bits(outs[0][1], outs[1][1]).all()
        expect_change = abs(s1 - 1) > thr * 1.001 or abs(s2 - 1) > thr * 1.001
        if abs(abs(s1 - 1) - thr) > 1e-3 and abs(abs(s2 - 1) - thr) > 1e-3:
            assert bool(outs[0][0]) == expect_change
        if outs[0][0] and case % 2:          # alpha = 1: singular values of the new F sit on the dead-zone edge (or were inside it)
            n_changed += 1
            irn = outs[0][1].astype(np.float64)
            irm = np.array([[irn[0], irn[2]], [irn[1], irn[3]]])
            Ds = np.stack([x[1] - x[0], x[2] - x[0]], axis=1).astype(np.float64)
            S = np.linalg.svd(Ds @ irm, compute_uv=False)
            for sv, s_old in zip(sorted(S), sorted([s1, s2])):
                want = s_old if abs(s_old - 1) <= thr else (1 - thr if s_old < 1 else 1 + thr)
                assert abs(sv - want) < 2e-3, (S, s1, s2)
    assert n_changed > 30
    for _ in range(200):                      # rest angles
        hx = rng.uniform(-1, 1, size=(4, 3)).astype(np.float32)
        rest = float(rng.uniform(-0.5, 0.5))
        a = o.orc_hinge_plasticity(fp(hx), C.c_float(rest), C.c_float(0.05), C.c_float(2.0), C.c_float(0.01))
        c = b.orc_hinge_plasticity(fp(hx), C.c_float(rest), C.c_float(0.05), C.c_float(2.0), C.c_float(0.01))
        assert same_bits(np.float32(a), np.float32(c))
        sx = rng.uniform(-1, 1, size=(3, 3)).astype(np.float32)
        a = o.orc_rod_bend_plasticity(fp(sx), C.c_float(np.pi), C.c_float(0.05), C.c_float(2.0), C.c_float(0.01))
        c = b.orc_rod_bend_plasticity(fp(sx), C.c_float(np.pi), C.c_float(0.05), C.c_float(2.0), C.c_float(0.01))
        assert same_bits(np.float32(a), np.float32(c))
        e0, e1 = sx[0].astype(np.float64) - sx[1], sx[2].astype(np.float64) - sx[1]
        theta = np.arccos(np.clip(e0 @ e1 / np.linalg.norm(e0) / np.linalg.norm(e1), -1, 1))
        delta = theta - np.float32(np.pi)
        alpha = 1 - np.exp(-2.0 * 0.01)
        want = np.pi + (alpha * (delta - np.sign(delta) * 0.05) if abs(delta) > 0.05 else 0.0)
        assert abs(a - want) < 1e-5


def test_quadratic_and_logarithmic_barriers_bit_exact(libs, built):
    """ParamSet.barrier = quad / log (barrier/quadratic.hpp:13-41, barrier/logarithm.hpp:14-44): gradient and curvature of
    the oracle and of the backend header against vectors produced by the reference's own headers (tests/golden/leaf_math.npz),
    bit for bit (the host build of the reference takes std::log, as do both restatements when compiled for the host)."""
    o, b = libs
    gold = np.load(os.path.join(REPO, "tests", "golden", "leaf_math.npz"))
    g, kind, want = gold["barrier_g"], gold["barrier_kind"], gold["barrier_out"]
    assert set(kind.tolist()) == {1, 2} and np.isinf(want).any() and (want == 0).any()
    for l in (o, b):
        got = np.zeros_like(want)
        for i in range(len(g)):
            l.orc_barrier(C.c_int(int(kind[i])), C.c_float(g[i]), C.c_float(1e-3), C.c_float(2e-4), fp(got[i]))
        assert same_bits(got, want).all()
    # kind 0 is the cubic barrier of the existing vectors
    out = np.zeros(2, np.float32)
    o.orc_barrier(C.c_int(0), C.c_float(5e-4), C.c_float(1e-3), C.c_float(0.0), fp(out))
    ref = np.zeros(2, np.float32)
    b.orc_barrier(C.c_int(0), C.c_float(5e-4), C.c_float(1e-3), C.c_float(0.0), fp(ref))
    assert same_bits(out, ref).all() and out[0] < 0 < out[1]
    # the gradient is the derivative of the energy: quad (ghat - g)^2, log -(g - ghat)^2 log(g / ghat)
    for k, energy in ((1, lambda x: (1e-3 - x) ** 2), (2, lambda x: -(x - 1e-3) ** 2 * np.log(x / 1e-3))):
        for x in (2e-4, 5e-4, 9e-4):
            o.orc_barrier(C.c_int(k), C.c_float(x), C.c_float(1e-3), C.c_float(0.0), fp(out))
            h = 1e-7
            assert abs(out[0] - (energy(x + h) - energy(x - h)) / (2 * h)) < 2e-3 * abs(out[0]) + 1e-9
            assert abs(out[1] - (energy(x + h) - 2 * energy(x) + energy(x - h)) / h ** 2) < 2e-2 * abs(out[1]) + 1e-6


@pytest.mark.parametrize("kind", ["quad", "log"])
def test_oracle_runs_the_twist_with_the_other_barriers(run_job, kind):
    """the contact-rich fast twist with ParamSet.barrier = quad / log: the step converges, contacts are found and the
    result stays free of intersections (the step's own final check) while differing from the cubic-barrier run"""
    r = run_job({"lib": "oracle", "scene": "fast-twist@" + kind, "n": 12, "state": None, "steps": 30}, force_subprocess=True)
    c = run_job({"lib": "oracle", "scene": "fast-twist", "n": 12, "state": None, "steps": 30}, force_subprocess=True)
    assert r["ok"].all() and np.asarray(r["log_num_contact"]).max() > 0
    assert np.isfinite(r["curr"]).all()
    assert not np.array_equal(r["curr"], c["curr"])


def test_static_collision_mesh_math_oracle_equals_backend_header(libs):
    """Constraint.mesh contacts (contact.cu:899-1151): dynamic vertex / face / edge against a static face / vertex / edge --
    the oracle's plain C and the backend's math_cmesh.cuh compiled for the host, bit for bit, on random close configurations;
    plus first principles: the force pushes the dynamic element away from the static one and vanishes beyond offset + ghat."""
    o, b = libs
    rng = np.random.default_rng(11)
    hits = [0, 0, 0]
    for trial in range(600):
        kind = trial % 3
        n = (1, 3, 2)[kind]
        ghat, offset = np.float32(2e-3), np.float32(rng.choice([0.0, 5e-4]))
        sta_n = (3, 1, 2)[kind]
        sta = rng.uniform(-0.05, 0.05, size=(sta_n, 3)).astype(np.float32)
        sta[:, 1] = rng.uniform(-2e-4, 2e-4, size=sta_n)                       # static element near the plane y = 0
        height = rng.uniform(float(offset) + 2e-4, float(offset + ghat) * 1.3)  # inside the gap most of the time
        dyn_eval = rng.uniform(-0.03, 0.03, size=(n, 3)).astype(np.float32)
        dyn_eval[:, 1] = height + rng.uniform(0, 2e-4, size=n)
        if kind == 2:                                                           # crossing edges
            sta[:, 0] = [-0.05, 0.05]; sta[:, 2] = 0.0
            dyn_eval[:, 2] = [-0.04, 0.04]; dyn_eval[:, 0] = rng.uniform(-0.01, 0.01, size=2)
        if kind == 1:                                                           # the static vertex under the dynamic face
            dyn_eval[:, 0] = [-0.03, 0.03, 0.0]; dyn_eval[:, 2] = [-0.02, -0.02, 0.03]; sta[0, 0] = sta[0, 2] = 0.0
        if kind == 0:
            sta[:, 0] = [-0.05, 0.05, 0.0]; sta[:, 2] = [-0.04, -0.04, 0.05]; dyn_eval[0, 0] = dyn_eval[0, 2] = 0.0
        dyn_start = (dyn_eval + rng.normal(0, 3e-4, size=dyn_eval.shape)).astype(np.float32)
        dyn = np.concatenate([dyn_eval.ravel(), dyn_start.ravel()]).astype(np.float32)
        mass = rng.uniform(1e-4, 1e-3, size=3).astype(np.float32)
        par = np.concatenate([[offset, ghat, rng.choice([0.0, 0.4]), 1e-5], mass]).astype(np.float32)
        A = rng.normal(size=(3 * n, 3 * n)); A = (A @ A.T * 10.0)
        B = np.zeros((n, n, 9), np.float32)
        for i in range(n):
            for j in range(n):
                B[i, j] = A[3 * i:3 * i + 3, 3 * j:3 * j + 3].T.ravel()          # column-major blocks
        outs = []
        for l in (o, b):
            out = np.full(16, np.nan, np.float32)
            l.orc_cmesh(C.c_int(kind), fp(dyn), fp(np.ascontiguousarray(sta.ravel())), fp(par), fp(np.ascontiguousarray(B.ravel())), fp(out))
            outs.append(out)
        assert same_bits(outs[0], outs[1]).all(), (kind, outs)
        out = outs[0]
        if out[0]:
            hits[kind] += 1
            assert out[2] < 0 or abs(out[2]) > 0                                # a force exists ...
            assert -out[2] > 0                                                  # ... and f = dE/dx points down: the element is pushed up (-f)
            assert abs(out[13:13 + n].sum() - 1.0) < 1e-5
            H = out[4:13].reshape(3, 3)
            assert np.allclose(H, H.T, atol=1e-6 * np.abs(H).max()) and np.linalg.eigvalsh(H.astype(np.float64)).min() > -1e-5 * np.abs(H).max()
        else:
            assert height + 2e-4 >= float(offset + ghat) - 4e-4                 # misses only near / beyond the gap
    assert min(hits) > 50


def test_oracle_drapes_a_sheet_over_static_meshes(run_job):
    """`drape-mesh`: the sheet thrown onto a static icosphere standing on a static ground square (Constraint.mesh): contacts
    appear, the line search cuts the steps, nothing passes through the meshes (the step's own intersection check covers the
    dynamic edges against the static faces) and the sheet ends up lying on the ball."""
    r = run_job({"lib": "oracle", "scene": "drape-mesh", "n": 16, "state": None, "steps": 14}, force_subprocess=True)
    assert bool(r["ok"].all())
    assert r["log_num_contact"].max() > 15 and r["log_toi"].min() < 0.5
    x = r["curr"].astype(np.float64)
    assert np.isfinite(x).all()
    assert np.linalg.norm(x, axis=1).min() > 0.37          # the icosphere's faces lie up to 4 % inside the radius 0.4
    assert x[:, 1].min() > -0.42


def test_torque_pin_math_oracle_equals_backend_header_and_first_principles(libs):
    """Torque pins (energy.cu:20-86 eigen_sym3x3, 643-701 group pre-pass, 221-246 per-vertex force): oracle's plain C and the
    backend's math_torque.cuh bit for bit; the axis is the requested principal direction of the group oriented toward the hint
    vertex; the forces carry no net force and exactly the requested torque about that axis."""
    o, b = libs
    rng = np.random.default_rng(21)
    for trial in range(60):
        n = int(rng.integers(6, 40))
        scale = np.array([1.0, 0.35, 0.08]) * rng.uniform(0.05, 0.5)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        x = ((rng.normal(size=(n, 3)) * scale) @ q.T + rng.uniform(-1, 1, 3)).astype(np.float32)
        uniform = trial % 2 == 0
        mass = (np.full(n, 5e-4) if uniform else rng.uniform(1e-4, 1e-3, n)).astype(np.float32)
        comp, hint, mag = trial % 3, int(rng.integers(0, n)), np.float32(rng.choice([-0.3, 0.7]))
        outs = []
        for l in (o, b):
            out = np.zeros(7 + 12 * n, np.float32)
            l.orc_torque(fp(np.ascontiguousarray(x.ravel())), fp(mass), C.c_uint(n), C.c_uint(comp), C.c_uint(hint), C.c_float(mag), fp(out))
            outs.append(out)
        assert same_bits(outs[0], outs[1]).all()
        out = outs[0].astype(np.float64)
        center, axis = out[0:3], out[3:6]
        assert np.allclose(center, (x * mass[:, None]).sum(0) / mass.sum(), atol=1e-5)
        cov = np.cov((x - center).T.astype(np.float64), bias=True)
        w, V = np.linalg.eigh(cov)
        want = V[:, ::-1][:, comp]
        assert abs(abs(axis @ want) - 1.0) < 2e-3 and abs(np.linalg.norm(axis) - 1.0) < 1e-5
        assert axis @ (x[hint] - center) >= -1e-6
        f = -out[7:].reshape(n, 12)[:, :3]                          # the shim returns the gradient; the force is its negative
        if uniform:                                                 # the centre is the MASS-weighted centroid: no net force for equal masses
            assert np.abs(f.sum(0)).max() < 2e-3 * np.abs(f).max() * n
        torque = np.cross(x - center, f).sum(0)
        assert abs(torque @ axis - float(mag)) < 2e-3 * abs(float(mag))
        H = out[7:].reshape(n, 12)[0, 3:].reshape(3, 3)
        assert np.allclose(H, H.T) and np.linalg.eigvalsh(H).min() > -1e-6 * max(np.abs(H).max(), 1e-12)


def test_inflation_pressure_oracle_equals_backend_header_and_the_volume_energy(libs):
    """FaceParam.pressure (inflate.hpp:16-150): E = -(P/6) x0 . (x1 x x2) per face.  Oracle and backend header bit for bit; the
    gradient is the derivative of E; the Hessian is positive semi-definite, never exceeds the positive part of the exact
    Hessian and equals its exact PSD projection when the face's tetrahedron with the origin is positively oriented."""
    o, b = libs
    rng = np.random.default_rng(31)
    exact = 0
    for trial in range(80):
        x = rng.uniform(-1, 1, 9).astype(np.float32)
        P = np.float32(rng.uniform(0.5, 4.0))
        res = []
        for l in (o, b):
            g = np.zeros(9, np.float32); H = np.zeros(81, np.float32)
            l.orc_inflate(C.c_float(P), fp(x), fp(g), fp(H))
            res.append((g, H))
        assert same_bits(res[0][0], res[1][0]).all() and same_bits(res[0][1], res[1][1]).all()
        g, H = res[0][0].astype(np.float64), res[0][1].reshape(9, 9).T.astype(np.float64)
        xd = x.astype(np.float64)
        E = lambda y: -(float(P) / 6.0) * np.linalg.det(y.reshape(3, 3).T)
        h = 1e-4
        gfd = np.array([(E(xd + h * np.eye(9)[i]) - E(xd - h * np.eye(9)[i])) / (2 * h) for i in range(9)])
        assert np.abs(g - gfd).max() < 1e-5
        He = np.zeros((9, 9))
        I = np.eye(9) * h
        for i in range(9):
            for j in range(9):
                He[i, j] = (E(xd + I[i] + I[j]) - E(xd + I[i] - I[j]) - E(xd - I[i] + I[j]) + E(xd - I[i] - I[j])) / (4 * h * h)
        w, V = np.linalg.eigh(He)
        Hp = (V * np.maximum(w, 0)) @ V.T
        assert np.allclose(H, H.T, atol=1e-6) and np.linalg.eigvalsh(H).min() > -1e-5
        assert np.linalg.eigvalsh(Hp - H).min() > -2e-4 * np.abs(Hp).max()          # a sub-sum of the positive modes
        if np.linalg.det(xd.reshape(3, 3).T) > 0.05:
            assert np.abs(H - Hp).max() < 2e-4 * np.abs(Hp).max()
            exact += 1
    assert exact > 15

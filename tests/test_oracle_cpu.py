"""Pins the oracle's leaf math (oracle/oracle_math.h, oracle_elastic.h) against the reference.

Known-answer vectors: tests/golden/leaf_math.npz, produced by tests/golden/make_leaf_golden.py from
the reference's own headers compiled for the host (oracle/_ref/libref_host.so).  Integer results and
fp32 payloads of the contact predicates are compared BIT FOR BIT (both sides are built with
-ffp-contract=off, so the operation order is the only degree of freedom); the elastic energies are
compared within 2e-4 of the per-case scale (the oracle evaluates them in x space, the vectors are in
F space, so the chain rule is applied here in float64).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP = C.POINTER(C.c_float)


def fp(a):
    return a.ctypes.data_as(FP)


def _build_csrc_shim():
    """The backend's own math headers (csrc/math_*.cuh) compiled for the host, same C signatures."""
    out = os.path.join(REPO, "tests", "native", "build", "libcsrc_math.so")
    src = os.path.join(REPO, "tests", "native", "csrc_shim.cpp")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(REPO, "ppf-contact-solver_b200", "csrc", f) for f in os.listdir(os.path.join(REPO, "ppf-contact-solver_b200", "csrc")) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-shared", "-x", "c++", src, "-o", out], check=True)
    return out


@pytest.fixture(scope="module", params=["oracle", "backend-math"])
def orc(built, request):
    """Both implementations of the leaf math run the same known-answer checks: the oracle's plain-C
    restatement and the CUDA backend's headers compiled for the host."""
    path = os.path.join(REPO, "oracle", "build", "liboracle_math.so") if request.param == "oracle" else _build_csrc_shim()
    lib = C.CDLL(path)
    lib.orc_ccd.restype = C.c_float
    if request.param == "oracle":
        lib.orc_edge_triangle_intersect.restype = C.c_int
    lib.which = request.param
    return lib


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(REPO, "tests", "golden", "leaf_math.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(a, b):
    """bitwise equality, NaN payloads treated as equal when both are NaN"""
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


def test_distance_coefficients_bit_exact(orc, gold):
    pts = gold["dist_pts"]
    m = len(pts)
    pt = np.zeros((m, 3), np.float32); ptu = np.zeros((m, 3), np.float32); pe = np.zeros((m, 2), np.float32); ee = np.zeros((m, 4), np.float32)
    for i in range(m):
        q = [fp(pts[i, k]) for k in range(4)]
        orc.orc_point_triangle_coeff(q[0], q[1], q[2], q[3], fp(pt[i]))
        orc.orc_point_triangle_coeff_unclassified(q[0], q[1], q[2], q[3], fp(ptu[i]))
        orc.orc_point_edge_coeff(q[0], q[1], q[2], fp(pe[i]))
        orc.orc_edge_edge_coeff(q[0], q[1], q[2], q[3], fp(ee[i]))
    for name, mine in (("pt_coeff", pt), ("pt_coeff_unclassified", ptu), ("pe_coeff", pe), ("ee_coeff", ee)):
        ok = same_bits(mine, gold[name])
        assert ok.all(), "%s: %d of %d cases differ from the reference" % (name, int((~ok.all(axis=1)).sum()), m)


def test_edge_triangle_intersection_exact(orc, gold):
    if orc.which != "oracle":
        pytest.skip("the backend's intersection test is device code (covered by the gpu step tests)")
    p5 = gold["isect_pts"]
    mine = np.array([orc.orc_edge_triangle_intersect(*[fp(p5[i, k]) for k in range(5)]) for i in range(len(p5))], np.int32)
    assert (mine == gold["isect"]).all()
    assert 0 < mine.sum() < len(mine)  # both outcomes are exercised


@pytest.mark.parametrize("name,kind", [("pt", 0), ("ee", 1), ("pe", 2), ("pp", 3)])
def test_accd_bit_exact(orc, gold, name, kind):
    x0 = gold["ccd_%s_x0" % name]; x1 = gold["ccd_%s_x1" % name]
    toi = np.zeros(len(x0), np.float32); ov = np.zeros(len(x0), np.uint32)
    for i in range(len(x0)):
        o = C.c_uint(0)
        toi[i] = orc.orc_ccd(kind, fp(x0[i]), fp(x1[i]), C.c_float(0.0), C.c_float(1e-3), C.c_float(1.25), C.c_float(1e-2), C.byref(o))
        ov[i] = o.value
    ok = same_bits(toi, gold["ccd_%s_toi" % name])
    assert ok.all(), "%d of %d TOIs differ" % (int((~ok).sum()), len(ok))
    assert (ov == gold["ccd_%s_overlap" % name]).all()
    t = gold["ccd_%s_toi" % name]
    assert (t < 1.25).any() and (t == np.float32(1.25)).any()  # hits and misses are both exercised


def test_cubic_barrier_bit_exact(orc, gold):
    g = gold["cubic_g"]
    mine = np.zeros((len(g), 2), np.float32)
    for i in range(len(g)):
        orc.orc_cubic(C.c_float(g[i]), C.c_float(1e-3), C.c_float(2e-4), fp(mine[i]))
    assert same_bits(mine, gold["cubic"]).all()


def test_baraff_witkin_matches_reference_energy(orc, gold):
    F = gold["bw_F"].astype(np.float64)
    rng = np.random.default_rng(3)
    worst = 0.0
    for i in range(len(F)):
        # a triangle whose deformation gradient is F[i]: rest triangle with inverse rest matrix IR (2x2, col-major)
        ir = np.array([[1.3, -0.2], [0.1, 0.9]])                 # IR (row, col)
        Fm = F[i].reshape(2, 3).T                                  # 3x2, column-major input
        D = Fm @ np.linalg.inv(ir)                                 # [x1-x0, x2-x0]
        x0 = rng.uniform(-1, 1, 3)
        x = np.stack([x0, x0 + D[:, 0], x0 + D[:, 1]]).astype(np.float32)
        ir32 = np.array([ir[0, 0], ir[1, 0], ir[0, 1], ir[1, 1]], np.float32)
        g = np.zeros(9, np.float32); H = np.zeros(81, np.float32)
        orc.orc_bw_face(fp(x), fp(ir32), C.c_float(1.3), C.c_float(0.7), C.c_float(2.0), fp(g), fp(H))
        # chain rule in float64 from the exact float32 inputs the oracle saw
        D32 = np.stack([x[1] - x[0], x[2] - x[0]], axis=1).astype(np.float64)
        ir64 = np.array([[ir32[0], ir32[2]], [ir32[1], ir32[3]]], np.float64)
        # dF/dx: F = D * IR, D columns = x1-x0, x2-x0 ; vec(F) column-major (3x2)
        dD = np.zeros((6, 9))
        for k in range(3):
            dD[k, 3 + k] = 1; dD[k, k] -= 1
            dD[3 + k, 6 + k] = 1; dD[3 + k, k] -= 1
        dF = np.kron(ir64.T, np.eye(3)) @ dD
        # the vectors were produced at F[i]; the float32 round trip moves F slightly, so compare with a tolerance
        gref = 2.0 * dF.T @ gold["bw_dedF"][i].astype(np.float64)
        Href = 2.0 * dF.T @ gold["bw_d2edF2"][i].astype(np.float64).reshape(6, 6).T @ dF
        assert np.abs(D32 @ ir64 - Fm).max() < 1e-5
        sg = max(np.abs(gref).max(), 1e-3); sH = np.abs(Href).max()
        worst = max(worst, np.abs(g - gref).max() / sg, np.abs(H.reshape(9, 9).T - Href).max() / sH)
    assert worst < 2e-4, worst


def test_hinge_matches_reference(orc, gold):
    hx = gold["hinge_x"]; rest = gold["hinge_rest"]
    worst_f = worst_H = 0.0
    for i in range(len(hx)):
        f = np.zeros(12, np.float32); H = np.zeros(144, np.float32)
        orc.orc_hinge(fp(hx[i]), C.c_float(rest[i]), fp(f), fp(H))
        fr = gold["hinge_f"][i]; Hr = gold["hinge_H"][i]
        worst_f = max(worst_f, np.abs(f - fr).max() / max(np.abs(fr).max(), 1e-6))
        worst_H = max(worst_H, np.abs(H - Hr).max() / max(np.abs(Hr).max(), 1e-6))
    assert worst_f < 1e-5 and worst_H < 1e-4, (worst_f, worst_H)


def test_friction_bit_exact(orc, gold):
    if "fric_fc" not in gold.files:
        pytest.skip("leaf_math.npz predates the friction vectors")
    orc.orc_friction.restype = C.c_float
    fc, dx, n, mu = gold["fric_fc"], gold["fric_dx"], gold["fric_n"], gold["fric_mu"]
    g = np.zeros((len(fc), 3), np.float32); h = np.zeros((len(fc), 9), np.float32); lam = np.zeros(len(fc), np.float32)
    for i in range(len(fc)):
        lam[i] = orc.orc_friction(fp(fc[i]), fp(dx[i]), fp(n[i]), C.c_float(mu[i]), C.c_float(1e-5), fp(g[i]), fp(h[i]))
    assert same_bits(lam, gold["fric_lambda"]).all()
    assert same_bits(g, gold["fric_grad"]).all()
    assert same_bits(h, gold["fric_hess"]).all()
    kin = np.linalg.norm(dx - n * (n * dx).sum(axis=1, keepdims=True), axis=1) > 1e-5
    assert kin.any() and (~kin).any()          # static and kinetic branches are both exercised


def test_rod_hook_bit_exact(orc, gold):
    """Hookean rod edge (hook.hpp:20-37): gradient and 6x6 Hessian, bit for bit."""
    rx = gold["rod_x"]
    g = np.zeros((len(rx), 6), np.float32); h = np.zeros((len(rx), 36), np.float32)
    for i in range(len(rx)):
        orc.orc_rod_hook(fp(rx[i]), C.c_float(gold["rod_l0"][i]), C.c_float(gold["rod_weight"][i]), fp(g[i]), fp(h[i]))
    assert same_bits(g, gold["rod_grad"]).all()
    assert same_bits(h, gold["rod_hess"]).all()
    stretched = np.linalg.norm(rx[:, 1] - rx[:, 0], axis=1) > gold["rod_l0"]
    assert stretched.any() and (~stretched).any()     # both branches of the c0 / c1 clamps


def test_rod_bend_bit_exact(orc, gold):
    """Interior rod vertex bending (dihedral_angle.hpp:451-541): force and PSD-projected 9x9 Hessian, bit for bit,
    over random triples, nearly straight yarn segments and (near-)collinear ones that hit the guards."""
    sx = gold["strand_x"]
    f = np.zeros((len(sx), 9), np.float32); H = np.zeros((len(sx), 81), np.float32)
    for i in range(len(sx)):
        orc.orc_strand(fp(sx[i]), C.c_float(gold["strand_rest"][i]), fp(f[i]), fp(H[i]))
    assert (same_bits(f, gold["strand_f"]) | ((f == 0) & (gold["strand_f"] == 0))).all()
    assert (same_bits(H, gold["strand_H"]) | ((H == 0) & (gold["strand_H"] == 0))).all()
    assert (np.abs(gold["strand_H"]).max(axis=1) == 0).any()   # guarded (collinear) cases are present
    # the projected Hessian is symmetric PSD
    for i in range(0, len(sx), 37):
        A = H[i].reshape(9, 9).astype(np.float64)
        assert np.abs(A - A.T).max() <= 1e-5 * max(np.abs(A).max(), 1e-12)
        assert np.linalg.eigvalsh(0.5 * (A + A.T)).min() >= -1e-4 * max(np.abs(A).max(), 1e-12)


def test_rod_bend_force_is_the_energy_gradient(orc):
    """finite differences of E = 1/2 (theta - theta0)^2 in float64 against the fp32 force"""
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(200):
        x = rng.uniform(-1, 1, size=(3, 3))
        rest = rng.uniform(1.0, np.pi)

        def energy(y):
            e0, e1 = y[0] - y[1], y[2] - y[1]
            th = np.arctan2(np.linalg.norm(np.cross(e0, e1)), e0 @ e1)
            return 0.5 * (th - rest) ** 2
        num = np.zeros((3, 3))
        for a in range(3):
            for k in range(3):
                d = np.zeros((3, 3)); d[a, k] = 1e-6
                num[a, k] = (energy(x + d) - energy(x - d)) / 2e-6
        f = np.zeros(9, np.float32); H = np.zeros(81, np.float32)
        x32 = x.astype(np.float32)
        orc.orc_strand(fp(x32), C.c_float(rest), fp(f), fp(H))
        worst = max(worst, np.abs(f.reshape(3, 3) - num).max() / max(np.abs(num).max(), 1e-3))
    assert worst < 2e-3, worst
    orc.orc_rod_bend_coefficient.restype = C.c_float
    k = orc.orc_rod_bend_coefficient(C.c_float(2e-3), C.c_float(4e-3), C.c_float(2e-3), C.c_float(3e-3), C.c_float(5e-4))
    assert abs(k - 3e-3 * 5e-4 * 1e-4 / (2.5e-3) ** 2) < 1e-9


def test_rod_strain_limit_is_the_barrier_gradient_and_hessian(orc):
    """strainlimiting.cu:122-190: with a zero elastic matrix the stiffness is mass / (limit - g)^2; force / stiffness must
    be the gradient of the cubic barrier energy E(g) = 2 g^3 / (3 limit) of the stretch g = l / l0 - 1 (cubic.hpp:13-21
    at the gap limit - g), and the block matrix its Hessian (both PSD clamps are inactive for g > 0)."""
    rng = np.random.default_rng(11)
    orc.orc_rod_strain.restype = C.c_int
    worst_g = worst_h = 0.0
    n_on = 0
    for _ in range(300):
        x = rng.uniform(-1, 1, size=(2, 3))
        l = np.linalg.norm(x[1] - x[0])
        l0 = l / rng.uniform(0.97, 1.04)              # stretched by up to 4 % or compressed by up to 3 %
        limit, mass = 0.05, 0.3
        x32 = x.astype(np.float32); c32 = (x + rng.normal(0, 0.01, size=(2, 3))).astype(np.float32)
        g6 = np.zeros(6, np.float32); h36 = np.zeros(36, np.float32); hel = np.zeros(27, np.float32)
        on = orc.orc_rod_strain(fp(x32), fp(c32), C.c_float(l0), C.c_float(limit), C.c_float(mass), fp(hel), fp(g6), fp(h36))
        g = np.linalg.norm(x32[1].astype(np.float64) - x32[0]) / np.float32(l0) - 1.0
        assert bool(on) == (g > 0)
        if not on:
            continue
        n_on += 1
        kappa = mass / (limit - g) ** 2

        def grad(y):
            gg = np.linalg.norm(y[1] - y[0]) / np.float64(np.float32(l0)) - 1.0
            n = (y[1] - y[0]) / np.linalg.norm(y[1] - y[0])
            dE = 2.0 * gg * gg / limit
            return np.concatenate([-n, n]) * dE / np.float64(np.float32(l0))
        ga = grad(x32.astype(np.float64))
        worst_g = max(worst_g, np.abs(g6 / kappa - ga).max() / max(np.abs(ga).max(), 1e-9))
        num = np.zeros((6, 6))
        for k in range(6):
            d = np.zeros(6); d[k] = 1e-6
            num[:, k] = (grad(x32.astype(np.float64) + d.reshape(2, 3)) - grad(x32.astype(np.float64) - d.reshape(2, 3))) / 2e-6
        H = h36.reshape(6, 6).T.astype(np.float64) / kappa
        worst_h = max(worst_h, np.abs(H - num).max() / max(np.abs(num).max(), 1e-9))
    assert n_on > 50 and worst_g < 1e-3 and worst_h < 1e-3, (n_on, worst_g, worst_h)
    # elasticity-inclusive stiffness: w^T H_el w with w = (c0 - centre, c1 - centre) adds to mass / (limit - g)^2
    x32 = np.array([[0, 0, 0], [1.02, 0, 0]], np.float32); c32 = np.array([[0, 0, 0], [1, 0, 0]], np.float32)
    hel = np.zeros(27, np.float32); hel[0] = 3.0; hel[18] = 5.0; hel[9] = -1.0     # xx entries of (e0,e0), (e1,e1), (e0,e1)
    g0 = np.zeros(6, np.float32); g1 = np.zeros(6, np.float32); h = np.zeros(36, np.float32)
    orc.orc_rod_strain(fp(x32), fp(c32), C.c_float(1.0), C.c_float(0.05), C.c_float(0.3), fp(np.zeros(27, np.float32)), fp(g0), fp(h))
    orc.orc_rod_strain(fp(x32), fp(c32), C.c_float(1.0), C.c_float(0.05), C.c_float(0.3), fp(hel), fp(g1), fp(h))
    k0 = 0.3 / (0.05 - 0.02) ** 2
    wHw = 0.25 * 3.0 + 0.25 * 5.0 + 2 * (-0.5) * 0.5 * (-1.0)
    assert abs(g1[3] / g0[3] - (k0 + wHw) / k0) < 1e-4


def test_rod_strain_line_search(orc):
    orc.orc_rod_strain_toi.restype = C.c_float
    d0 = np.array([1.0, 0, 0], np.float32)
    for grow, limit in ((1.01, 0.02), (1.03, 0.02), (1.10, 0.02), (0.9, 0.02)):
        d1 = (d0 * grow).astype(np.float32)
        t = orc.orc_rod_strain_toi(fp(d0), fp(d1), C.c_float(1.0), C.c_float(limit), C.c_float(1.25))
        if 1.25 * (grow - 1.0) < limit * 0.999:
            assert t == np.float32(1.25)
        else:
            assert abs(t * (grow - 1.0) - limit) < 1e-5 and t * (grow - 1.0) <= limit * (1 + 1e-5)   # fp32 bisection


def _strain_case(rng, stretch):
    """a rest triangle in the plane, deformed by a random rotation times a stretch"""
    rest = np.array([[0.0, 0.0], [1.0, 0.1], [0.2, 0.9]]) * 0.05
    Dm = np.stack([rest[1] - rest[0], rest[2] - rest[0]], axis=1)
    ir = np.linalg.inv(Dm)                                    # (row, col)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    Fm = q[:, :2] @ np.diag(stretch) @ np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
    x0 = rng.uniform(-1, 1, 3)
    x = np.stack([x0, x0 + Fm @ Dm[:, 0], x0 + Fm @ Dm[:, 1]]).astype(np.float32)
    ir32 = np.array([ir[0, 0], ir[1, 0], ir[0, 1], ir[1, 1]], np.float32)   # column-major like the ABI
    return x, ir32


def test_isotropic_shell_model_tables_bit_exact(orc, gold):
    """ARAP / StVK / SNHk derivative tables w.r.t. the two singular values (arap.hpp:36-49, stvk.hpp:66-72 + detsqr.hpp:22-33,
    snhk.hpp:45-56), bit for bit against the reference's headers"""
    a, model = gold["iso_a"], gold["iso_model"]
    t = np.zeros((len(a), 5), np.float32)
    for i in range(len(a)):
        orc.orc_iso_table2(C.c_int(int(model[i])), fp(a[i]), C.c_float(gold["iso_mu"][i]), C.c_float(gold["iso_lmd"][i]), fp(t[i]))
    assert same_bits(t, gold["iso_table"]).all()
    assert set(model.tolist()) == {0, 1, 3}


def _iso_energy(model, S, mu, lmd):
    a, b = S
    if model == 0:
        return mu * ((a - 1) ** 2 + (b - 1) ** 2) + 0.5 * lmd * (a + b - 2) ** 2
    if model == 1:
        return mu * ((a - 1) ** 4 + (b - 1) ** 4) / 4.0 + 0.5 * lmd * (a * b - 1) ** 2
    return 0.5 * mu * (a * a + b * b - 2.0) - mu * (a * b - 1.0) + 0.5 * lmd * (a * b - 1.0) ** 2


@pytest.mark.parametrize("model", [0, 1, 3])
def test_isotropic_shell_face_is_the_energy_gradient_and_hessian(orc, model):
    """embed_face_force_hessian for the SVD-based models (energy.cu:337-352): force = mass x dE/dx of the model's energy of
    the singular values; where every eigen-mode is convex (a stretched face) the projected Hessian is the exact one."""
    rng = np.random.default_rng(21 + model)
    mu, lmd, mass = 3.0, 2.0, 0.7
    worst_g = worst_h = 0.0
    for _ in range(60):
        x, ir32 = _strain_case(rng, (rng.uniform(1.05, 1.3), rng.uniform(1.05, 1.3)))
        ir = np.array([[ir32[0], ir32[2]], [ir32[1], ir32[3]]], np.float64)

        def energy(y):
            Ds = np.stack([y[1] - y[0], y[2] - y[0]], axis=1)
            S = np.linalg.svd(Ds @ ir, compute_uv=False)
            return mass * _iso_energy(model, S, mu, lmd)
        x64 = x.astype(np.float64)
        num = np.zeros(9); numH = np.zeros((9, 9))
        h = 1e-5

        def grad(y):
            g = np.zeros(9)
            for k in range(9):
                d = np.zeros(9); d[k] = h
                g[k] = (energy(y + d.reshape(3, 3)) - energy(y - d.reshape(3, 3))) / (2 * h)
            return g
        num = grad(x64)
        for k in range(9):
            d = np.zeros(9); d[k] = 1e-4
            numH[:, k] = (grad(x64 + d.reshape(3, 3)) - grad(x64 - d.reshape(3, 3))) / 2e-4
        g9 = np.zeros(9, np.float32); H81 = np.zeros(81, np.float32)
        orc.orc_iso_face(fp(x), fp(ir32), C.c_int(model), C.c_float(mu), C.c_float(lmd), C.c_float(mass), C.c_float(1e-2), fp(g9), fp(H81))
        worst_g = max(worst_g, np.abs(g9 - num).max() / max(np.abs(num).max(), 1e-9))
        H = H81.reshape(9, 9).T.astype(np.float64)
        assert np.linalg.eigvalsh(0.5 * (H + H.T)).min() > -1e-3 * np.abs(H).max()       # PSD by construction
        if np.linalg.eigvalsh(0.5 * (numH + numH.T)).min() > -1e-6 * np.abs(numH).max():     # exact Hessian already PSD: nothing was projected away
            worst_h = max(worst_h, np.abs(H - numH).max() / np.abs(numH).max())
    assert worst_g < 2e-3, worst_g
    assert worst_h < 2e-2, worst_h


def test_strain_limit_svd_and_gradient(orc):
    """Jacobi SVD reconstructs F; the strain-limit force is the gradient of the barrier energy on the
    singular values (checked against float64 finite differences of numpy's SVD); the Hessian is symmetric PSD."""
    orc.orc_strain_toi.restype = C.c_float
    rng = np.random.default_rng(11)
    for trial in range(60):
        F = rng.normal(size=(3, 2)).astype(np.float32)
        U = np.zeros(6, np.float32); S = np.zeros(2, np.float32); Vt = np.zeros(4, np.float32)
        orc.orc_svd32(fp(F), fp(U), fp(S), fp(Vt))
        rec = U.reshape(3, 2) @ np.diag(S) @ Vt.reshape(2, 2)
        assert np.abs(rec - F).max() < 5e-6 and S[0] <= S[1]
        assert np.allclose(S, np.linalg.svd(F.astype(np.float64), compute_uv=False)[::-1], atol=5e-6)
    strainlimit = 0.05
    worst = 0.0
    for trial in range(40):
        x, ir = _strain_case(rng, [1.0 + rng.uniform(0.002, 0.03), 1.0 + rng.uniform(-0.02, 0.04)])
        Hel = np.zeros(81, np.float32)
        f = np.zeros(9, np.float32); H = np.zeros(81, np.float32)
        on = orc.orc_strain_face(fp(x), fp(x), fp(ir), C.c_float(strainlimit), C.c_float(strainlimit), C.c_float(2.0), fp(Hel), C.c_float(1e-2), fp(f), fp(H))
        assert on == 1
        irm = np.array([[ir[0], ir[2]], [ir[1], ir[3]]], np.float64)

        def energy(xx):
            Fm = np.stack([xx[1] - xx[0], xx[2] - xx[0]], axis=1) @ irm
            sv = np.linalg.svd(Fm, compute_uv=False) - 1.0
            e = 0.0
            for g in sv:
                if g > 0:
                    y = strainlimit - g                       # cubic barrier energy (cubic.hpp): -2/(3 ghat) (y - ghat)^3 for y < ghat
                    e += -2.0 * (y - strainlimit) ** 3 / (3.0 * strainlimit)
            return e

        x64 = x.astype(np.float64)
        Fm = np.stack([x64[1] - x64[0], x64[2] - x64[0]], axis=1) @ irm
        smax = np.linalg.svd(Fm, compute_uv=False).max() - 1.0
        stiffness = 2.0 / (strainlimit - smax) ** 2            # elastic blocks are zero here: mass / d^2
        gnum = np.zeros((3, 3))
        h = 1e-6
        for a in range(3):
            for k in range(3):
                xp = x64.copy(); xm = x64.copy(); xp[a, k] += h; xm[a, k] -= h
                gnum[a, k] = (energy(xp) - energy(xm)) / (2 * h)
        scale = max(np.abs(gnum).max() * stiffness, 1e-9)
        worst = max(worst, np.abs(f.reshape(3, 3) - stiffness * gnum).max() / scale)
        Hm = np.zeros((9, 9))
        for a in range(3):
            for b in range(3):
                Hm[3 * a:3 * a + 3, 3 * b:3 * b + 3] = H[(a * 3 + b) * 9:(a * 3 + b) * 9 + 9].reshape(3, 3).T
        assert np.abs(Hm - Hm.T).max() <= 1e-4 * np.abs(Hm).max()
        assert np.linalg.eigvalsh(0.5 * (Hm + Hm.T)).min() >= -1e-4 * np.abs(Hm).max()
    assert worst < 2e-3, worst


def test_strain_limit_line_search(orc):
    orc.orc_strain_toi.restype = C.c_float
    rng = np.random.default_rng(12)
    for trial in range(40):
        xa, ir = _strain_case(rng, [1.0, 1.0])
        grow = 1.0 + rng.uniform(0.0, 0.2)
        xb = (xa[0] + (xa - xa[0]) * grow).astype(np.float32)   # uniform scaling about vertex 0: strain = t (grow - 1)
        limit = 0.05
        t = orc.orc_strain_toi(fp(xa), fp(xb), fp(ir), C.c_float(limit), C.c_float(1.25))
        if 1.25 * (grow - 1.0) < limit * 0.999:
            assert t == np.float32(1.25)
        elif 1.25 * (grow - 1.0) > limit * 1.001:
            assert abs(t * (grow - 1.0) - limit) < 2e-4 and t * (grow - 1.0) <= limit * (1 + 1e-5)


def test_oracle_agrees_with_live_reference_headers(orc, gold):
    """When oracle/_ref/libref_host.so is present (build container), cross-check fresh random cases too."""
    path = os.path.join(REPO, "oracle", "_ref", "libref_host.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_host.so not built (no /root/reference here)")
    ref = C.CDLL(path)
    rng = np.random.default_rng(77)
    pts = rng.uniform(-1, 1, size=(2000, 4, 3)).astype(np.float32)
    a = np.zeros(4, np.float32); b = np.zeros(4, np.float32)
    for i in range(len(pts)):
        q = [fp(pts[i, k]) for k in range(4)]
        orc.orc_edge_edge_coeff(q[0], q[1], q[2], q[3], fp(a)); ref.ref_edge_edge_coeff(q[0], q[1], q[2], q[3], fp(b))
        assert same_bits(a, b).all()
        orc.orc_point_triangle_coeff(q[0], q[1], q[2], q[3], fp(a)); ref.ref_point_triangle_coeff(q[0], q[1], q[2], q[3], fp(b))
        assert same_bits(a[:3], b[:3]).all()

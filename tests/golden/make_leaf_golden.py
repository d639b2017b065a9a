#!/usr/bin/env python3
"""make_leaf_golden.py -- known-answer vectors for the per-element math of the hot path.

Runs in the build container (needs oracle/_ref/libref_host.so, i.e. the reference's OWN headers
distance.hpp / accd.hpp / intersect_core.hpp / cubic.hpp / baraffwitkin.hpp / dihedral_angle.hpp / hook.hpp
compiled for the host by oracle/Makefile.ref with -ffp-contract=off) and writes
tests/golden/leaf_math.npz: seeded inputs plus the reference's outputs.  The CPU tests check the
oracle restatement against these vectors bit for bit (float32 payloads compared as uint32).
"""
import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ref = C.CDLL(os.path.join(REPO, "oracle", "_ref", "libref_host.so"))
FP = C.POINTER(C.c_float)


def fp(a):
    return a.ctypes.data_as(FP)


def lattice(rng, n, k):
    """n cases of k points on a coarse lattice: hits every degenerate branch (parallel edges,
    projections on vertices / edges, coincident points)."""
    return (rng.integers(0, 3, size=(n, k, 3)).astype(np.float32) * np.float32(0.5)).astype(np.float32)


def main():
    rng = np.random.default_rng(20240917)
    out = {}
    # ---- distance coefficients ----
    n = 3000
    pts = np.concatenate([rng.uniform(-1, 1, size=(n, 4, 3)).astype(np.float32), lattice(rng, 1500, 4),
                          (rng.uniform(-1, 1, size=(500, 4, 3)) * 1e-3 + 0.3).astype(np.float32)]).copy()
    m = len(pts)
    pt = np.zeros((m, 3), np.float32); ptu = np.zeros((m, 3), np.float32); pe = np.zeros((m, 2), np.float32); ee = np.zeros((m, 4), np.float32)
    it = np.zeros(m, np.int32)
    for i in range(m):
        q = [fp(pts[i, k]) for k in range(4)]
        ref.ref_point_triangle_coeff(q[0], q[1], q[2], q[3], fp(pt[i]))
        ref.ref_point_triangle_coeff_unclassified(q[0], q[1], q[2], q[3], fp(ptu[i]))
        ref.ref_point_edge_coeff(q[0], q[1], q[2], fp(pe[i]))
        ref.ref_edge_edge_coeff(q[0], q[1], q[2], q[3], fp(ee[i]))
    out.update(dist_pts=pts, pt_coeff=pt, pt_coeff_unclassified=ptu, pe_coeff=pe, ee_coeff=ee)
    # ---- edge / triangle intersection (5 points) ----
    p5 = np.concatenate([rng.uniform(-1, 1, size=(3000, 5, 3)).astype(np.float32), lattice(rng, 1500, 5)]).copy()
    ref.ref_edge_triangle_intersect.restype = C.c_int
    it = np.array([ref.ref_edge_triangle_intersect(*[fp(p5[i, k]) for k in range(5)]) for i in range(len(p5))], np.int32)
    out.update(isect_pts=p5, isect=it)
    # ---- ACCD ----
    for name, kind, fn in (("pt", 0, ref.ref_point_triangle_ccd), ("ee", 1, ref.ref_edge_edge_ccd), ("pe", 2, ref.ref_point_edge_ccd), ("pp", 3, ref.ref_point_point_ccd)):
        fn.restype = C.c_float
        nn = 1500
        x0 = rng.uniform(-0.05, 0.05, size=(nn, 4, 3)).astype(np.float32)
        x1 = (x0 + rng.normal(0, 0.04, size=(nn, 4, 3))).astype(np.float32)
        x0l = (lattice(rng, 400, 4) * np.float32(0.1)).astype(np.float32)
        x1l = (lattice(rng, 400, 4) * np.float32(0.1)).astype(np.float32)
        x0 = np.concatenate([x0, x0l]).copy(); x1 = np.concatenate([x1, x1l]).copy()
        toi = np.zeros(len(x0), np.float32); ov = np.zeros(len(x0), np.uint32)
        for i in range(len(x0)):
            o = C.c_uint(0)
            toi[i] = fn(fp(x0[i]), fp(x1[i]), C.c_float(0.0), C.c_float(1e-3), C.c_float(1.25), C.c_float(1e-2), C.byref(o))
            ov[i] = o.value
        out.update({"ccd_%s_x0" % name: x0, "ccd_%s_x1" % name: x1, "ccd_%s_toi" % name: toi, "ccd_%s_overlap" % name: ov})
    # ---- cubic barrier ----
    g = np.concatenate([rng.uniform(0, 2e-3, 2000), [0.0, 1e-3, 2e-3]]).astype(np.float32)
    cub = np.zeros((len(g), 3), np.float32)
    for i in range(len(g)):
        ref.ref_cubic(C.c_float(g[i]), C.c_float(1e-3), C.c_float(2e-4), fp(cub[i]))
    out.update(cubic_g=g, cubic=cub[:, 1:].copy())
    # ---- Baraff-Witkin in F space and hinge ----
    F = (np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), (500, 1)) + rng.normal(0, 0.2, size=(500, 6))).astype(np.float32)
    dedF = np.zeros((500, 6), np.float32); d2 = np.zeros((500, 36), np.float32)
    for i in range(500):
        ref.ref_baraffwitkin(fp(F[i]), C.c_float(1.3), C.c_float(0.7), fp(dedF[i]), fp(d2[i]))
    out.update(bw_F=F, bw_dedF=dedF, bw_d2edF2=d2)
    hx = rng.uniform(-1, 1, size=(800, 4, 3)).astype(np.float32)
    hf = np.zeros((800, 12), np.float32); hH = np.zeros((800, 144), np.float32); ha = np.zeros(800, np.float32)
    rest = rng.uniform(-0.5, 0.5, 800).astype(np.float32)
    for i in range(800):
        a = C.c_float(0)
        ref.ref_hinge(fp(hx[i]), C.c_float(rest[i]), fp(hf[i]), fp(hH[i]), C.byref(a))
        ha[i] = a.value
    out.update(hinge_x=hx, hinge_rest=rest, hinge_f=hf, hinge_H=hH)
    # ---- friction (appended last so that the earlier arrays keep their random streams) ----
    nfr = 1500
    fc = rng.normal(0, 1.0, size=(nfr, 3)).astype(np.float32)
    fdx = (rng.normal(0, 1.0, size=(nfr, 3)) * 10.0 ** rng.uniform(-7, -2, size=(nfr, 1))).astype(np.float32)
    fn = rng.normal(size=(nfr, 3)); fn = (fn / np.linalg.norm(fn, axis=1, keepdims=True)).astype(np.float32)
    fmu = rng.choice([0.0, 0.1, 0.5], size=nfr).astype(np.float32)
    fg = np.zeros((nfr, 3), np.float32); fh = np.zeros((nfr, 9), np.float32); fl = np.zeros(nfr, np.float32)
    for i in range(nfr):
        lam = C.c_float(0)
        ref.ref_friction(fp(fc[i]), fp(fdx[i]), fp(fn[i]), C.c_float(fmu[i]), C.c_float(1e-5), fp(fg[i]), fp(fh[i]), C.byref(lam))
        fl[i] = lam.value
    out.update(fric_fc=fc, fric_dx=fdx, fric_n=fn, fric_mu=fmu, fric_grad=fg, fric_hess=fh, fric_lambda=fl)
    # ---- rods: Hookean stretch (hook.hpp) and interior-vertex bending (dihedral_angle.hpp strand_*) ----
    nr = 1200
    rx = rng.uniform(-1, 1, size=(nr, 2, 3)).astype(np.float32)
    rl0 = (np.linalg.norm(rx[:, 1] - rx[:, 0], axis=1) * rng.uniform(0.6, 1.4, nr)).astype(np.float32)
    rw = rng.uniform(0.1, 50.0, nr).astype(np.float32)
    rg = np.zeros((nr, 6), np.float32); rh = np.zeros((nr, 36), np.float32)
    for i in range(nr):
        ref.ref_hook(fp(rx[i]), C.c_float(rl0[i]), C.c_float(rw[i]), fp(rg[i]), fp(rh[i]))
    out.update(rod_x=rx, rod_l0=rl0, rod_weight=rw, rod_grad=rg, rod_hess=rh)
    ns = 1500
    sx = rng.uniform(-1, 1, size=(ns, 3, 3)).astype(np.float32)
    # nearly straight yarn segments (the regime of the woven scenes) and exactly collinear ones (guards)
    seg = rng.uniform(2e-3, 4e-3, size=(500, 1)).astype(np.float32)
    dirs = rng.normal(size=(500, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    mid = rng.uniform(-1, 1, size=(500, 3))
    bend = rng.normal(0, 0.05, size=(500, 3)) * seg
    near = np.stack([mid - dirs * seg, mid + bend, mid + dirs * seg * rng.uniform(0.7, 1.3, size=(500, 1))], axis=1).astype(np.float32)
    col = np.stack([mid[:50] - dirs[:50], mid[:50], mid[:50] + 2 * dirs[:50]], axis=1).astype(np.float32)
    col[:, 1] = (0.5 * (col[:, 0] + col[:, 2])).astype(np.float32)
    sx = np.concatenate([sx, near, col]).copy()
    srest = np.concatenate([rng.uniform(0.5, np.pi, ns), np.full(500, np.pi), np.full(50, np.pi)]).astype(np.float32)
    sf = np.zeros((len(sx), 9), np.float32); sh = np.zeros((len(sx), 81), np.float32)
    for i in range(len(sx)):
        ref.ref_strand(fp(sx[i]), C.c_float(srest[i]), fp(sf[i]), fp(sh[i]))
    out.update(strand_x=sx, strand_rest=srest, strand_f=sf, strand_H=sh)
    # ---- isotropic shell models ARAP / StVK / SNHk: derivative tables w.r.t. the singular values (arap.hpp, stvk.hpp, snhk.hpp) ----
    nt = 600
    ta = rng.uniform(0.5, 1.6, size=(nt, 2)).astype(np.float32)
    tmodel = rng.choice([0, 1, 3], size=nt).astype(np.int32)
    tmu = rng.uniform(0.5, 500.0, nt).astype(np.float32); tl = rng.uniform(0.5, 800.0, nt).astype(np.float32)
    tt = np.zeros((nt, 5), np.float32)
    for i in range(nt):
        ref.ref_iso_table2(C.c_int(int(tmodel[i])), fp(ta[i]), C.c_float(tmu[i]), C.c_float(tl[i]), fp(tt[i]))
    out.update(iso_a=ta, iso_model=tmodel, iso_mu=tmu, iso_lmd=tl, iso_table=tt)
    # ---- quadratic and logarithmic barriers (barrier/quadratic.hpp, barrier/logarithm.hpp) ----
    bg = np.concatenate([rng.uniform(1e-5, 2.4e-3, 1500), [2e-4, 1.2e-3, 1.3e-3]]).astype(np.float32)
    bk = rng.integers(1, 3, size=len(bg)).astype(np.int32)
    bo = np.zeros((len(bg), 2), np.float32)
    for i in range(len(bg)):
        ref.ref_barrier(C.c_int(int(bk[i])), C.c_float(bg[i]), C.c_float(1e-3), C.c_float(2e-4), fp(bo[i]))
    out.update(barrier_g=bg, barrier_kind=bk, barrier_out=bo)
    path = os.path.join(REPO, "tests", "golden", "leaf_math.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

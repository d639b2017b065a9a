// csrc_shim.cpp -- exports the BACKEND's per-element math (ppf-contact-solver_b200/csrc/math_*.cuh,
// compiled for the host in the tests only, -ffp-contract=off) under the same C signatures as
// oracle/oracle_shim.c, so tests/test_oracle_cpu.py runs one set of known-answer checks over both.
#include "../../ppf-contact-solver_b200/csrc/math_accd.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_cmesh.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_collider.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_contact.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_damping.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_distance.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_elastic.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_hinge.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_inflate.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_plasticity.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_rod.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_stitch.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_strain.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_torque.cuh"
using namespace ppf;
#define SHIM extern "C" __attribute__((visibility("default")))
static V3 v(const float *p) { return mk(p[0], p[1], p[2]); }

SHIM void orc_point_triangle_coeff(const float *p, const float *t0, const float *t1, const float *t2, float *o) { const C3 c = point_triangle_coeff(v(p), v(t0), v(t1), v(t2)); o[0] = c.a; o[1] = c.b; o[2] = c.c; }
SHIM void orc_point_triangle_coeff_unclassified(const float *p, const float *t0, const float *t1, const float *t2, float *o) { const C3 c = point_triangle_coeff_unclassified(v(p), v(t0), v(t1), v(t2)); o[0] = c.a; o[1] = c.b; o[2] = c.c; }
SHIM void orc_point_edge_coeff(const float *p, const float *e0, const float *e1, float *o) { const C2 c = point_edge_coeff(v(p), v(e0), v(e1)); o[0] = c.a; o[1] = c.b; }
SHIM void orc_edge_edge_coeff(const float *a0, const float *a1, const float *b0, const float *b1, float *o) { const C4 c = edge_edge_coeff(v(a0), v(a1), v(b0), v(b1)); o[0] = c.a; o[1] = c.b; o[2] = c.c; o[3] = c.d; }
SHIM float orc_ccd(int kind, const float *x0, const float *x1, float offset, float ghat, float max_t, float ccd_eps, unsigned *overlap) {
    CcdOverlap ov{0u, 0.0f, 0.0f};
    float r;
    if (kind == 0) r = point_triangle_ccd(v(x0), v(x1), v(x0 + 3), v(x0 + 6), v(x0 + 9), v(x1 + 3), v(x1 + 6), v(x1 + 9), offset, ghat, max_t, ccd_eps, ov);
    else if (kind == 1) r = edge_edge_ccd(v(x0), v(x0 + 3), v(x0 + 6), v(x0 + 9), v(x1), v(x1 + 3), v(x1 + 6), v(x1 + 9), offset, ghat, max_t, ccd_eps, ov);
    else if (kind == 2) r = point_edge_ccd(v(x0), v(x1), v(x0 + 3), v(x0 + 6), v(x1 + 3), v(x1 + 6), offset, ghat, max_t, ccd_eps, ov);
    else r = point_point_ccd(v(x0), v(x1), v(x0 + 3), v(x1 + 3), offset, ghat, max_t, ccd_eps, ov);
    if (overlap) *overlap = ov.flag;
    return r;
}
SHIM void orc_cubic(float g, float ghat, float offset, float *out2) { out2[0] = cubic_gradient(g, ghat, offset); out2[1] = cubic_curvature(g, ghat, offset); }
SHIM void orc_bw_face(const float *x9, const float *ir4, float mu, float lmd, float mass, float *g9, float *H81) {
    FaceElastic fe;
    bw_face(v(x9), v(x9 + 3), v(x9 + 6), ir4, mu, lmd, mass, fe);
    for (int a = 0; a < 3; ++a) { g9[3 * a] = fe.f[a].x; g9[3 * a + 1] = fe.f[a].y; g9[3 * a + 2] = fe.f[a].z; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) H81[(3 * b + c) * 9 + 3 * a + r] = fe.H[a][b].m[c * 3 + r];
}
SHIM void orc_hinge(const float *x12, float rest_angle, float *f12, float *H144) {
    static const int map[4] = {2, 1, 0, 3};
    V3 X[4];
    for (int a = 0; a < 4; ++a) X[a] = v(x12 + 3 * map[a]);
    HingeElastic he;
    hinge_force_hessian(X, rest_angle, he);
    for (int a = 0; a < 4; ++a) { f12[3 * a] = he.f[a].x; f12[3 * a + 1] = he.f[a].y; f12[3 * a + 2] = he.f[a].z; }
    for (int i = 0; i < 144; ++i) H144[i] = he.H[i];
}

SHIM int orc_strain_face(const float *xe9, const float *xc9, const float *ir4, float strainlimit, float limit, float mass, const float *Hel81,
                         float eig_eps, float *f9, float *H81) {
    V3 xe[3], xc[3];
    M3 Hel[3][3];
    for (int a = 0; a < 3; ++a) { xe[a] = v(xe9 + 3 * a); xc[a] = v(xc9 + 3 * a); }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int e = 0; e < 9; ++e) Hel[a][b].m[e] = Hel81[(a * 3 + b) * 9 + e];
    StrainFace sf;
    if (!strain_face(xe, xc, ir4, strainlimit, limit, mass, Hel, eig_eps, sf)) return 0;
    for (int a = 0; a < 3; ++a) { f9[3 * a] = sf.f[a].x; f9[3 * a + 1] = sf.f[a].y; f9[3 * a + 2] = sf.f[a].z; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int e = 0; e < 9; ++e) H81[(a * 3 + b) * 9 + e] = sf.H[a][b].m[e];
    return 1;
}
SHIM float orc_strain_toi(const float *xa9, const float *xb9, const float *ir4, float limit, float max_t) {
    V3 xa[3], xb[3];
    for (int a = 0; a < 3; ++a) { xa[a] = v(xa9 + 3 * a); xb[a] = v(xb9 + 3 * a); }
    return strain_face_toi(xa, xb, ir4, limit, max_t);
}
SHIM void orc_svd32(const float *F6, float *U6, float *S2, float *Vt4) {
    float F[3][2];
    Svd32 s;
    for (int r = 0; r < 3; ++r) { F[r][0] = F6[2 * r]; F[r][1] = F6[2 * r + 1]; }
    jacobi_svd32(F, s);
    for (int r = 0; r < 3; ++r) { U6[2 * r] = s.U[r][0]; U6[2 * r + 1] = s.U[r][1]; }
    S2[0] = s.S[0]; S2[1] = s.S[1];
    Vt4[0] = s.Vt[0][0]; Vt4[1] = s.Vt[0][1]; Vt4[2] = s.Vt[1][0]; Vt4[3] = s.Vt[1][1];
}

SHIM float orc_friction(const float *fc, const float *dx, const float *n, float mu, float min_dx, float *grad, float *hess) {
    const FrictionOut o = friction_eval(v(fc), v(dx), v(n), mu, min_dx);
    grad[0] = o.grad.x; grad[1] = o.grad.y; grad[2] = o.grad.z;
    for (int e = 0; e < 9; ++e) hess[e] = o.hess.m[e];
    return o.lambda;
}

// rods: same C signatures as the oracle shim (gradient 3x2 / Hessian 6x6 column-major; force 3x3 / Hessian 9x9)
SHIM void orc_rod_hook(const float *x6, float l0, float weight, float *grad6, float *hess36) {
    RodElastic re;
    rod_hook(v(x6), v(x6 + 3), l0, weight, re);
    for (int a = 0; a < 2; ++a) { grad6[3 * a] = re.f[a].x; grad6[3 * a + 1] = re.f[a].y; grad6[3 * a + 2] = re.f[a].z; }
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
        hess36[(3 * b + c) * 6 + 3 * a + r] = (a == b ? 1.0f : -1.0f) * re.K.m[c * 3 + r];
}
SHIM void orc_strand(const float *x9, float rest_angle, float *force9, float *hess81) {
    StrandElastic se;
    strand_force_hessian(v(x9), v(x9 + 3), v(x9 + 6), rest_angle, se);
    for (int a = 0; a < 3; ++a) { force9[3 * a] = se.f[a].x; force9[3 * a + 1] = se.f[a].y; force9[3 * a + 2] = se.f[a].z; }
    for (int i = 0; i < 81; ++i) hess81[i] = se.H[i];
}
SHIM float orc_rod_bend_coefficient(float b0, float b1, float l0, float l1, float mass) { return rod_bend_coefficient(b0, b1, l0, l1, mass); }
// rod strain limiting: Hel27 = blocks (e0,e0), (e0,e1), (e1,e1) column-major; outputs like orc_rod_hook; returns 0 when not stretched
SHIM int orc_rod_strain(const float *x6, const float *c6, float l0, float limit, float mass, const float *Hel27, float *grad6, float *hess36) {
    M3 h[3];
    for (int q = 0; q < 3; ++q) for (int e = 0; e < 9; ++e) h[q].m[e] = Hel27[9 * q + e];
    RodElastic re;
    if (!rod_strain(v(x6), v(x6 + 3), v(c6), v(c6 + 3), l0, limit, mass, h[0], h[1], h[2], re)) return 0;
    for (int a = 0; a < 2; ++a) { grad6[3 * a] = re.f[a].x; grad6[3 * a + 1] = re.f[a].y; grad6[3 * a + 2] = re.f[a].z; }
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
        hess36[(3 * b + c) * 6 + 3 * a + r] = (a == b ? 1.0f : -1.0f) * re.K.m[c * 3 + r];
    return 1;
}
SHIM float orc_rod_strain_toi(const float *d0, const float *d1, float l0, float limit, float max_t) { return rod_strain_toi(v(d0), v(d1), l0, limit, max_t); }

// analytic colliders: sphere / floor records as 10 floats each (layout documented in oracle/oracle_math.h)
static void unpack_colliders(const float *sp, unsigned ns, const float *fl, unsigned nf, PpfSphere *S, PpfFloor *F) {
    for (unsigned j = 0; j < ns; ++j) {
        const float *s = sp + 10 * j;
        PpfSphere q{};
        q.center.v[0] = s[0]; q.center.v[1] = s[1]; q.center.v[2] = s[2];
        q.ghat = s[3]; q.friction = s[4]; q.radius = s[5]; q.thickness = s[6]; q.bowl = s[7] != 0.0f; q.reverse = s[8] != 0.0f; q.kinematic = s[9] != 0.0f;
        S[j] = q;
    }
    for (unsigned j = 0; j < nf; ++j) {
        const float *s = fl + 10 * j;
        PpfFloor q{};
        q.ground.v[0] = s[0]; q.ground.v[1] = s[1]; q.ground.v[2] = s[2];
        q.ghat = s[3]; q.friction = s[4]; q.thickness = s[5]; q.up.v[0] = s[6]; q.up.v[1] = s[7]; q.up.v[2] = s[8]; q.kinematic = s[9] != 0.0f;
        F[j] = q;
    }
}
SHIM unsigned orc_collider_vertex(const float *x, const float *dx, float mass, float vfriction, const float *Hii, const float *spheres, unsigned ns,
                                  const float *floors, unsigned nf, int friction_mode, float friction_eps, float constraint_tol, float *f, float *H) {
    PpfSphere S[16]; PpfFloor F[16];
    unpack_colliders(spheres, ns, floors, nf, S, F);
    M3 h;
    for (int e = 0; e < 9; ++e) h.m[e] = Hii[e];
    ColliderOut o;
    collider_vertex(v(x), v(dx), mass, vfriction, h, S, ns, F, nf, friction_mode, friction_eps, constraint_tol, o);
    f[0] = o.f.x; f[1] = o.f.y; f[2] = o.f.z;
    for (int e = 0; e < 9; ++e) H[e] = o.H.m[e];
    return o.count;
}
SHIM float orc_collider_toi(const float *y0, const float *y1, int pinned, const float *spheres, unsigned ns, const float *floors, unsigned nf, float max_t, int *infeasible) {
    PpfSphere S[16]; PpfFloor F[16];
    unpack_colliders(spheres, ns, floors, nf, S, F);
    bool bad = false;
    const float t = collider_toi(v(y0), v(y1), pinned != 0, S, ns, F, nf, max_t, &bad);
    *infeasible = bad ? 1 : 0;
    return t;
}

// isotropic shell models (ARAP 0 / StVK 1 / SNHk 3): derivative table and one face (force 3x3, Hessian 9x9 column-major)
SHIM void orc_iso_table2(int model, const float *a2, float mu, float lmd, float *out5) {
    float deda[2], d2[3];
    iso_table2(model, a2, mu, lmd, deda, d2);
    out5[0] = deda[0]; out5[1] = deda[1]; out5[2] = d2[0]; out5[3] = d2[1]; out5[4] = d2[2];
}
SHIM void orc_iso_face(const float *x9, const float *ir4, int model, float mu, float lmd, float mass, float eig_eps, float *g9, float *H81) {
    StrainFace sf;
    iso_face(v(x9), v(x9 + 3), v(x9 + 6), ir4, model, mu, lmd, mass, eig_eps, sf);
    for (int a = 0; a < 3; ++a) { g9[3 * a] = sf.f[a].x; g9[3 * a + 1] = sf.f[a].y; g9[3 * a + 2] = sf.f[a].z; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) H81[(3 * b + c) * 9 + 3 * a + r] = sf.H[a][b].m[c * 3 + r];
}

// Rayleigh damping of one element (n = 2, 3 or 4 vertices): K (3n)^2 column-major in/out, f 3n in/out; lagged != 0 uses K_lag
SHIM void orc_damping(int n, float *K, const float *K_lag, const float *u, float beta, float dt, float *f, int lagged) {
    V3 uu[4], ff[4];
    for (int a = 0; a < n; ++a) { uu[a] = v(u + 3 * a); ff[a] = v(f + 3 * a); }
    if (n == 2) { if (lagged) stiffness_damping_lagged<2>(K, K_lag, uu, beta, dt, ff); else stiffness_damping<2>(K, uu, beta, dt, ff); }
    else if (n == 3) { if (lagged) stiffness_damping_lagged<3>(K, K_lag, uu, beta, dt, ff); else stiffness_damping<3>(K, uu, beta, dt, ff); }
    else { if (lagged) stiffness_damping_lagged<4>(K, K_lag, uu, beta, dt, ff); else stiffness_damping<4>(K, uu, beta, dt, ff); }
    for (int a = 0; a < n; ++a) { f[3 * a] = ff[a].x; f[3 * a + 1] = ff[a].y; f[3 * a + 2] = ff[a].z; }
}

// one stitch: x18 = six slot positions, ghat6 / offset6 / w6; outputs gradient 18, Hessian 18x18 column-major
SHIM void orc_stitch(const float *x18, const float *ghat6, const float *offset6, const float *w6, float stiffness, float length_factor, float *f18, float *H324) {
    V3 x[6];
    for (int a = 0; a < 6; ++a) x[a] = v(x18 + 3 * a);
    StitchOut o;
    stitch_force_hessian(x, ghat6, offset6, w6, stiffness, length_factor, o);
    for (int a = 0; a < 6; ++a) { f18[3 * a] = o.f[a].x; f18[3 * a + 1] = o.f[a].y; f18[3 * a + 2] = o.f[a].z; }
    for (int i = 0; i < 324; ++i) H324[i] = o.H[i];
}

// plastic creep: face (ir4 in/out, returns changed), hinge (x12 in the ORIGINAL hinge order h0..h3), rod bend (x9 = j, i, k)
SHIM int orc_face_plasticity(const float *x9, float *ir4, float threshold, float rate, float dt) {
    return face_plasticity(v(x9), v(x9 + 3), v(x9 + 6), ir4, threshold, plastic_alpha(rate, dt)) ? 1 : 0;
}
SHIM float orc_hinge_plasticity(const float *x12, float rest, float threshold, float rate, float dt) {
    static const int map[4] = {2, 1, 0, 3};
    V3 X[4];
    for (int a = 0; a < 4; ++a) X[a] = v(x12 + 3 * map[a]);
    return hinge_plasticity(X, rest, threshold, plastic_alpha(rate, dt));
}
SHIM float orc_rod_bend_plasticity(const float *x9, float rest, float threshold, float rate, float dt) {
    return rod_bend_plasticity(v(x9), v(x9 + 3), v(x9 + 6), rest, threshold, plastic_alpha(rate, dt));
}

SHIM void orc_barrier(int kind, float g, float ghat, float offset, float *out2) { out2[0] = barrier_gradient(g, ghat, offset, kind); out2[1] = barrier_curvature(g, ghat, offset, kind); }

// static collision mesh (math_cmesh.cuh); same argument layout as the oracle's shim
SHIM void orc_cmesh(int kind, const float *dyn, const float *sta, const float *par, const float *B, float *out16) {
    CmeshOut o;
    o.f = mk(0.f, 0.f, 0.f); o.H = m3_zero(); o.c[0] = o.c[1] = o.c[2] = 0.0f;
    bool hit;
    if (kind == 0) {
        M3 Hii; for (int e = 0; e < 9; ++e) Hii.m[e] = B[e];
        hit = cmesh_vertex_face(v(dyn), v(dyn + 3), v(sta), v(sta + 3), v(sta + 6), par[0], par[1], par[2], par[4], Hii, par[3], o);
    } else if (kind == 1) {
        V3 X[3], Xs[3]; M3 Bl[3][3];
        for (int a = 0; a < 3; ++a) { X[a] = v(dyn + 3 * a); Xs[a] = v(dyn + 9 + 3 * a); for (int b = 0; b < 3; ++b) for (int e = 0; e < 9; ++e) Bl[a][b].m[e] = B[(a * 3 + b) * 9 + e]; }
        hit = cmesh_face_vertex(v(sta), X, Xs, par[0], par[1], par[2], par + 4, Bl, par[3], o);
    } else {
        V3 P[2], Ps[2]; M3 Bl[2][2];
        for (int a = 0; a < 2; ++a) { P[a] = v(dyn + 3 * a); Ps[a] = v(dyn + 6 + 3 * a); for (int b = 0; b < 2; ++b) for (int e = 0; e < 9; ++e) Bl[a][b].m[e] = B[(a * 2 + b) * 9 + e]; }
        hit = cmesh_edge_edge(P, Ps, v(sta), v(sta + 3), par[0], par[1], par[2], par + 4, Bl, par[3], o);
    }
    out16[0] = hit ? 1.0f : 0.0f;
    out16[1] = o.f.x; out16[2] = o.f.y; out16[3] = o.f.z;
    for (int k = 0; k < 9; ++k) out16[4 + k] = o.H.m[k];
    for (int k = 0; k < 3; ++k) out16[13 + k] = o.c[k];
}

// torque pins (math_torque.cuh); same layout as the oracle's shim
SHIM void orc_torque(const float *x, const float *mass, unsigned n, unsigned axis_component, unsigned hint, float magnitude, float *out) {
    PpfTorqueVertex tv[256];
    for (unsigned k = 0; k < n; ++k) { tv[k].index = k; tv[k].group_id = 0; tv[k].magnitude = magnitude; }
    PpfTorqueGroup g{axis_component, 0u, n, hint};
    const TorqueGroupResult r = torque_group(g, tv, [&](unsigned i) { return v(x + 3 * i); }, [&](unsigned i) { return mass[i]; });
    out[0] = r.center.x; out[1] = r.center.y; out[2] = r.center.z; out[3] = r.axis.x; out[4] = r.axis.y; out[5] = r.axis.z; out[6] = r.inv_r_perp_sq_sum;
    for (unsigned k = 0; k < n; ++k) {
        V3 f; M3 H;
        torque_vertex(v(x + 3 * k), magnitude, r, f, H);
        float *o = out + 7 + 12 * k;
        o[0] = f.x; o[1] = f.y; o[2] = f.z;
        for (int e = 0; e < 9; ++e) o[3 + e] = H.m[e];
    }
}

// inflation pressure (math_inflate.cuh)
SHIM void orc_inflate(float pressure, const float *x9, float *g9, float *H81) {
    const V3 v0 = v(x9), v1 = v(x9 + 3), v2 = v(x9 + 6);
    V3 g[3];
    inflate_gradient(pressure, v0, v1 - v0, v2 - v0, g);
    for (int a = 0; a < 3; ++a) { g9[3 * a] = g[a].x; g9[3 * a + 1] = g[a].y; g9[3 * a + 2] = g[a].z; }
    inflate_hessian(pressure, v0, v1, v2, H81);
}

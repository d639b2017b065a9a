/* oracle_shim.c -- TEST INFRASTRUCTURE: exports the oracle's leaf math (oracle_math.h,
 * oracle_elastic.h) as plain C symbols so tests/test_oracle_cpu.py can pin it, bit for bit,
 * against the reference's own headers compiled for the host (oracle/_ref/libref_host.so) and
 * against the committed known-answer vectors in tests/golden/leaf_math.npz. */
#include "oracle_elastic.h"
#include "oracle_math.h"
#define SHIM __attribute__((visibility("default")))

SHIM void orc_point_triangle_coeff(const float *p, const float *t0, const float *t1, const float *t2, float *o) { o_point_triangle_coeff(p, t0, t1, t2, o); }
SHIM void orc_point_triangle_coeff_unclassified(const float *p, const float *t0, const float *t1, const float *t2, float *o) { o_point_triangle_coeff_unclassified(p, t0, t1, t2, o); }
SHIM void orc_point_edge_coeff(const float *p, const float *e0, const float *e1, float *o) { o_point_edge_coeff(p, e0, e1, o); }
SHIM void orc_edge_edge_coeff(const float *a0, const float *a1, const float *b0, const float *b1, float *o) { o_edge_edge_coeff(a0, a1, b0, b1, o); }
/* kind: 0 point-triangle, 1 edge-edge, 2 point-edge, 3 point-point; x0/x1: vertex-major positions at t=0 / t=1 */
SHIM float orc_ccd(int kind, const float *x0, const float *x1, float offset, float ghat, float max_t, float ccd_eps, unsigned *overlap) {
    static const int nc[4] = {4, 4, 3, 2}, split[4] = {1, 2, 1, 1};
    float a[4][3], b[4][3];
    for (int c = 0; c < nc[kind]; ++c) for (int k = 0; k < 3; ++k) { a[c][k] = x0[3 * c + k]; b[c][k] = x1[3 * c + k]; }
    unsigned ov = 0;
    const float r = o_accd(kind, nc[kind], split[kind], a, b, offset, ghat, max_t, ccd_eps, &ov);
    if (overlap) *overlap = ov;
    return r;
}
SHIM int orc_edge_triangle_intersect(const float *a0, const float *a1, const float *b0, const float *b1, const float *b2) { return o_edge_triangle_intersect(a0, a1, b0, b1, b2); }
SHIM void orc_cubic(float g, float ghat, float offset, float *out2) { out2[0] = o_cubic_gradient(g, ghat, offset); out2[1] = o_cubic_curvature(g, ghat, offset); }
SHIM void orc_bw_face(const float *x9, const float *ir4, float mu, float lmd, float mass, float *g9, float *H81) {
    float g[3][3];
    o_bw_face(x9, x9 + 3, x9 + 6, ir4, mu, lmd, mass, g, H81);
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) g9[3 * a + k] = g[a][k];
}
/* x12: hinge vertices in ORIGINAL order h0..h3; outputs in the remapped order (h2,h1,h0,h3) like the reference scatters */
SHIM void orc_hinge(const float *x12, float rest_angle, float *f12, float *H144) {
    static const int map[4] = {2, 1, 0, 3};
    float X[4][3], f[4][3];
    for (int a = 0; a < 4; ++a) for (int k = 0; k < 3; ++k) X[a][k] = x12[3 * map[a] + k];
    o_hinge(X, rest_angle, f, H144);
    for (int a = 0; a < 4; ++a) for (int k = 0; k < 3; ++k) f12[3 * a + k] = f[a][k];
}

/* strain limiting of one face: xe9/xc9 evaluation / start-of-step positions, Hel81 = 3x3 blocks [a][b][9] (column-major);
 * outputs scaled by the stiffness; returns 0 when the face is not stretched.  svd6: U(3x2, column-major) S Vt(2x2 row-major) */
SHIM int orc_strain_face(const float *xe9, const float *xc9, const float *ir4, float strainlimit, float limit, float mass, const float *Hel81,
                         float eig_eps, float *f9, float *H81) {
    float xe[3][3], xc[3][3], Hel[3][3][9], f[3][3], H[3][3][9];
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) { xe[a][k] = xe9[3 * a + k]; xc[a][k] = xc9[3 * a + k]; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int e = 0; e < 9; ++e) Hel[a][b][e] = Hel81[(a * 3 + b) * 9 + e];
    const int on = o_strain_face(xe, xc, ir4, strainlimit, limit, mass, Hel, eig_eps, f, H);
    if (!on) return 0;
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) f9[3 * a + k] = f[a][k];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int e = 0; e < 9; ++e) H81[(a * 3 + b) * 9 + e] = H[a][b][e];
    return 1;
}
SHIM float orc_strain_toi(const float *xa9, const float *xb9, const float *ir4, float limit, float max_t) {
    float xa[3][3], xb[3][3];
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) { xa[a][k] = xa9[3 * a + k]; xb[a][k] = xb9[3 * a + k]; }
    return o_strain_face_toi(xa, xb, ir4, limit, max_t);
}
SHIM void orc_svd32(const float *F6 /* row-major 3x2 */, float *U6, float *S2, float *Vt4) {
    float F[3][2];
    o_svd32 s;
    for (int r = 0; r < 3; ++r) { F[r][0] = F6[2 * r]; F[r][1] = F6[2 * r + 1]; }
    o_jacobi_svd32(F, &s);
    for (int r = 0; r < 3; ++r) { U6[2 * r] = s.U[r][0]; U6[2 * r + 1] = s.U[r][1]; }
    S2[0] = s.S[0]; S2[1] = s.S[1];
    Vt4[0] = s.Vt[0][0]; Vt4[1] = s.Vt[0][1]; Vt4[2] = s.Vt[1][0]; Vt4[3] = s.Vt[1][1];
}

SHIM float orc_friction(const float *fc, const float *dx, const float *n, float mu, float min_dx, float *grad, float *hess) { return o_friction(fc, dx, n, mu, min_dx, grad, hess); }

/* rods: gradient 3x2 / Hessian 6x6 (column-major) of one Hookean rod edge; force 3x3 / Hessian 9x9 of one bending vertex */
SHIM void orc_rod_hook(const float *x6, float l0, float weight, float *grad6, float *hess36) {
    float f[2][3], K[9];
    o_rod_hook(x6, x6 + 3, l0, weight, f, K);
    for (int a = 0; a < 2; ++a) for (int k = 0; k < 3; ++k) grad6[3 * a + k] = f[a][k];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
        hess36[(3 * b + c) * 6 + 3 * a + r] = (a == b ? 1.0f : -1.0f) * K[c * 3 + r];
}
SHIM void orc_strand(const float *x9, float rest_angle, float *force9, float *hess81) {
    float f[3][3];
    o_strand(x9, x9 + 3, x9 + 6, rest_angle, f, hess81);
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) force9[3 * a + k] = f[a][k];
}
SHIM float orc_rod_bend_coefficient(float b0, float b1, float l0, float l1, float mass) { return o_rod_bend_coefficient(b0, b1, l0, l1, mass); }
/* rod strain limiting: Hel27 = blocks (e0,e0), (e0,e1), (e1,e1) column-major; returns 0 when the edge is not stretched */
SHIM int orc_rod_strain(const float *x6, const float *c6, float l0, float limit, float mass, const float *Hel27, float *grad6, float *hess36) {
    float f[2][3], K[9];
    if (!o_rod_strain(x6, x6 + 3, c6, c6 + 3, l0, limit, mass, Hel27, Hel27 + 9, Hel27 + 18, f, K)) return 0;
    for (int a = 0; a < 2; ++a) for (int k = 0; k < 3; ++k) grad6[3 * a + k] = f[a][k];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
        hess36[(3 * b + c) * 6 + 3 * a + r] = (a == b ? 1.0f : -1.0f) * K[c * 3 + r];
    return 1;
}
SHIM float orc_rod_strain_toi(const float *d0, const float *d1, float l0, float limit, float max_t) { return o_rod_strain_toi(d0, d1, l0, limit, max_t); }

/* analytic colliders (10 floats per sphere / floor record, see oracle_math.h) */
SHIM unsigned orc_collider_vertex(const float *x, const float *dx, float mass, float vfriction, const float *Hii, const float *spheres, unsigned ns,
                                  const float *floors, unsigned nf, int friction_mode, float friction_eps, float constraint_tol, float *f, float *H) {
    return o_collider_vertex(x, dx, mass, vfriction, Hii, spheres, ns, floors, nf, friction_mode, friction_eps, constraint_tol, f, H);
}
SHIM float orc_collider_toi(const float *y0, const float *y1, int pinned, const float *spheres, unsigned ns, const float *floors, unsigned nf, float max_t, int *infeasible) {
    *infeasible = 0;
    return o_collider_toi(y0, y1, pinned, spheres, ns, floors, nf, max_t, infeasible);
}

/* isotropic shell models (ARAP 0 / StVK 1 / SNHk 3) */
SHIM void orc_iso_table2(int model, const float *a2, float mu, float lmd, float *out5) {
    float deda[2], d2[3];
    o_iso_table2(model, a2, mu, lmd, deda, d2);
    out5[0] = deda[0]; out5[1] = deda[1]; out5[2] = d2[0]; out5[3] = d2[1]; out5[4] = d2[2];
}
SHIM void orc_iso_face(const float *x9, const float *ir4, int model, float mu, float lmd, float mass, float eig_eps, float *g9, float *H81) {
    float f[3][3], H[3][3][9];
    o_iso_face(x9, x9 + 3, x9 + 6, ir4, model, mu, lmd, mass, eig_eps, f, H);
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) g9[3 * a + k] = f[a][k];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) H81[(3 * b + c) * 9 + 3 * a + r] = H[a][b][c * 3 + r];
}

/* Rayleigh damping of one element (n = 2, 3 or 4 vertices) */
SHIM void orc_damping(int n, float *K, const float *K_lag, const float *u, float beta, float dt, float *f, int lagged) {
    float uu[4][3], ff[4][3];
    for (int a = 0; a < n; ++a) for (int k = 0; k < 3; ++k) { uu[a][k] = u[3 * a + k]; ff[a][k] = f[3 * a + k]; }
    if (lagged) o_stiffness_damping_lagged(n, K, K_lag, uu, beta, dt, ff);
    else o_stiffness_damping(n, K, uu, beta, dt, ff);
    for (int a = 0; a < n; ++a) for (int k = 0; k < 3; ++k) f[3 * a + k] = ff[a][k];
}

/* one stitch (six slots) */
SHIM void orc_stitch(const float *x18, const float *ghat6, const float *offset6, const float *w6, float stiffness, float length_factor, float *f18, float *H324) {
    float x[6][3], f[6][3];
    for (int a = 0; a < 6; ++a) for (int k = 0; k < 3; ++k) x[a][k] = x18[3 * a + k];
    o_stitch(x, ghat6, offset6, w6, stiffness, length_factor, f, H324);
    for (int a = 0; a < 6; ++a) for (int k = 0; k < 3; ++k) f18[3 * a + k] = f[a][k];
}

/* plastic creep */
SHIM int orc_face_plasticity(const float *x9, float *ir4, float threshold, float rate, float dt) {
    return o_face_plasticity(x9, x9 + 3, x9 + 6, ir4, threshold, o_plastic_alpha(rate, dt));
}
SHIM float orc_hinge_plasticity(const float *x12, float rest, float threshold, float rate, float dt) {
    static const int map[4] = {2, 1, 0, 3};
    float X[4][3];
    for (int a = 0; a < 4; ++a) for (int k = 0; k < 3; ++k) X[a][k] = x12[3 * map[a] + k];
    return o_hinge_plasticity(X, rest, threshold, o_plastic_alpha(rate, dt));
}
SHIM float orc_rod_bend_plasticity(const float *x9, float rest, float threshold, float rate, float dt) {
    return o_rod_bend_plasticity(x9, x9 + 3, x9 + 6, rest, threshold, o_plastic_alpha(rate, dt));
}

/* barrier gradient / curvature of ParamSet.barrier (0 cubic, 1 quad, 2 log) */
SHIM void orc_barrier(int kind, float g, float ghat, float offset, float *out2) { out2[0] = o_barrier_gradient(g, ghat, offset, kind); out2[1] = o_barrier_curvature(g, ghat, offset, kind); }

/* static collision mesh: kind 0 dynamic vertex / static face, 1 dynamic face / static vertex, 2 dynamic edge / static edge.
 * dyn = the dynamic element's vertices (evaluation positions, then start-of-step positions), sta = the static element's,
 * par = offset, ghat, mu, friction_eps, then the vertex masses; B = the element's elastic blocks [n][n][9].
 * out = hit flag, f[3], H[9], c[3] */
SHIM void orc_cmesh(int kind, const float *dyn, const float *sta, const float *par, const float *B, float *out16) {
    float f[3] = {0, 0, 0}, H[9] = {0}, c[3] = {0, 0, 0};
    int hit;
    if (kind == 0) hit = o_cmesh_vertex_face(dyn, dyn + 3, sta, sta + 3, sta + 6, par[0], par[1], par[2], par[4], B, par[3], f, H, c);
    else if (kind == 1) hit = o_cmesh_face_vertex(sta, (const float(*)[3])dyn, (const float(*)[3])(dyn + 9), par[0], par[1], par[2], par + 4, B, par[3], f, H, c);
    else hit = o_cmesh_edge_edge((const float(*)[3])dyn, (const float(*)[3])(dyn + 6), sta, sta + 3, par[0], par[1], par[2], par + 4, B, par[3], f, H, c);
    out16[0] = (float)hit;
    for (int k = 0; k < 3; ++k) { out16[1 + k] = f[k]; out16[13 + k] = c[k]; }
    for (int k = 0; k < 9; ++k) out16[4 + k] = H[k];
}

/* torque pins: one group over all n vertices; out = group result (7), then per vertex f[3], H[9] */
SHIM void orc_torque(const float *x, const float *mass, unsigned n, unsigned axis_component, unsigned hint, float magnitude, float *out) {
    unsigned idx[256];
    for (unsigned k = 0; k < n; ++k) idx[k] = k;
    o_torque_group(axis_component, 0, n, hint, idx, x, mass, out);
    for (unsigned k = 0; k < n; ++k) o_torque_vertex(x + 3 * k, magnitude, out, out + 7 + 12 * k, out + 7 + 12 * k + 3);
}

/* inflation pressure of one face: gradient (9) and PSD-projected Hessian (81, column-major) */
SHIM void orc_inflate(float pressure, const float *x9, float *g9, float *H81) {
    float g[3][3];
    o_inflate_gradient(pressure, x9, x9 + 3, x9 + 6, g);
    for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) g9[3 * a + k] = g[a][k];
    o_inflate_hessian(pressure, x9, x9 + 3, x9 + 6, H81);
}

/* oracle_math.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C-style, float arrays) of the
 * per-element math of the reference's per-Newton-step path.  Nothing under ppf-contact-solver_b200/
 * includes, links or executes this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg use the oracle.
 *
 * Each function cites the reference file:line it follows (paths under
 * /root/reference/crates/ppf-cts-solver/src/cpp/).  fp32 throughout, same operation order as the
 * reference so that the contact predicates and the ACCD time of impact can be compared bit for bit
 * (tests/test_oracle_cpu.py pins every function here against the reference's own headers compiled
 * for the host, oracle/_ref/libref_host.so).
 */
#ifndef PPF_ORACLE_MATH_H
#define PPF_ORACLE_MATH_H
#include <float.h>
#include <math.h>
#include <string.h>

typedef float o3[3];

static inline void o_sub(const float *a, const float *b, float *r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
/* linalg/smat.hpp:330-335 -- accumulator starts at 0 */
static inline float o_dot(const float *a, const float *b) {
    float acc = 0.0f;
    for (int i = 0; i < 3; ++i) acc = acc + a[i] * b[i];
    return acc;
}
static inline void o_cross(const float *a, const float *b, float *r) { /* smat.hpp:337-343 */
    const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline float o_min_coeff(const float *m, int n) { float b = m[0]; for (int i = 1; i < n; ++i) if (m[i] < b) b = m[i]; return b; }
static inline float o_max_coeff(const float *m, int n) { float b = m[0]; for (int i = 1; i < n; ++i) if (b < m[i]) b = m[i]; return b; }
static inline int o_interior(const float *c, int n) { return o_max_coeff(c, n) < 1.0f && o_min_coeff(c, n) > 0.0f; }

/* contact/distance.hpp:15-27 */
static inline void o_point_edge_coeff(const float *p, const float *e0, const float *e1, float *c) {
    o3 r, d;
    o_sub(e1, e0, r);
    const float dd = o_dot(r, r);
    if (dd > 0.0f) {
        o_sub(p, e0, d);
        const float t = o_dot(r, d) / dd;
        c[0] = 1.0f - t; c[1] = t;
    } else { c[0] = 0.5f; c[1] = 0.5f; }
}
/* contact/distance.hpp:47-84 */
static inline void o_point_triangle_coeff(const float *p, const float *t0, const float *t1, const float *t2, float *c) {
    o3 r0, r1, d, nrm, tmp;
    o_sub(t1, t0, r0); o_sub(t2, t0, r1); o_sub(p, t0, d);
    o_cross(r0, r1, nrm);
    const float nn = o_dot(nrm, nrm);
    if (nn > 0.0f) {
        const float inv = 1.0f / nn;
        o_cross(d, r1, tmp);
        const float u = o_dot(tmp, nrm) * inv;
        o_cross(r0, d, tmp);
        const float v = o_dot(tmp, nrm) * inv;
        c[0] = 1.0f - u - v; c[1] = u; c[2] = v;
    } else {
        o3 e1, e2;
        float w[2];
        o_sub(t2, t1, e1); o_sub(t0, t2, e2);
        const float l0 = o_dot(r0, r0), l1 = o_dot(e1, e1), l2 = o_dot(e2, e2);
        if (l0 >= l1 && l0 >= l2) { o_point_edge_coeff(p, t0, t1, w); c[0] = w[0]; c[1] = w[1]; c[2] = 0.0f; }
        else if (l1 >= l2) { o_point_edge_coeff(p, t1, t2, w); c[0] = 0.0f; c[1] = w[0]; c[2] = w[1]; }
        else { o_point_edge_coeff(p, t2, t0, w); c[0] = w[1]; c[1] = 0.0f; c[2] = w[0]; }
    }
}
static inline float o_clamp_unit(float v) { return v > 0.0f ? (v < 1.0f ? v : 1.0f) : 0.0f; } /* distance.hpp:90-92 */
/* contact/distance.hpp:113-196 */
static inline void o_edge_edge_coeff(const float *a0, const float *a1, const float *b0, const float *b1, float *out) {
    o3 r0, r1, d, n, tmp;
    o_sub(a1, a0, r0); o_sub(b1, b0, r1); o_sub(a0, b0, d);
    const float a = o_dot(r0, r0), e = o_dot(r1, r1), b = o_dot(r0, r1), c = o_dot(r0, d), f = o_dot(r1, d);
    const float inv_a = a > 0.0f ? 1.0f / a : 0.0f, inv_e = e > 0.0f ? 1.0f / e : 0.0f;
    o_cross(r0, r1, n);
    const float nn = o_dot(n, n);
    const float inv_nn = nn > 0.0f ? 1.0f / nn : 0.0f;
    o_cross(r1, d, tmp);
    const float s_stat = o_clamp_unit(o_dot(n, tmp) * inv_nn);
    const float t_stat = o_clamp_unit((b * s_stat + f) * inv_e);
    const float cs[5] = {s_stat, 0.0f, 1.0f, o_clamp_unit(-c * inv_a), o_clamp_unit((b - c) * inv_a)};
    const float ct[5] = {t_stat, o_clamp_unit(f * inv_e), o_clamp_unit((f + b) * inv_e), 0.0f, 1.0f};
    float bs = 0.0f, bt = 0.0f, best = FLT_MAX;
    for (int i = 0; i < 5; ++i) {
        o3 v;
        for (int k = 0; k < 3; ++k) v[k] = d[k] + cs[i] * r0[k] - ct[i] * r1[k];
        const float dist = o_dot(v, v);
        if (dist < best) { best = dist; bs = cs[i]; bt = ct[i]; }
    }
    out[0] = 1.0f - bs; out[1] = bs; out[2] = 1.0f - bt; out[3] = bt;
}
/* contact/distance.hpp:224-283 */
static inline void o_point_triangle_coeff_unclassified(const float *p, const float *t0, const float *t1, const float *t2, float *out) {
    o3 r0, r1, d;
    o_sub(t1, t0, r0); o_sub(t2, t0, r1); o_sub(p, t0, d);
    float cand[4][3], w[2];
    int adm[4] = {0, 1, 1, 1};
    o_point_triangle_coeff(p, t0, t1, t2, cand[0]);
    adm[0] = o_min_coeff(cand[0], 3) >= 0.0f;
    o_point_edge_coeff(p, t0, t1, w); const float s01 = o_clamp_unit(w[1]);
    o_point_edge_coeff(p, t1, t2, w); const float s12 = o_clamp_unit(w[1]);
    o_point_edge_coeff(p, t2, t0, w); const float s20 = o_clamp_unit(w[1]);
    cand[1][0] = 1.0f - s01; cand[1][1] = s01; cand[1][2] = 0.0f;
    cand[2][0] = 0.0f; cand[2][1] = 1.0f - s12; cand[2][2] = s12;
    cand[3][0] = s20; cand[3][1] = 0.0f; cand[3][2] = 1.0f - s20;
    int bi = 1;
    float best = FLT_MAX;
    for (int i = 0; i < 4; ++i) {
        if (!adm[i]) continue;
        o3 v;
        for (int k = 0; k < 3; ++k) v[k] = d[k] - cand[i][1] * r0[k] - cand[i][2] * r1[k];
        const float dist = o_dot(v, v);
        if (dist < best) { best = dist; bi = i; }
    }
    memcpy(out, cand[bi], 3 * sizeof(float));
}
/* contact/distance.hpp:285-299 */
static inline void o_point_edge_coeff_unclassified(const float *p, const float *e0, const float *e1, float *c) {
    o_point_edge_coeff(p, e0, e1, c);
    if (c[0] >= 0.0f && c[0] <= 1.0f) return;
    if (c[0] > 1.0f) { c[0] = 1.0f; c[1] = 0.0f; }
    else { c[0] = 0.0f; c[1] = 1.0f; }
}

/* ---- ACCD (contact/accd.hpp) : columns x[c][3], nc columns, the first `split` belong to primitive A ---- */
enum { O_CCD_PT = 0, O_CCD_EE = 1, O_CCD_PE = 2, O_CCD_PP = 3 };

static inline float o_accd_sqdist(int kind, const float (*x)[3], float *dir) { /* accd.hpp:473-561 */
    if (kind == O_CCD_PT) {
        float c[3];
        o_point_triangle_coeff_unclassified(x[0], x[1], x[2], x[3], c);
        for (int k = 0; k < 3; ++k) dir[k] = c[0] * (x[1][k] - x[0][k]) + c[1] * (x[2][k] - x[0][k]) + c[2] * (x[3][k] - x[0][k]);
    } else if (kind == O_CCD_EE) {
        float c[4];
        o_edge_edge_coeff(x[0], x[1], x[2], x[3], c);
        for (int k = 0; k < 3; ++k) {
            const float a = c[0] * x[0][k] + c[1] * x[1][k], b = c[2] * x[2][k] + c[3] * x[3][k];
            dir[k] = b - a;
        }
    } else if (kind == O_CCD_PE) {
        float c[2];
        o_point_edge_coeff_unclassified(x[0], x[1], x[2], c);
        for (int k = 0; k < 3; ++k) dir[k] = (c[0] * x[1][k] + c[1] * x[2][k]) - x[0][k];
    } else {
        for (int k = 0; k < 3; ++k) dir[k] = x[1][k] - x[0][k];
    }
    return o_dot(dir, dir);
}

static inline float o_accd_directional(int nc, int split, const float (*x)[3], const float (*dx)[3], const float *w, float w_norm, float dip,
                                       float park, float max_t) { /* accd.hpp:242-282 */
    const float round_slack = 2.0f * FLT_EPSILON, coord_slack = 6.2e-7f;
    const float w_abs[3] = {fabsf(w[0]), fabsf(w[1]), fabsf(w[2])};
    const float dip_w = dip * w_norm, park_w = park * w_norm;
    float step = FLT_MAX;
    for (int i = 0; i < split; ++i)
        for (int j = split; j < nc; ++j) {
            o3 gap, rel, ga, ra;
            for (int k = 0; k < 3; ++k) { gap[k] = x[j][k] - x[i][k]; rel[k] = dx[i][k] - dx[j][k]; ga[k] = fabsf(gap[k]); ra[k] = fabsf(rel[k]); }
            const float err = round_slack * (o_dot(w_abs, ga) + max_t * o_dot(w_abs, ra) + park_w) + coord_slack * w_norm;
            const float projected = o_dot(w, gap) - err;
            if (!(projected - dip_w > 0.0f)) return 0.0f;
            const float rate = o_dot(w, rel);
            if (rate > 0.0f) step = fminf(step, fmaxf(0.0f, projected - park_w) / rate);
        }
    return step;
}

/* full driver: centerize 65-74, rescale 77-120, max_relative_u 142-151, ccd_helper 292-471, *_ccd 565-696.
 * x0/x1: nc points at t=0 and t=1.  *overlap is set when the reference would latch g_ccd_overlap. */
static inline float o_accd(int kind, int nc, int split, const float (*p0)[3], const float (*p1)[3], float offset, float ghat, float max_t,
                           float ccd_eps, unsigned *overlap) {
    float x0[4][3], dx[4][3], x[4][3];
    for (int c = 0; c < nc; ++c)
        for (int k = 0; k < 3; ++k) { x0[c][k] = p0[c][k]; dx[c][k] = p1[c][k] - p0[c][k]; }
    for (int pass = 0; pass < 2; ++pass) { /* centerize x0 then dx */
        float(*m)[3] = pass ? dx : x0;
        float mov[3] = {0.0f, 0.0f, 0.0f};
        const float sc = (float)(1.0 / nc);
        for (int c = 0; c < nc; ++c)
            for (int k = 0; k < 3; ++k) mov[k] = mov[k] + sc * m[c][k];
        for (int c = 0; c < nc; ++c)
            for (int k = 0; k < 3; ++k) m[c][k] = m[c][k] - mov[k];
    }
    float m0 = fabsf(x0[0][0]), m1 = fabsf(x0[0][0] + max_t * dx[0][0]);
    for (int c = 0; c < nc; ++c)
        for (int k = 0; k < 3; ++k) {
            const float a = fabsf(x0[c][k]), b = fabsf(x0[c][k] + max_t * dx[c][k]);
            if (m0 < a) m0 = a;
            if (m1 < b) m1 = b;
        }
    float scale = 0.99f / fmaxf(m0, m1);
    if (!(scale > 0.0f) || !isfinite(scale)) { *overlap = 1u; return 0.0f; }
    float scaled = 1.0f;
    for (;;) {
        const float s = 8.0f;
        if (scale > s) {
            for (int c = 0; c < nc; ++c) for (int k = 0; k < 3; ++k) { x0[c][k] *= s; dx[c][k] *= s; }
            scaled *= s; scale /= s;
        } else {
            for (int c = 0; c < nc; ++c) for (int k = 0; k < 3; ++k) { x0[c][k] *= scale; dx[c][k] *= scale; }
            scaled *= scale;
            break;
        }
    }
    float max_u = 0.0f;
    for (int i = 0; i < split; ++i)
        for (int j = split; j < nc; ++j) {
            o3 du;
            o_sub(dx[i], dx[j], du);
            const float sq = o_dot(du, du);
            if (max_u < sq) max_u = sq;
        }
    const float u_max = sqrtf(max_u);
    if (!u_max) return max_t;
    const float off = scaled * offset, eps = scaled * ccd_eps, floor_gap = scaled * (1e-2f * ghat);
    /* ccd_helper */
    float w[3], lower_t = 0.0f, toi = 0.0f;
    memcpy(x, x0, sizeof(x));
    float d2 = o_accd_sqdist(kind, x0, w);
    if (!(d2 > off * off)) { *overlap = 1u; return 0.0f; }
    const float clearance = sqrtf(d2) - off;
    const float floor_clear = fmaxf(2.0f * eps, floor_gap);
    const float park = off + fminf(floor_clear, fmaxf(clearance - 0.5f * eps, 0.5f * clearance));
    const float overshoot = fmaxf(0.0f, (park - off) - 0.5f * floor_clear);
    const float dip = park - overshoot;
    const float park_sq = fminf(park * park, nextafterf(d2, 0.0f));
    const float inv_u = 1.0f / u_max;
    unsigned probes = 0;
    for (;;) {
        if (!(d2 > park_sq)) break;
        lower_t = toi;
        const float dist = sqrtf(d2);
        float step = (dist - dip) * inv_u;
        if (toi + step <= max_t) step = fmaxf(step, o_accd_directional(nc, split, x, dx, w, dist * (1.0f + 2.0f * FLT_EPSILON), dip, park, max_t));
        toi += step;
        if (toi > max_t) return max_t;
        else if (++probes > 4096u) return lower_t;
        for (int c = 0; c < nc; ++c) for (int k = 0; k < 3; ++k) x[c][k] = x0[c][k] + toi * dx[c][k];
        d2 = o_accd_sqdist(kind, x, w);
    }
    float upper_t = toi, window = upper_t - lower_t;
    for (;;) {
        float tmp[3];
        toi = 0.5f * (upper_t + lower_t);
        for (int c = 0; c < nc; ++c) for (int k = 0; k < 3; ++k) x[c][k] = x0[c][k] + toi * dx[c][k];
        d2 = o_accd_sqdist(kind, x, tmp);
        if (d2 > park_sq) lower_t = toi; else upper_t = toi;
        const float nw = upper_t - lower_t;
        if (nw == window) break;
        window = nw;
    }
    if (!(lower_t > 0.0f)) { *overlap = 1u; return 0.0f; }
    return lower_t;
}

/* ---- cubic barrier (barrier/cubic.hpp:13-41) ---- */
static inline float o_cubic_gradient(float g, float ghat, float offset) { g -= offset; const float y = g - ghat; return y < 0.0f ? -2.0f * y * y / ghat : 0.0f; }
static inline float o_cubic_curvature(float g, float ghat, float offset) { g -= offset; return (g - ghat < 0.0f) ? 4.0f * (1.0f - g / ghat) : 0.0f; }

/* ---- edge-triangle pierce test (contact/intersect_core.hpp:21-79 through contact.cu:2745-2759) ---- */
static inline int o_edge_triangle_intersect(const float *a0, const float *a1, const float *b0, const float *b1, const float *b2) {
    o3 d1, d2, A0, A1, n;
    o_sub(b1, b0, d1); o_sub(b2, b0, d2); o_sub(a0, b0, A0); o_sub(a1, b0, A1);
    o_cross(d1, d2, n);
    const float s1 = A0[0] * n[0] + A0[1] * n[1] + A0[2] * n[2], s2 = A1[0] * n[0] + A1[1] * n[1] + A1[2] * n[2];
    if (s1 * s2 < 0.0f) {
        const float t = s1 / (s1 - s2);
        o3 r;
        for (int k = 0; k < 3; ++k) r[k] = (1.0f - t) * A0[k] + t * A1[k];
        const float a00 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], a01 = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
        const float a11 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
        const float q0 = d1[0] * r[0] + d1[1] * r[1] + d1[2] * r[2], q1 = d2[0] * r[0] + d2[1] * r[1] + d2[2] * r[2];
        const float det = a00 * a11 - a01 * a01;
        if (det == 0.0f) return 0;
        const float w0 = (a11 * q0 - a01 * q1) / det, w1 = (a00 * q1 - a01 * q0) / det, w2 = 1.0f - w0 - w1;
        float wmin = w0 < w1 ? w0 : w1; wmin = wmin < w2 ? wmin : w2;
        float wmax = w0 > w1 ? w0 : w1; wmax = wmax > w2 ? wmax : w2;
        return wmin >= 0.0f && wmax <= 1.0f;
    }
    return 0;
}


/* the other two barriers of ParamSet.barrier (barrier/quadratic.hpp:13-41, barrier/logarithm.hpp:14-44); kind 0 cubic, 1 quad, 2 log */
static inline float o_barrier_gradient(float g, float ghat, float offset, int kind) {
    if (kind == 0) return o_cubic_gradient(g, ghat, offset);
    g -= offset;
    if (kind == 1) { const float y = ghat - g; return y > 0.0f ? -2.0f * y : 0.0f; }
    if (g <= 0.0f) return -INFINITY;
    if (g >= ghat) return 0.0f;
    return (ghat - g) * (2.0f * g * logf(g / ghat) + g - ghat) / g;
}
static inline float o_barrier_curvature(float g, float ghat, float offset, int kind) {
    if (kind == 0) return o_cubic_curvature(g, ghat, offset);
    g -= offset;
    if (kind == 1) return (ghat - g > 0.0f) ? 2.0f : 0.0f;
    if (g <= 0.0f) return INFINITY;
    if (g >= ghat) return 0.0f;
    return -2.0f * logf(g / ghat) + ghat * (ghat + 2.0f * g) / (g * g) - 3.0f;
}

/* ---- lagged Coulomb friction (energy/model/friction.hpp:19-82): gradient and Hessian to ADD to the contact's ---- */
static inline float o_friction(const float *force_contact, const float *dx, const float *normal, float mu, float min_dx, float *grad, float *hess /* column-major 3x3 */) {
    const float contact = -o_dot(normal, force_contact), nd = o_dot(normal, dx);
    float u[3];
    for (int k = 0; k < 3; ++k) u[k] = dx[k] - normal[k] * nd;
    const float un = sqrtf(o_dot(u, u));
    const float lambda = mu > 0.0f ? mu * contact / (min_dx > un ? min_dx : un) : 0.0f;
    for (int k = 0; k < 3; ++k) grad[k] = lambda * u[k];
    if (un > min_dx) { /* kinetic: lambda * w w^T with w = n x slip direction */
        float s[3], w[3];
        for (int k = 0; k < 3; ++k) s[k] = u[k] / un;
        o_cross(normal, s, w);
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) hess[c * 3 + r] = lambda * (w[r] * w[c]);
    } else            /* static: lambda * (I - n n^T) */
        for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) hess[c * 3 + r] = lambda * ((r == c ? 1.0f : 0.0f) - normal[r] * normal[c]);
    return lambda;
}

/* ---- analytic colliders: invisible spheres / bowls and walls (TEST INFRASTRUCTURE restatement) ------------------
 * push barrier energy/model/push.hpp:15-50; embed_vertex_constraint_force_hessian contact/contact.cu:1154-1383;
 * vertex_constraint_line_search contact/contact.cu:2318-2470.  Sphere / floor records are passed as plain floats:
 *   sphere: center[3], ghat, friction, radius, thickness, bowl, reverse, kinematic   (10 floats)
 *   floor : ground[3], ghat, friction, thickness, up[3], kinematic                     (10 floats) */
static inline void o_push(const float *e, const float *n, float eps, float scale, float *f, float *H) { /* f += scale grad, H += scale hess */
    const float d = o_dot(e, n);
    if (!(d < 0.0f)) return;
    const float g = -(d * d) / eps, c = -2.0f * d / eps;
    for (int k = 0; k < 3; ++k) f[k] += scale * (g * n[k]);
    for (int cc = 0; cc < 3; ++cc) for (int r = 0; r < 3; ++r) H[cc * 3 + r] += scale * ((c * n[r]) * n[cc]);
}
static inline float o_combine_friction(float a, float b, int mode) { return mode == 1 ? (a > b ? a : b) : (mode == 2 ? 0.5f * (a + b) : (a < b ? a : b)); }
static inline float o_nHn(const float *H, const float *n) {
    float y[3];
    for (int r = 0; r < 3; ++r) y[r] = H[r] * n[0] + H[3 + r] * n[1] + H[6 + r] * n[2];
    return o_dot(n, y);
}
static inline unsigned o_collider_vertex(const float *x, const float *dx, float mass, float vfriction, const float *Hii, const float *spheres, unsigned n_sphere,
                                         const float *floors, unsigned n_floor, int friction_mode, float friction_eps, float constraint_tol, float *f, float *H) {
    unsigned count = 0;
    for (int k = 0; k < 3; ++k) f[k] = 0.0f;
    for (int k = 0; k < 9; ++k) H[k] = 0.0f;
    for (unsigned j = 0; j < n_sphere; ++j) {
        const float *s = spheres + 10 * j;
        const float ghat = s[3], friction = o_combine_friction(s[4], vfriction, friction_mode), sr = s[5], thickness = s[6];
        const int bowl = s[7] != 0.0f, kinematic = s[9] != 0.0f;
        int reverse = s[8] != 0.0f;
        float radius = sr;
        float center[3] = {s[0], s[1], s[2]}, ov[3];
        if (bowl && x[1] > center[1]) center[1] = x[1];
        o_sub(x, center, ov);
        const float d2 = o_dot(ov, ov);
        if (kinematic) {
            if (!d2) continue;
            const float dist = sqrtf(d2);
            float normal[3] = {ov[0] / dist, ov[1] / dist, ov[2] / dist}, e[3];
            const float eff = reverse ? (radius - ghat) : (radius + ghat);
            for (int k = 0; k < 3; ++k) e[k] = ov[k] - eff * normal[k];
            if (bowl) reverse = 1;
            const float stiff_k = mass / (ghat * ghat), r2 = eff * eff;
            const float depth = reverse ? (dist - sr) : (sr - dist);
            if (thickness > 0.0f && depth > thickness) continue;
            if (reverse && d2 > r2) { float nn[3] = {-normal[0], -normal[1], -normal[2]}; ++count; o_push(e, nn, ghat, stiff_k, f, H); }
            else if (!reverse && d2 < r2) { ++count; o_push(e, normal, ghat, stiff_k, f, H); }
        } else {
            if (reverse) radius -= ghat; else radius += ghat;
            const float r2 = radius * radius;
            if (!(reverse ? d2 > r2 : d2 < r2)) continue;
            const float dist = sqrtf(d2);
            const float depth = reverse ? (dist - sr) : (sr - dist);
            if (thickness > 0.0f && depth > thickness) continue;
            ++count;
            float normal[3] = {ov[0] / dist, ov[1] / dist, ov[2] / dist}, e[3];
            for (int k = 0; k < 3; ++k) e[k] = ov[k] - radius * normal[k];
            if (reverse) for (int k = 0; k < 3; ++k) normal[k] = -normal[k];
            const float gap = reverse ? sr - dist : dist - sr;
            const float stiff_k = o_nHn(Hii, normal) + mass / (gap * gap);
            float fp[3] = {0, 0, 0}, g[3], h[9];
            o_push(e, normal, ghat, stiff_k, fp, H);
            for (int k = 0; k < 3; ++k) f[k] += fp[k];
            o_friction(fp, dx, normal, friction, friction_eps, g, h);
            for (int k = 0; k < 3; ++k) f[k] += g[k];
            for (int k = 0; k < 9; ++k) H[k] += h[k];
        }
    }
    for (unsigned j = 0; j < n_floor; ++j) {
        const float *fl = floors + 10 * j;
        const float ghat = fl[3], friction = o_combine_friction(fl[4], vfriction, friction_mode), thickness = fl[5];
        const float *up = fl + 6;
        float e[3], p[3], q[3];
        for (int k = 0; k < 3; ++k) e[k] = x[k] - (fl[k] + ghat * up[k]);
        const float eu = o_dot(e, up);
        if (!(eu < 0.0f)) continue;
        if (thickness > 0.0f && -eu > thickness) continue;
        ++count;
        for (int k = 0; k < 3; ++k) { p[k] = -((-eu) * up[k]); q[k] = x[k] - fl[k]; }
        float gap = o_dot(q, up);
        if (fl[9] != 0.0f) gap = fmaxf(gap, constraint_tol * ghat);
        const float stiff_k = o_nHn(Hii, up) + mass / (gap * gap);
        float fp[3] = {0, 0, 0}, g[3], h[9];
        o_push(p, up, ghat, stiff_k, fp, H);
        for (int k = 0; k < 3; ++k) f[k] += fp[k];
        o_friction(fp, dx, up, friction, friction_eps, g, h);
        for (int k = 0; k < 3; ++k) f[k] += g[k];
        for (int k = 0; k < 9; ++k) H[k] += h[k];
    }
    return count;
}
static inline float o_collider_toi(const float *y0, const float *y1, int pinned, const float *spheres, unsigned n_sphere, const float *floors, unsigned n_floor,
                                   float max_t, int *infeasible) {
    float x1[3], toi = max_t;
    for (int k = 0; k < 3; ++k) x1[k] = max_t * (y1[k] - y0[k]) + y0[k];
    for (unsigned j = 0; j < n_sphere; ++j) {
        const float *s = spheres + 10 * j;
        if (s[9] != 0.0f) continue;
        const int bowl = s[7] != 0.0f, reverse = s[8] != 0.0f;
        float c0[3] = {s[0], s[1], s[2]}, c1[3] = {s[0], s[1], s[2]}, a[3], b[3];
        if (bowl && y0[1] > s[1]) c0[1] = y0[1];
        if (bowl && x1[1] > s[1]) c1[1] = x1[1];
        o_sub(y0, c0, a); o_sub(x1, c1, b);
        const float r = s[5], r0 = sqrtf(o_dot(a, a)), r1 = sqrtf(o_dot(b, b));
        if (pinned) {
            const float d0 = reverse ? (r - r0) : (r0 - r), d1 = reverse ? (r - r1) : (r1 - r);
            if (d0 >= 0.0f && d1 < 0.0f) *infeasible = 1;
            continue;
        }
        const float depth0 = reverse ? (r0 - r) : (r - r0);
        if (s[6] > 0.0f && depth0 > s[6]) continue;
        if ((r0 - r) * (r1 - r) <= 0.0f) toi = fminf(toi, max_t * ((r - r0) / (r1 - r0)));
    }
    for (unsigned j = 0; j < n_floor; ++j) {
        const float *fl = floors + 10 * j;
        if (fl[9] != 0.0f) continue;
        float a[3], b[3];
        o_sub(y0, fl, a); o_sub(x1, fl, b);
        const float h0 = o_dot(fl + 6, a), h1 = o_dot(fl + 6, b);
        if (pinned) { if (h0 >= 0.0f && h1 < 0.0f) *infeasible = 1; continue; }
        if (fl[5] > 0.0f && -h0 > fl[5]) continue;
        if (h1 < 0.0f) toi = fminf(toi, max_t * (-h0 / (h1 - h0)));
    }
    return toi;
}


/* ---- static collision mesh (TEST INFRASTRUCTURE restatement of contact/contact.cu:899-1151) -----------------------
 * One dynamic vertex / face / edge against one static face / vertex / edge: the push penalty of the analytic colliders
 * with the elasticity-inclusive stiffness of the dynamic element.  Outputs: f[3] (force on the contact point), H[9]
 * (column-major), c[] (weights of the element's vertices: vertex k receives c[k] f, block (j,k) receives c[j] c[k] H). */
static inline float o_cmesh_stiffness(int n, const float *B /* [n][n][9] column-major blocks */, const float *vmass, float gap_squared, const float *c,
                                      const float *normal) {
    float v[9], Av[9];
    for (int j = 0; j < 3 * n; ++j) v[j] = c[j / 3] * normal[j % 3];
    for (int r = 0; r < 3 * n; ++r) {
        float acc = 0.0f;
        for (int k = 0; k < 3 * n; ++k) {
            float a = B[((r / 3) * n + (k / 3)) * 9 + (k % 3) * 3 + (r % 3)];
            if (r == k) a += vmass[r / 3] / gap_squared;
            acc += a * v[k];
        }
        Av[r] = acc;
    }
    float num = 0.0f, den = 0.0f;
    for (int r = 0; r < 3 * n; ++r) { num += Av[r] * v[r]; den += v[r] * v[r]; }
    return num / den;
}
static inline void o_cmesh_force(const float *a, const float *normal, const float *dx, float ghat, float mu, float stiff_k, float friction_eps, float *f,
                                 float *H) {
    float fg[3], fh[9];
    for (int k = 0; k < 3; ++k) f[k] = 0.0f;
    for (int k = 0; k < 9; ++k) H[k] = 0.0f;
    o_push(a, normal, ghat, stiff_k, f, H);
    o_friction(f, dx, normal, mu, friction_eps, fg, fh);
    for (int k = 0; k < 3; ++k) f[k] += fg[k];
    for (int k = 0; k < 9; ++k) H[k] = H[k] + fh[k];
}
static inline void o_normalize(const float *e, float *n) {
    const float l = sqrtf(o_dot(e, e));
    for (int k = 0; k < 3; ++k) n[k] = l > 0.0f ? e[k] / l : e[k];
}
static inline int o_cmesh_vertex_face(const float *p, const float *p_start, const float *T0, const float *T1, const float *T2, float offset, float ghat,
                                      float mu, float mass, const float *Hii, float friction_eps, float *f, float *H, float *c3) {
    const float zero[3] = {0.0f, 0.0f, 0.0f};
    float t0[3], t1[3], t2[3], c[3], y[3], e[3], mq[3], normal[3], a[3];
    o_sub(T0, p, t0); o_sub(T1, p, t1); o_sub(T2, p, t2);
    o_point_triangle_coeff_unclassified(zero, t0, t1, t2, c);
    for (int k = 0; k < 3; ++k) { y[k] = c[0] * t0[k] + c[1] * t1[k] + c[2] * t2[k]; e[k] = -y[k]; mq[k] = -(p_start[k] - p[k]); }
    const float s = offset + ghat;
    if (!(o_dot(e, e) < s * s)) return 0;
    o_normalize(e, normal);
    for (int k = 0; k < 3; ++k) a[k] = -(y[k] + normal[k] * s);
    const float gap = sqrtf(o_dot(e, e)) - offset;
    const float stiff_k = o_nHn(Hii, normal) + mass / (gap * gap);
    o_cmesh_force(a, normal, mq, ghat, mu, stiff_k, friction_eps, f, H);
    c3[0] = 1.0f; c3[1] = c3[2] = 0.0f;
    return 1;
}
static inline int o_cmesh_face_vertex(const float *y, const float (*X)[3], const float (*Xs)[3], float offset, float ghat, float mu, const float *vmass,
                                      const float *B /* [3][3][9] */, float friction_eps, float *f, float *H, float *c3) {
    const float zero[3] = {0.0f, 0.0f, 0.0f};
    float t[3][3], sv[3][3], c[3], p[3], q[3], normal[3], a[3], dx[3];
    for (int j = 0; j < 3; ++j) { o_sub(X[j], y, t[j]); o_sub(Xs[j], y, sv[j]); }
    o_point_triangle_coeff_unclassified(zero, t[0], t[1], t[2], c);
    for (int k = 0; k < 3; ++k) { p[k] = c[0] * t[0][k] + c[1] * t[1][k] + c[2] * t[2][k]; q[k] = c[0] * sv[0][k] + c[1] * sv[1][k] + c[2] * sv[2][k]; }
    const float s = offset + ghat;
    if (!(o_dot(p, p) < s * s)) return 0;
    o_normalize(p, normal);
    const float gap = sqrtf(o_dot(p, p)) - offset;
    for (int k = 0; k < 3; ++k) c3[k] = c[k];
    const float stiff_k = o_cmesh_stiffness(3, B, vmass, gap * gap, c3, normal);
    for (int k = 0; k < 3; ++k) { a[k] = p[k] - normal[k] * s; dx[k] = p[k] - q[k]; }
    o_cmesh_force(a, normal, dx, ghat, mu, stiff_k, friction_eps, f, H);
    return 1;
}
static inline int o_cmesh_edge_edge(const float (*P)[3], const float (*Ps)[3], const float *Q0, const float *Q1, float offset, float ghat, float mu,
                                    const float *vmass, const float *B /* [2][2][9] */, float friction_eps, float *f, float *H, float *c3) {
    float c[4], x[3], y[3], z[3], e[3], normal[3], a[3], dx[3];
    o_edge_edge_coeff(P[0], P[1], Q0, Q1, c);
    for (int k = 0; k < 3; ++k) {
        x[k] = c[0] * P[0][k] + c[1] * P[1][k]; y[k] = c[2] * Q0[k] + c[3] * Q1[k]; z[k] = c[0] * Ps[0][k] + c[1] * Ps[1][k];
        e[k] = x[k] - y[k];
    }
    const float s = offset + ghat;
    if (!(o_dot(e, e) < s * s)) return 0;
    o_normalize(e, normal);
    const float gap = sqrtf(o_dot(e, e)) - offset;
    c3[0] = c[0]; c3[1] = c[1]; c3[2] = 0.0f;
    const float stiff_k = o_cmesh_stiffness(2, B, vmass, gap * gap, c3, normal);
    for (int k = 0; k < 3; ++k) { a[k] = e[k] - s * normal[k]; dx[k] = x[k] - z[k]; }
    o_cmesh_force(a, normal, dx, ghat, mu, stiff_k, friction_eps, f, H);
    return 1;
}


/* ---- torque pins (TEST INFRASTRUCTURE restatement of energy/energy.cu:20-86, 643-701, 221-246) ------------------- */
static inline void o_eigen_sym3x3(float a00, float a01, float a02, float a11, float a12, float a22, float evecs[3][3] /* [column][component] */) {
    const float q = (a00 + a11 + a22) / 3.0f;
    const float p2 = (a00 - q) * (a00 - q) + (a11 - q) * (a11 - q) + (a22 - q) * (a22 - q) + 2.0f * (a01 * a01 + a02 * a02 + a12 * a12);
    const float p = sqrtf(fmaxf(p2 / 6.0f, 0.0f));
    float lam0, lam1, lam2;
    if (p < 1e-12f) { lam0 = a00; lam1 = a11; lam2 = a22; }
    else {
        const float inv_p = 1.0f / p;
        const float b00 = (a00 - q) * inv_p, b01 = a01 * inv_p, b02 = a02 * inv_p, b11 = (a11 - q) * inv_p, b12 = a12 * inv_p, b22 = (a22 - q) * inv_p;
        const float det_b = b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02);
        const float r = fminf(1.0f, fmaxf(-1.0f, det_b * 0.5f));
        const float phi = acosf(r) / 3.0f;
        lam0 = q + 2.0f * p * cosf(phi);
        lam2 = q + 2.0f * p * cosf(phi + 2.0943951f);
        lam1 = 3.0f * q - lam0 - lam2;
    }
    float t;
    if (lam1 > lam0) { t = lam0; lam0 = lam1; lam1 = t; }
    if (lam2 > lam0) { t = lam0; lam0 = lam2; lam2 = t; }
    if (lam2 > lam1) { t = lam1; lam1 = lam2; lam2 = t; }
    const float vals[3] = {lam0, lam1, lam2};
    for (int c = 0; c < 3; ++c) {
        const float l = vals[c];
        const float r0[3] = {a00 - l, a01, a02}, r1[3] = {a01, a11 - l, a12}, r2[3] = {a02, a12, a22 - l};
        float v[3];
        o_cross(r0, r1, v);
        float vn = sqrtf(o_dot(v, v));
        if (vn < 1e-8f) { o_cross(r0, r2, v); vn = sqrtf(o_dot(v, v)); }
        if (vn < 1e-8f) { o_cross(r1, r2, v); vn = sqrtf(o_dot(v, v)); }
        for (int k = 0; k < 3; ++k) evecs[c][k] = vn > 1e-8f ? v[k] / vn : (k == c ? 1.0f : 0.0f);
    }
    float e0[3], e1[3], e2[3];
    o_normalize(evecs[0], e0);
    const float d = o_dot(e0, evecs[1]);
    for (int k = 0; k < 3; ++k) e1[k] = evecs[1][k] - e0[k] * d;
    o_normalize(e1, e1);
    o_cross(e0, e1, e2);
    for (int k = 0; k < 3; ++k) { evecs[0][k] = e0[k]; evecs[1][k] = e1[k]; evecs[2][k] = e2[k]; }
}
/* group g: vertices tv_index[vs .. vs+vn); x = 3 floats per vertex, mass per vertex; out7 = centre, axis, 1 / sum |r_perp|^2 */
static inline void o_torque_group(unsigned axis_component, unsigned vs, unsigned vn, unsigned hint_vertex, const unsigned *tv_index, const float *x,
                                  const float *mass, float *out7) {
    float center[3] = {0.0f, 0.0f, 0.0f}, total_mass = 0.0f;
    for (unsigned k = vs; k < vs + vn; ++k) {
        const unsigned idx = tv_index[k];
        const float m = mass[idx];
        for (int q = 0; q < 3; ++q) center[q] = center[q] + x[3 * (size_t)idx + q] * m;
        total_mass += m;
    }
    if (total_mass > 1e-12f) { const float s = 1.0f / total_mass; for (int q = 0; q < 3; ++q) center[q] = center[q] * s; }
    float c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
    for (unsigned k = vs; k < vs + vn; ++k) {
        float d[3];
        o_sub(x + 3 * (size_t)tv_index[k], center, d);
        c00 += d[0] * d[0]; c01 += d[0] * d[1]; c02 += d[0] * d[2]; c11 += d[1] * d[1]; c12 += d[1] * d[2]; c22 += d[2] * d[2];
    }
    const float inv_n = 1.0f / (float)vn;
    c00 *= inv_n; c01 *= inv_n; c02 *= inv_n; c11 *= inv_n; c12 *= inv_n; c22 *= inv_n;
    float evecs[3][3], ax[3], hint[3];
    o_eigen_sym3x3(c00, c01, c02, c11, c12, c22, evecs);
    o_normalize(evecs[axis_component < 3 ? axis_component : 0], ax);
    o_sub(x + 3 * (size_t)hint_vertex, center, hint);
    if (o_dot(ax, hint) < 0.0f) for (int q = 0; q < 3; ++q) ax[q] = -ax[q];
    float sum = 0.0f;
    for (unsigned k = vs; k < vs + vn; ++k) {
        float r[3], rp[3];
        o_sub(x + 3 * (size_t)tv_index[k], center, r);
        const float d = o_dot(ax, r);
        for (int q = 0; q < 3; ++q) rp[q] = r[q] - ax[q] * d;
        sum += o_dot(rp, rp);
    }
    for (int q = 0; q < 3; ++q) { out7[q] = center[q]; out7[3 + q] = ax[q]; }
    out7[6] = sum > 1e-12f ? 1.0f / sum : 0.0f;
}
static inline void o_torque_vertex(const float *y, float magnitude, const float *grp7, float *f, float *H) {
    const float *center = grp7, *axis = grp7 + 3;
    float r[3], rp[3], cr[3];
    o_sub(y, center, r);
    const float d = o_dot(axis, r);
    for (int q = 0; q < 3; ++q) rp[q] = r[q] - axis[q] * d;
    const float scale = magnitude * grp7[6];
    o_cross(axis, rp, cr);
    for (int q = 0; q < 3; ++q) f[q] = -(cr[q] * scale);
    const float sp = fmaxf(0.0f, scale);
    for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) H[c * 3 + rr] = sp * ((rr == c ? 1.0f : 0.0f) - axis[rr] * axis[c]);
}


/* ---- inflation pressure (TEST INFRASTRUCTURE restatement of energy/model/inflate.hpp:16-150, utility/utility.cu:60-183) ---- */
static inline void o_jacobi_svd3(const float *F /* column-major */, float *U, float *S, float *V) {
    float A[3][3], Vv[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { A[r][c] = F[c * 3 + r]; Vv[r][c] = r == c ? 1.0f : 0.0f; }
    for (int sweep = 0; sweep < 5; ++sweep)
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 3; ++q) {
                float a = 0.0f, b = 0.0f, g = 0.0f;
                for (int r = 0; r < 3; ++r) { a += A[r][p] * A[r][p]; b += A[r][q] * A[r][q]; g += A[r][p] * A[r][q]; }
                const float zeta = (b - a) / (2.0f * g);
                const float t = (g == 0.0f) ? 0.0f : (zeta >= 0.0f ? 1.0f : -1.0f) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
                const float c = 1.0f / sqrtf(1.0f + t * t), s = c * t;
                for (int r = 0; r < 3; ++r) {
                    const float ap = A[r][p], aq = A[r][q];
                    A[r][p] = c * ap - s * aq; A[r][q] = s * ap + c * aq;
                    const float vp = Vv[r][p], vq = Vv[r][q];
                    Vv[r][p] = c * vp - s * vq; Vv[r][q] = s * vp + c * vq;
                }
            }
    float s_[3], u_[3][3];
    for (int c = 0; c < 3; ++c) {
        float n2 = 0.0f;
        for (int r = 0; r < 3; ++r) n2 += A[r][c] * A[r][c];
        const float n = sqrtf(n2);
        s_[c] = n;
        const float inv = n > 0.0f ? 1.0f / n : 0.0f;
        for (int r = 0; r < 3; ++r) u_[r][c] = A[r][c] * inv;
    }
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) if (s_[idx[j]] < s_[idx[i]]) { const int tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp; }
    for (int c = 0; c < 3; ++c) { S[c] = s_[idx[c]]; for (int r = 0; r < 3; ++r) { U[c * 3 + r] = u_[r][idx[c]]; V[c * 3 + r] = Vv[r][idx[c]]; } }
}
static inline float o_det3(const float *m) { return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]); }
static inline void o_svd3_rv(const float *F, float *U, float *S, float *V) {
    o_jacobi_svd3(F, U, S, V);
    const float du = o_det3(U), dv = o_det3(V);
    int mi = 0;
    for (int i = 1; i < 3; ++i) if (S[i] < S[mi]) mi = i;
    if (du < 0.0f && dv > 0.0f) { for (int r = 0; r < 3; ++r) U[mi * 3 + r] = -U[mi * 3 + r]; S[mi] = -S[mi]; }
    else if (du > 0.0f && dv < 0.0f) { for (int r = 0; r < 3; ++r) V[mi * 3 + r] = -V[mi * 3 + r]; S[mi] = -S[mi]; }
}
static inline void o_jacobi_eig3(float A[3][3], float *val, float vec[3][3]) {
    float Q[3][3] = {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.f, 1.f}};
    for (int sweep = 0; sweep < 8; ++sweep)
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const float apq = A[p][q];
                if (apq == 0.0f) continue;
                const float theta = (A[q][q] - A[p][p]) / (2.0f * apq);
                const float t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(1.0f + theta * theta));
                const float c = 1.0f / sqrtf(1.0f + t * t), s = c * t;
                for (int k = 0; k < 3; ++k) { const float akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const float apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const float qkp = Q[k][p], qkq = Q[k][q]; Q[k][p] = c * qkp - s * qkq; Q[k][q] = s * qkp + c * qkq; }
            }
    for (int k = 0; k < 3; ++k) { val[k] = A[k][k]; for (int r = 0; r < 3; ++r) vec[k][r] = Q[r][k]; }
}
static inline void o_inflate_gradient(float pressure, const float *x0, const float *x1, const float *x2, float g[3][3]) {
    const float coeff = -pressure / 6.0f;
    float e1[3], e2[3], d[3], a[3], b[3];
    o_sub(x1, x0, e1); o_sub(x2, x0, e2); o_sub(e1, e2, d);
    o_cross(e1, e2, a); o_cross(d, x0, b);
    for (int k = 0; k < 3; ++k) g[0][k] = coeff * (a[k] + b[k]);
    o_cross(e2, x0, a); for (int k = 0; k < 3; ++k) g[1][k] = coeff * a[k];
    o_cross(x0, e1, a); for (int k = 0; k < 3; ++k) g[2][k] = coeff * a[k];
}
static inline void o_inflate_hessian(float pressure, const float *x0, const float *x1, const float *x2, float *H /* 9x9 column-major */) {
    float X[9], U[9], V[9], s[3];
    for (int k = 0; k < 3; ++k) { X[k] = x0[k]; X[3 + k] = x1[k]; X[6 + k] = x2[k]; }
    o_svd3_rv(X, U, s, V);
    float A[3][3] = {{0.f, s[2], s[1]}, {s[2], 0.f, s[0]}, {s[1], s[0], 0.f}};
    float ev[3], evec[3][3];
    o_jacobi_eig3(A, ev, evec);
    const float c = pressure / 6.0f, inv_sqrt2 = 0.7071067811865475f;
    float weights[5], Q[5][3][3];
    int kept = 0;
    const int swap_ij[3][2] = {{1, 0}, {2, 0}, {2, 1}};
    const float swap_s[3] = {s[2], s[1], s[0]};
    for (int p = 0; p < 3; ++p) {
        const float w = c * swap_s[p];
        if (!(w > 0.0f)) continue;
        const int i = swap_ij[p][0], j = swap_ij[p][1];
        weights[kept] = w;
        for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) Q[kept][k][r] = inv_sqrt2 * (V[j * 3 + k] * U[i * 3 + r] + V[i * 3 + k] * U[j * 3 + r]);
        ++kept;
    }
    for (int d = 0; d < 3 && kept < 5; ++d) {
        if (!(ev[d] < -1e-12f)) continue;
        const float w = c * fabsf(ev[d]);
        if (!(w > 0.0f)) continue;
        weights[kept] = w;
        for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r)
            Q[kept][k][r] = ((evec[d][0] * V[0 * 3 + k]) * U[0 * 3 + r] + (evec[d][1] * V[1 * 3 + k]) * U[1 * 3 + r]) + (evec[d][2] * V[2 * 3 + k]) * U[2 * 3 + r];
        ++kept;
    }
    for (int i = 0; i < 81; ++i) H[i] = 0.0f;
    for (int m = 0; m < kept; ++m)
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int cc = 0; cc < 3; ++cc) for (int r = 0; r < 3; ++r)
            H[(3 * b + cc) * 9 + 3 * a + r] += weights[m] * Q[m][a][r] * Q[m][b][cc];
}

#endif

// math_distance.cuh -- closest-point barycentric coefficients (fp32).
//
// These predicates decide which primitive pairs are contacts, so they follow the
// reference's arithmetic operation by operation:
//   point-edge      crates/ppf-cts-solver/src/cpp/contact/distance.hpp:15-27
//   point-triangle  distance.hpp:47-84   (cross-product form, degenerate fallback)
//   edge-edge       distance.hpp:113-196 (five-candidate scan)
//   unclassified    distance.hpp:224-301
#pragma once
#include "vecmath.cuh"

namespace ppf {

struct C2 { float a, b; };
struct C3 { float a, b, c; };
struct C4 { float a, b, c, d; };

PPF_HD C2 point_edge_coeff(const V3 &p, const V3 &e0, const V3 &e1) {
    const V3 r = e1 - e0;
    const float d = dot(r, r);
    C2 o;
    if (d > 0.0f) {
        const float t = dot(r, p - e0) / d;
        o.a = 1.0f - t;
        o.b = t;
    } else {
        o.a = 1.0f / 2.0f;
        o.b = 1.0f / 2.0f;
    }
    return o;
}

PPF_HD C3 point_triangle_coeff(const V3 &p, const V3 &t0, const V3 &t1, const V3 &t2) {
    const V3 r0 = t1 - t0;
    const V3 r1 = t2 - t0;
    const V3 d = p - t0;
    const V3 nrm = cross(r0, r1);
    const float nn = dot(nrm, nrm);
    C3 o;
    if (nn > 0.0f) {
        const float inv_nn = 1.0f / nn;
        const float u = dot(cross(d, r1), nrm) * inv_nn;
        const float v = dot(cross(r0, d), nrm) * inv_nn;
        o.a = 1.0f - u - v;
        o.b = u;
        o.c = v;
    } else {
        const V3 e1 = t2 - t1;
        const V3 e2 = t0 - t2;
        const float len0 = dot(r0, r0), len1 = dot(e1, e1), len2 = dot(e2, e2);
        if (len0 >= len1 && len0 >= len2) {
            const C2 w = point_edge_coeff(p, t0, t1);
            o.a = w.a; o.b = w.b; o.c = 0.0f;
        } else if (len1 >= len2) {
            const C2 w = point_edge_coeff(p, t1, t2);
            o.a = 0.0f; o.b = w.a; o.c = w.b;
        } else {
            const C2 w = point_edge_coeff(p, t2, t0);
            o.a = w.b; o.b = 0.0f; o.c = w.a;
        }
    }
    return o;
}

// smat.hpp:346-361 minCoeff / maxCoeff (first-element seeded scans; NaN-preserving)
PPF_HD float min_coeff(const float *m, int n) { float best = m[0]; for (int i = 1; i < n; ++i) if (m[i] < best) best = m[i]; return best; }
PPF_HD float max_coeff(const float *m, int n) { float best = m[0]; for (int i = 1; i < n; ++i) if (best < m[i]) best = m[i]; return best; }
PPF_HD float clamp_unit(float v) { return v > 0.0f ? (v < 1.0f ? v : 1.0f) : 0.0f; }

PPF_HD C4 edge_edge_coeff(const V3 &ea0, const V3 &ea1, const V3 &eb0, const V3 &eb1) {
    const V3 r0 = ea1 - ea0;
    const V3 r1 = eb1 - eb0;
    const V3 d = ea0 - eb0;
    const float a = dot(r0, r0);
    const float e = dot(r1, r1);
    const float b = dot(r0, r1);
    const float c = dot(r0, d);
    const float f = dot(r1, d);
    const float inv_a = a > 0.0f ? 1.0f / a : 0.0f;
    const float inv_e = e > 0.0f ? 1.0f / e : 0.0f;
    const V3 n = cross(r0, r1);
    const float nn = dot(n, n);
    const float inv_nn = nn > 0.0f ? 1.0f / nn : 0.0f;
    const float s_stat = clamp_unit(dot(n, cross(r1, d)) * inv_nn);
    const float t_stat = clamp_unit((b * s_stat + f) * inv_e);
    const float t_a0 = clamp_unit(f * inv_e);
    const float t_a1 = clamp_unit((f + b) * inv_e);
    const float s_b0 = clamp_unit(-c * inv_a);
    const float s_b1 = clamp_unit((b - c) * inv_a);
    const float cand_s[5] = {s_stat, 0.0f, 1.0f, s_b0, s_b1};
    const float cand_t[5] = {t_stat, t_a0, t_a1, 0.0f, 1.0f};
    float best_s = 0.0f, best_t = 0.0f, best_dist = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const V3 v = d + cand_s[i] * r0 - cand_t[i] * r1;
        const float dist = dot(v, v);
        if (dist < best_dist) {
            best_dist = dist;
            best_s = cand_s[i];
            best_t = cand_t[i];
        }
    }
    C4 o;
    o.a = 1.0f - best_s; o.b = best_s; o.c = 1.0f - best_t; o.d = best_t;
    return o;
}

PPF_HD C3 point_triangle_coeff_unclassified(const V3 &p, const V3 &t0, const V3 &t1, const V3 &t2) {
    const V3 r0 = t1 - t0;
    const V3 r1 = t2 - t0;
    const V3 d = p - t0;
    C3 cand[4];
    bool admissible[4] = {false, true, true, true};
    cand[0] = point_triangle_coeff(p, t0, t1, t2);
    { const float m0[3] = {cand[0].a, cand[0].b, cand[0].c}; admissible[0] = min_coeff(m0, 3) >= 0.0f; }
    const float s01 = clamp_unit(point_edge_coeff(p, t0, t1).b);
    const float s12 = clamp_unit(point_edge_coeff(p, t1, t2).b);
    const float s20 = clamp_unit(point_edge_coeff(p, t2, t0).b);
    cand[1].a = 1.0f - s01; cand[1].b = s01; cand[1].c = 0.0f;
    cand[2].a = 0.0f; cand[2].b = 1.0f - s12; cand[2].c = s12;
    cand[3].a = s20; cand[3].b = 0.0f; cand[3].c = 1.0f - s20;
    C3 best = cand[1];
    float best_dist = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!admissible[i]) continue;
        const V3 v = d - cand[i].b * r0 - cand[i].c * r1;
        const float dist = dot(v, v);
        if (dist < best_dist) {
            best_dist = dist;
            best = cand[i];
        }
    }
    return best;
}

PPF_HD C2 point_edge_coeff_unclassified(const V3 &p, const V3 &e0, const V3 &e1) {
    const C2 c = point_edge_coeff(p, e0, e1);
    if (c.a >= 0.0f && c.a <= 1.0f) return c;
    C2 o;
    if (c.a > 1.0f) { o.a = 1.0f; o.b = 0.0f; }
    else { o.a = 0.0f; o.b = 1.0f; }
    return o;
}

PPF_HD bool strictly_interior(const C2 &c) { const float m[2] = {c.a, c.b}; return max_coeff(m, 2) < 1.0f && min_coeff(m, 2) > 0.0f; }
PPF_HD bool strictly_interior(const C3 &c) { const float m[3] = {c.a, c.b, c.c}; return max_coeff(m, 3) < 1.0f && min_coeff(m, 3) > 0.0f; }
PPF_HD bool strictly_interior(const C4 &c) { const float m[4] = {c.a, c.b, c.c, c.d}; return max_coeff(m, 4) < 1.0f && min_coeff(m, 4) > 0.0f; }

} // namespace ppf

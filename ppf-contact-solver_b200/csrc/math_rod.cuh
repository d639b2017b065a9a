// math_rod.cuh -- rod (yarn) elements: Hookean stretch of one rod edge and the bending energy
// E = 1/2 (theta - theta0)^2 of an interior rod vertex with its analytically PSD-projected Hessian (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/energy/model/hook.hpp:20-37 (make_diff_table) and
// energy/model/dihedral_angle.hpp:397-425 (gradient_theta), 451-541 (strand_compute_force_hessian:
// closed-form angle-energy eigen-system of Wu & Kim 2023, four in-plane and two out-of-plane eigenpairs
// in the 6-D space of the two segment vectors, mapped to the 9 DOFs by J^T [qa; qb] = [qa; -qa-qb; qb]),
// plus the rod strain-limit barrier of strainlimiting/strainlimiting.cu:125-190.
#pragma once
#include "math_contact.cuh"
#include "vecmath.cuh"

namespace ppf {

// Hookean rod edge (x0, x1): gradient per vertex and the one 3x3 block K that makes up the 6x6 Hessian
// [[K, -K], [-K, K]]  (hook.hpp:20-37: weight * (c0 g g^T + c1 dtdx^T dtdx), g = [-n; n], dtdx = [-I, I])
struct RodElastic {
    V3 f[2];
    M3 K;
};
PPF_HD void rod_hook(const V3 &x0, const V3 &x1, float l0, float weight, RodElastic &o) {
    const V3 t = x1 - x0;
    const float l = norm(t);
    const V3 n = t / l;
    const V3 dedt = (l / l0 - 1.0f) * n;
    const float r = (l - l0) / l;
    const float c0 = fmaxf(0.0f, 1.0f - r) / l0;
    const float c1 = fmaxf(0.0f, r / l0);
    o.f[0] = (-weight) * dedt;
    o.f[1] = weight * dedt;
    const float nv[3] = {n.x, n.y, n.z};
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) o.K.m[c * 3 + rr] = weight * ((c0 * nv[rr]) * nv[c] + (rr == c ? c1 : 0.0f));
}

// bending of the polyline (x0, x1, x2) about its middle vertex
struct StrandElastic {
    V3 f[3];
    float H[81]; // 9x9 column-major, vertex order (x0, x1, x2)
};
// gradient of the turning angle w.r.t. the segment vectors e0 = x0 - x1, e1 = x2 - x1 (dihedral_angle.hpp:397-425)
PPF_HD void strand_gradient_theta(const V3 &e0, const V3 &e1, V3 &g0, V3 &g1) {
    const V3 cr = cross(e0, e1);
    const float cross_sqnm = sqnorm(cr);
    const float e0_sqnm = dot(e0, e0);
    const float e1_sqnm = dot(e1, e1);
    const float kSinSqEps = 1e-12f;
    if (cross_sqnm <= kSinSqEps * e0_sqnm * e1_sqnm) {
        g0 = mk(0.f, 0.f, 0.f);
        g1 = mk(0.f, 0.f, 0.f);
        return;
    }
    const V3 n = cr / sqrtf(cross_sqnm);
    const V3 e0perp = cross(e0, n);
    const V3 e1perp = cross(e1, n);
    g0 = e0perp / e0_sqnm;
    g1 = (-e1perp) / e1_sqnm;
}
PPF_HD void strand_accum(float *H, float lam, const V3 &qa, const V3 &qb) {
    if (lam <= 0.0f) return;
    const float n2 = dot(qa, qa) + dot(qb, qb);
    if (n2 <= 1e-20f) return;
    float p9[9];
    for (int c = 0; c < 3; ++c) {
        const float a = comp(qa, c), b = comp(qb, c);
        p9[c] = a;
        p9[3 + c] = -a - b;
        p9[6 + c] = b;
    }
    const float s = lam / n2;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) H[j * 9 + i] = H[j * 9 + i] + s * p9[i] * p9[j];
}
PPF_HD void strand_force_hessian(const V3 &x0, const V3 &x1, const V3 &x2, float rest_angle, StrandElastic &o) {
    const V3 e0 = x0 - x1;
    const V3 e1 = x2 - x1;
    const V3 bcr = cross(e0, e1);
    const float theta = atan2f(norm(bcr), dot(e0, e1));
    V3 g0, g1;
    strand_gradient_theta(e0, e1, g0, g1);
    const float w = theta - rest_angle;
    o.f[0] = w * g0;
    o.f[1] = w * ((-g0) - g1);
    o.f[2] = w * g1;
    for (int i = 0; i < 81; ++i) o.H[i] = 0.0f;

    const float l0 = norm(e0);
    const float l1 = norm(e1);
    const float lb = norm(bcr);
    if (lb <= 1e-9f * l0 * l1 || l0 <= 1e-9f || l1 <= 1e-9f) return;
    const V3 tb = bcr / lb;
    const V3 e0p = cross(e0, tb);
    const V3 e1p = cross(e1, tb);
    const float cos_t = dot(e0, e1) / (l0 * l1);
    const float sin_t = lb / (l0 * l1);
    const float gamma = (l1 * l1) / (l0 * l0);
    const float gm1 = gamma - 1.0f;
    const float gp1 = gamma + 1.0f;
    const float r = sqrtf(4.0f * w * w * (gm1 / gp1) * (gm1 / gp1) + 1.0f);
    const float Rm = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f - r)));
    const float Rp = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f + r)));
    const float Lh = 1.0f / (l0 * l0) + 1.0f / (l1 * l1);
    const float cc = 4.0f * w / gp1;
    // four in-plane eigenpairs
    const int sr[4] = {-1, -1, 1, 1};
    const float RR[4] = {Rm, Rm, Rp, Rp};
    const int sR[4] = {-1, 1, -1, 1};
    for (int m = 0; m < 4; ++m) {
        const float R = RR[m];
        const float a = w * (gm1 * gm1 / gp1 + sr[m] * gp1 * r + sR[m] * gm1 * R);
        const float b = -0.5f * gamma * R * R + 2.0f * w * w - sR[m] * 0.5f * (gm1 + sr[m] * gp1 * r) * R;
        const float d = 1.0f + sr[m] * r + sR[m] * R;
        const float lam = 0.25f * Lh * d;
        strand_accum(o.H, lam, a * e0 + b * e0p, cc * e1 + d * e1p);
    }
    // two out-of-plane eigenpairs
    const float beta = (l1 / l0 - l0 / l1) * cos_t;
    const float alpha0 = 0.5f * (-beta + sqrtf(beta * beta + 4.0f));
    const float ginv = w / sin_t;
    const float lam4 = ginv * (cos_t / (l1 * l1) - alpha0 / (l0 * l1));
    const float lam5 = ginv * (cos_t / (l1 * l1) + 1.0f / (alpha0 * l0 * l1));
    strand_accum(o.H, lam4, alpha0 * tb, tb);
    strand_accum(o.H, lam5, tb, (-alpha0) * tb);
}

// per-vertex bending coefficient of embed_rod_bend_force_hessian (energy.cu:572-603): bend and the rest lengths
// of the two incident rod edges, mass of the vertex; REF_LENGTH = 1 cm anchors the numeric range of `bend`
PPF_HD float rod_bend_coefficient(float bend0, float bend1, float len0, float len1, float mass) {
    const float bend = 0.5f * (bend0 + bend1);
    const float REF_LENGTH = 1e-2f;
    const float voronoi = 0.5f * (len0 + len1);
    return voronoi > 0.0f ? bend * mass * (REF_LENGTH * REF_LENGTH) / (voronoi * voronoi) : 0.0f;
}

// ---- rod strain limiting (strainlimiting/strainlimiting.cu:100-190) ----
// Barrier on the stretch g = l / l0 - 1 of a rod edge beyond its DRAWN length l0 (EdgeProp.initial_length), gap = the
// strain limit.  Same block structure as the Hookean term: Hessian [[K, -K], [-K, K]].  Hel00 / Hel01 / Hel11 are the
// blocks (e0,e0), (e0,e1), (e1,e1) of the elastic-only matrix (FixedCSRMat::operator(), transposed when stored as (e1,e0));
// c0, c1 are the start-of-step positions the stiffness is evaluated at.  Returns false when the edge is not stretched.
PPF_HD bool rod_strain(const V3 &x0, const V3 &x1, const V3 &c0, const V3 &c1, float l0, float limit, float mass, const M3 &Hel00,
                       const M3 &Hel01, const M3 &Hel11, RodElastic &o, int barrier = 0) {
    const V3 d = x1 - x0;
    const float l = norm(d);
    const float g = l / l0 - 1.0f;
    if (!(g > 0.0f)) return false;
    const float y = limit - g;
    const float dEdg = -barrier_gradient(y, limit, 0.0f, barrier);
    const float d2Edg2 = barrier_curvature(y, limit, 0.0f, barrier);
    const V3 n = d / l;
    const float inv_ll0 = 1.0f / (l * l0);
    const float a_psd = fmaxf(0.0f, d2Edg2);
    const float b_psd = fmaxf(0.0f, dEdg);
    // compute_rod_stiffness: w^T H_el w + mass / (limit - g)^2 with w = (c0 - center, c1 - center)
    const V3 center = 0.5f * c0 + 0.5f * c1;
    const V3 w0 = c0 - center, w1 = c1 - center;
    const V3 y0 = mul(Hel00, w0) + mul(Hel01, w1);
    const V3 y1 = mul_t(Hel01, w0) + mul(Hel11, w1);
    const float dd = limit - g;
    const float stiffness = (dot(w0, y0) + dot(w1, y1)) + mass / (dd * dd);
    const float nv[3] = {n.x, n.y, n.z};
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            const float P = (r == c ? 1.0f : 0.0f) - nv[r] * nv[c];
            o.K.m[c * 3 + r] = stiffness * (a_psd * ((nv[r] / l0) * (nv[c] / l0)) + b_psd * (inv_ll0 * P));
        }
    const V3 j1 = n / l0;
    o.f[0] = stiffness * (dEdg * (-j1));
    o.f[1] = stiffness * (dEdg * j1);
    return true;
}
// strainlimiting::rod_line_search (strainlimiting.cu:192-245): largest t in (0, max_t] keeping the stretch of the
// segment d0 + t (d1 - d0) below the limit, by bisection until the bracket stops shrinking
PPF_HD float rod_strain_toi(const V3 &d0, const V3 &d1, float l0, float limit, float max_t) {
    float t = max_t;
    const V3 dd = d1 - d0;
    const float g_t = norm(d0 + t * dd) / l0 - 1.0f;
    if (g_t >= limit) {
        float upper_t = t, lower_t = 0.0f;
        float window = upper_t - lower_t;
        while (true) {
            t = 0.5f * (upper_t + lower_t);
            const float g_here = norm(d0 + t * dd) / l0 - 1.0f;
            if (g_here - limit < 0.0f) lower_t = t;
            else upper_t = t;
            const float new_window = upper_t - lower_t;
            if (new_window == window) break;
            window = new_window;
        }
        t = lower_t;
    }
    return t;
}

} // namespace ppf

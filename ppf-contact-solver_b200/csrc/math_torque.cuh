// math_torque.cuh -- torque pins: a group of free vertices driven by a torque about a principal axis of the group (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/energy/energy.cu:20-86 (eigen_sym3x3: Cardano eigenvalues, eigenvectors from row
// cross products, Gram-Schmidt), 643-701 (compute_torque_groups: mass-weighted centroid, covariance, chosen principal axis
// oriented toward a hint vertex, sum of squared lever arms) and 221-246 (the per-vertex force axis x r_perp * magnitude /
// sum |r_perp|^2 with the PSD part of its symmetric Hessian).
#pragma once
#include "../../include/ppf_abi.h"
#include "vecmath.cuh"

namespace ppf {

PPF_HD float cos_bounded(float x) { // float_math.hpp:52-61: the fast intrinsic on the device
#if defined(__CUDA_ARCH__)
    return __cosf(x);
#else
    return cosf(x);
#endif
}

// eigenvectors of the symmetric matrix (a00 a01 a02; . a11 a12; . . a22), columns sorted by descending eigenvalue
PPF_HD void eigen_sym3x3(float a00, float a01, float a02, float a11, float a12, float a22, V3 *evecs) {
    const float q = (a00 + a11 + a22) / 3.0f;
    const float p2 = (a00 - q) * (a00 - q) + (a11 - q) * (a11 - q) + (a22 - q) * (a22 - q) + 2.0f * (a01 * a01 + a02 * a02 + a12 * a12);
    const float p = sqrtf(fmaxf(p2 / 6.0f, 0.0f));
    float lam0, lam1, lam2;
    if (p < 1e-12f) {
        lam0 = a00; lam1 = a11; lam2 = a22;
    } else {
        const float inv_p = 1.0f / p;
        const float b00 = (a00 - q) * inv_p, b01 = a01 * inv_p, b02 = a02 * inv_p;
        const float b11 = (a11 - q) * inv_p, b12 = a12 * inv_p;
        const float b22 = (a22 - q) * inv_p;
        const float det_b = b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02);
        const float r = fminf(1.0f, fmaxf(-1.0f, det_b * 0.5f));
        const float phi = acosf(r) / 3.0f;
        lam0 = q + 2.0f * p * cos_bounded(phi);
        lam2 = q + 2.0f * p * cos_bounded(phi + 2.0943951f);
        lam1 = 3.0f * q - lam0 - lam2;
    }
    if (lam1 > lam0) { const float t = lam0; lam0 = lam1; lam1 = t; }
    if (lam2 > lam0) { const float t = lam0; lam0 = lam2; lam2 = t; }
    if (lam2 > lam1) { const float t = lam1; lam1 = lam2; lam2 = t; }
    const float vals[3] = {lam0, lam1, lam2};
    for (int c = 0; c < 3; ++c) {
        const float l = vals[c];
        const V3 r0 = mk(a00 - l, a01, a02), r1 = mk(a01, a11 - l, a12), r2 = mk(a02, a12, a22 - l);
        V3 v = cross(r0, r1);
        float vn = norm(v);
        if (vn < 1e-8f) { v = cross(r0, r2); vn = norm(v); }
        if (vn < 1e-8f) { v = cross(r1, r2); vn = norm(v); }
        evecs[c] = vn > 1e-8f ? v / vn : mk(c == 0 ? 1.f : 0.f, c == 1 ? 1.f : 0.f, c == 2 ? 1.f : 0.f);
    }
    const V3 e0 = normalized(evecs[0]);
    V3 e1 = evecs[1] - e0 * dot(e0, evecs[1]);
    e1 = normalized(e1);
    evecs[0] = e0; evecs[1] = e1; evecs[2] = cross(e0, e1);
}

struct TorqueGroupResult {
    V3 center, axis;
    float inv_r_perp_sq_sum;
};
// pos(v) / mass(v): evaluation position and mass of vertex v
template <class Pos, class Mass>
PPF_HD TorqueGroupResult torque_group(const PpfTorqueGroup &grp, const PpfTorqueVertex *tv, Pos pos, Mass mass) {
    const unsigned vs = grp.vertex_start, vn = grp.vertex_count;
    V3 center = mk(0.f, 0.f, 0.f);
    float total_mass = 0.0f;
    for (unsigned k = vs; k < vs + vn; ++k) {
        const unsigned idx = tv[k].index;
        const float m = mass(idx);
        center += pos(idx) * m;
        total_mass += m;
    }
    if (total_mass > 1e-12f) center = center * (1.0f / total_mass);
    float c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
    for (unsigned k = vs; k < vs + vn; ++k) {
        const V3 d = pos(tv[k].index) - center;
        c00 += d.x * d.x; c01 += d.x * d.y; c02 += d.x * d.z;
        c11 += d.y * d.y; c12 += d.y * d.z; c22 += d.z * d.z;
    }
    const float inv_n = 1.0f / (float)vn;
    c00 *= inv_n; c01 *= inv_n; c02 *= inv_n; c11 *= inv_n; c12 *= inv_n; c22 *= inv_n;
    V3 evecs[3];
    eigen_sym3x3(c00, c01, c02, c11, c12, c22, evecs);
    TorqueGroupResult r;
    r.center = center;
    V3 ax = normalized(evecs[grp.axis_component < 3 ? grp.axis_component : 0]);
    const V3 hint_dir = pos(grp.hint_vertex) - center;
    if (dot(ax, hint_dir) < 0.0f) ax = -ax;
    r.axis = ax;
    float sum = 0.0f;
    for (unsigned k = vs; k < vs + vn; ++k) {
        const V3 rr = pos(tv[k].index) - center;
        const V3 r_perp = rr - ax * dot(ax, rr);
        sum += sqnorm(r_perp);
    }
    r.inv_r_perp_sq_sum = sum > 1e-12f ? 1.0f / sum : 0.0f;
    return r;
}
// force (to ADD to the gradient) and Hessian of one torque vertex at y
PPF_HD void torque_vertex(const V3 &y, float magnitude, const TorqueGroupResult &g, V3 &f, M3 &H) {
    const V3 r = y - g.center;
    const V3 r_perp = r - g.axis * dot(g.axis, r);
    const float scale = magnitude * g.inv_r_perp_sq_sum;
    const V3 torque_force = cross(g.axis, r_perp) * scale;
    f = -torque_force;
    H = fmaxf(0.0f, scale) * (m3_identity() - m3_outer(g.axis, g.axis));
}

} // namespace ppf

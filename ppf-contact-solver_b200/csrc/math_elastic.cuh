// math_elastic.cuh -- per-element elastic force / PSD-projected Hessian (fp32).
//
// Shell membrane (Baraff-Witkin with the closed-form eigen-systems of Kim 2020):
//   crates/ppf-cts-solver/src/cpp/energy/energy.cu:304-364 (embed_face_force_hessian)
//   energy/model/baraffwitkin.hpp:15-229 (stretch / shear gradient + clamped eigen Hessians)
//   utility/utility.cu:185-282 (deformation gradient, convert_force, convert_hessian)
// Inertia / air drag:
//   energy/energy.cu:182-270, energy/model/momentum.hpp, energy/model/air_damper.hpp
// The results only need to agree with the reference to fp32 tolerance (they feed the linear
// system, not a contact predicate), so the algebra is reorganised where that saves work:
// the stretch Hessian is formed as q q^T + max(0, 1-1/|f|)(I - q q^T) instead of three
// explicit eigenvectors (identical in exact arithmetic).
#pragma once
#include "vecmath.cuh"

namespace ppf {

struct FaceElastic {
    V3 f[3];    // dE/dx per vertex (the reference's "force" = gradient)
    M3 H[3][3]; // d2E/dx2 blocks
};

PPF_HD void v6_set(float *o, const V3 &a, const V3 &b) { o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; }
PPF_HD float v6_dot(const float *a, const float *b) { float s = 0.0f; for (int i = 0; i < 6; ++i) s = s + a[i] * b[i]; return s; }
PPF_HD float v6_sq(const float *a) { return v6_dot(a, a); }

// picks, among three candidates, the one with the strictly largest squared norm (third wins ties),
// as baraffwitkin.hpp does for its null-space vectors
PPF_HD int pick3(float n0, float n1, float n2) {
    if (n0 > n1 && n0 > n2) return 0;
    if (n1 > n0 && n1 > n2) return 1;
    return 2;
}

// PSD stretch Hessian of one column f of F: eigenpairs (1, f/|f|), (1 - 1/|f|, two tangents)
PPF_HD M3 bw_stretch_block(const V3 &f) {
    const float n = norm(f);
    const V3 q = normalized(f);
    const float lam = fmaxf(0.0f, 1.0f - 1.0f / n);
    const M3 qq = m3_outer(q, q);
    M3 r = qq + lam * (m3_identity() - qq);
    return r;
}

// shear Hessian (6x6, column-major h[c*6+r]) = 0.5 * lmd * sum_k max(0, lam_k) q_k q_k^T  (baraffwitkin.hpp:117-225)
PPF_HD void bw_shear_hessian(const V3 &fu, const V3 &fv, float lmd, float *h) {
    const float I6 = dot(fu, fv);
    const float I2 = sqnorm(fu) + sqnorm(fv);
    float g[6], Hg[6];
    v6_set(g, fv, fu);
    v6_set(Hg, fu, fv);
    const float root = sqrtf(I2 * I2 + 12.0f * I6 * I6);
    const float lmd0 = I2 + root;
    float q0[6];
    for (int i = 0; i < 6; ++i) q0[i] = 2.0f * I6 * Hg[i] + lmd0 * g[i];
    float z[6], w1[6] = {0, 0, 0, 0, 0, 0}, w2[6] = {0, 0, 0, 0, 0, 0}, w3[6] = {0, 0, 0, 0, 0, 0};
    float lmd1;
    // NOTE: vectors here index the 6-vector vec(F) = [Fu; Fv]; w_k couple component k of Fu and Fv.
    // The reference writes w1 = (1,+-1,0,0,0,0) etc. over the SAME storage order; keep it literally.
    if (I6 > 0) {
        for (int i = 0; i < 6; ++i) z[i] = q0[i];
        lmd1 = 2.0f * I6;
        w1[0] = 1; w1[1] = 1; w2[2] = 1; w2[3] = 1; w3[4] = 1; w3[5] = 1;
    } else {
        lmd1 = -2.0f * I6;
        const float tmp = I2 - root;
        for (int i = 0; i < 6; ++i) z[i] = 2.0f * I6 * Hg[i] + tmp * g[i];
        w1[0] = 1; w1[1] = -1; w2[2] = 1; w2[3] = -1; w3[4] = 1; w3[5] = -1;
    }
    const float lmd2 = lmd1;
    const float a1 = v6_dot(z, w1), a2 = v6_dot(z, w2), a3 = v6_dot(z, w3);
    float c1[3][6];
    for (int i = 0; i < 6; ++i) {
        c1[0][i] = a3 * w1[i] + a3 * w2[i] - (a1 - a2) * w3[i];
        c1[1][i] = a2 * w1[i] - (a1 - a3) * w2[i] + a2 * w3[i];
        c1[2][i] = -(a2 - a3) * w1[i] + a1 * w2[i] + a1 * w3[i];
    }
    const int k1 = pick3(v6_sq(c1[0]), v6_sq(c1[1]), v6_sq(c1[2]));
    float q1[6];
    for (int i = 0; i < 6; ++i) q1[i] = c1[k1][i];
    const float b1 = v6_dot(q1, w1), b2 = v6_dot(q1, w2), b3 = v6_dot(q1, w3);
    // 2x2 systems A1=[a2 a3; b2 b3], A2=[a3 a1; b3 b1], A3=[a1 a2; b1 b2]; adj(A) = [A11 -A01; -A10 A00]
    const float detA1 = a2 * b3 - a3 * b2, detA2 = a3 * b1 - a1 * b3, detA3 = a1 * b2 - a2 * b1;
    // c45 = adj(A3) * (-a3,-b3), c56 = adj(A1) * (-a1,-b1), c64 = adj(A2) * (-a2,-b2)
    const float c45_0 = b2 * (-a3) + (-a2) * (-b3), c45_1 = (-b1) * (-a3) + a1 * (-b3);
    const float c56_0 = b3 * (-a1) + (-a3) * (-b1), c56_1 = (-b2) * (-a1) + a2 * (-b1);
    const float c64_0 = b1 * (-a2) + (-a1) * (-b2), c64_1 = (-b3) * (-a2) + a3 * (-b2);
    float c2[3][6];
    for (int i = 0; i < 6; ++i) {
        c2[0][i] = c45_0 * w1[i] + c45_1 * w2[i] + detA3 * w3[i];
        c2[1][i] = detA1 * w1[i] + c56_0 * w2[i] + c56_1 * w3[i];
        c2[2][i] = c64_1 * w1[i] + detA2 * w2[i] + c64_0 * w3[i];
    }
    const int k2 = pick3(v6_sq(c2[0]), v6_sq(c2[1]), v6_sq(c2[2]));
    float q2[6];
    for (int i = 0; i < 6; ++i) q2[i] = c2[k2][i];
    const float n0 = sqrtf(v6_sq(q0)), n1 = sqrtf(v6_sq(q1)), n2 = sqrtf(v6_sq(q2));
    for (int i = 0; i < 6; ++i) {
        if (n0 > 0.0f) q0[i] = q0[i] / n0;
        if (n1 > 0.0f) q1[i] = q1[i] / n1;
        if (n2 > 0.0f) q2[i] = q2[i] / n2;
    }
    const float l0 = fmaxf(0.0f, lmd0), l1 = fmaxf(0.0f, lmd1), l2 = fmaxf(0.0f, lmd2);
    const float s = 0.5f * lmd;
    for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) h[c * 6 + r] = s * (l0 * q0[r] * q0[c] + l1 * q1[r] * q1[c] + l2 * q2[r] * q2[c]);
}

// inv_rest is the 2x2 column-major matrix m[c*2+r] of the ABI (DataSet.inv_rest2x2)
PPF_HD void bw_face(const V3 &x0, const V3 &x1, const V3 &x2, const float *inv_rest, float mu, float lmd, float mass,
                    FaceElastic &o) {
    const V3 d0 = x1 - x0, d1 = x2 - x0;
    const float r00 = inv_rest[0], r10 = inv_rest[1], r01 = inv_rest[2], r11 = inv_rest[3];
    // F = [d0 d1] * inv_rest
    const V3 fu = d0 * r00 + d1 * r10;
    const V3 fv = d0 * r01 + d1 * r11;
    // gradient in F space
    const float nu = norm(fu), nv = norm(fv);
    const float w = dot(fu, fv);
    const V3 gu = mu * (fu * ((nu - 1.0f) / nu)) + lmd * (w * fv);
    const V3 gv = mu * (fv * ((nv - 1.0f) / nv)) + lmd * (w * fu);
    // Hessian in F space: stretch is block diagonal, shear is full 6x6
    float hs[36];
    bw_shear_hessian(fu, fv, lmd, hs);
    const M3 su = mu * bw_stretch_block(fu), sv = mu * bw_stretch_block(fv);
    M3 Hde[2][2];
    for (int d = 0; d < 2; ++d)
        for (int e = 0; e < 2; ++e)
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) Hde[d][e].m[c * 3 + r] = hs[(3 * e + c) * 6 + (3 * d + r)];
    Hde[0][0] += su;
    Hde[1][1] += sv;
    // chain rule: dF/dx_a = g[a] (utility.cu convert_force / convert_hessian); g0 = -row0 - row1
    const float g[3][2] = {{-r00 - r10, -r01 - r11}, {r00, r01}, {r10, r11}};
    for (int a = 0; a < 3; ++a) o.f[a] = mass * (g[a][0] * gu + g[a][1] * gv);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            M3 blk = m3_zero();
            for (int d = 0; d < 2; ++d)
                for (int e = 0; e < 2; ++e) blk += (g[a][d] * g[b][e]) * Hde[d][e];
            o.H[a][b] = mass * blk;
        }
}

} // namespace ppf

// math_hinge.cuh -- discrete-shell hinge bending: gradient and PSD-projected Hessian of
// E = 1/2 (theta - theta0)^2 (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/energy/model/dihedral_angle.hpp:
//   remap 15-17, face_dihedral_angle_grad 19-46, face_dihedral_angle 81-96,
//   face_compute_force_hessian 130-375 (closed-form eigen-system of Wu & Kim 2023:
//   six "F-part" eigenpairs + a rank-4 "W-part" projected with a 4x4 Jacobi sweep),
// and linalg/eigsolve.hpp:245-322 (psd_project_symmetric<4>).
// Output is in the REMAPPED vertex order (h2, h1, h0, h3), the order the reference scatters
// with and the order DataSet.hinge_hess_slots is built in (builder.rs:1437-1445).
#pragma once
#include "vecmath.cuh"

namespace ppf {

struct HingeElastic {
    V3 f[4];
    float H[144]; // 12x12 column-major, remapped vertex order
};

PPF_HD float hinge_angle(const V3 &v0, const V3 &v1, const V3 &v2, const V3 &v3) {
    const V3 n1 = cross(v1 - v0, v2 - v0);
    const V3 n2 = cross(v2 - v3, v1 - v3);
    const V3 e = v1 - v2;
    const float e_norm = norm(e);
    if (e_norm <= 0.0f) return 0.0f;
    const float sin_t = dot(cross(n2, n1), e) / e_norm;
    const float cos_t = dot(n1, n2);
    return atan2f(sin_t, cos_t);
}

// gradient of the angle w.r.t. (X0,X1,X2,X3) = remapped vertices; argument naming of the reference:
// face_dihedral_angle_grad(v2, v0, v1, v3) called with (X0, X1, X2, X3)
PPF_HD void hinge_angle_grad(const V3 &v2, const V3 &v0, const V3 &v1, const V3 &v3, V3 *out) {
    const V3 e0 = v1 - v0, e1 = v2 - v0, e2 = v3 - v0, e3 = v2 - v1, e4 = v3 - v1;
    const V3 n1 = cross(e0, e1), n2 = cross(e2, e0);
    const float n1_sq = sqnorm(n1), n2_sq = sqnorm(n2), e0_n = norm(e0);
    out[0] = (-e0_n / n1_sq) * n1;
    out[1] = (-dot(e0, e3) / (e0_n * n1_sq)) * n1 - (dot(e0, e4) / (e0_n * n2_sq)) * n2;
    out[2] = (dot(e0, e1) / (e0_n * n1_sq)) * n1 + (dot(e0, e2) / (e0_n * n2_sq)) * n2;
    out[3] = (-e0_n / n2_sq) * n2;
}

// symmetric 4x4 PSD projection by cyclic Jacobi (eigsolve.hpp:245-322), floor = 0
PPF_HD void psd_project4(float A[4][4]) {
    float V[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0f : 0.0f;
    float scale = 0.0f;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) scale = fmaxf(scale, fabsf(A[i][j]));
    if (scale == 0.0f) return;
    const float tiny = 1e-7f * scale;
    for (int sweep = 0; sweep < 24; ++sweep) {
        float off = 0.0f;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) off = fmaxf(off, fabsf(A[p][q]));
        if (off <= tiny) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const float apq = A[p][q];
                if (fabsf(apq) <= tiny) continue;
                const float app = A[p][p], aqq = A[q][q];
                const float tau = (aqq - app) / (2.0f * apq);
                const float t = (tau >= 0.0f ? 1.0f : -1.0f) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
                const float c = 1.0f / sqrtf(1.0f + t * t);
                const float s = t * c;
                A[p][p] = c * c * app - 2.0f * s * c * apq + s * s * aqq;
                A[q][q] = s * s * app + 2.0f * s * c * apq + c * c * aqq;
                A[p][q] = 0.0f;
                A[q][p] = 0.0f;
                for (int k = 0; k < 4; ++k) {
                    if (k == p || k == q) continue;
                    const float akp = A[k][p], akq = A[k][q];
                    const float np = c * akp - s * akq, nq = s * akp + c * akq;
                    A[k][p] = np; A[p][k] = np; A[k][q] = nq; A[q][k] = nq;
                }
                for (int k = 0; k < 4; ++k) {
                    const float vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    float lam[4];
    for (int i = 0; i < 4; ++i) lam[i] = fmaxf(A[i][i], 0.0f);
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            float sum = 0.0f;
            for (int k = 0; k < 4; ++k) sum += V[i][k] * lam[k] * V[j][k];
            A[i][j] = sum;
            A[j][i] = sum;
        }
}

// X[0..3]: positions of the REMAPPED hinge vertices (h2, h1, h0, h3)
PPF_HD void hinge_force_hessian(const V3 *X, float rest_angle, HingeElastic &o) {
    const V3 X0 = X[0], X1 = X[1], X2 = X[2], X3 = X[3];
    const float angle = hinge_angle(X0, X1, X2, X3);
    V3 grad[4];
    hinge_angle_grad(X0, X1, X2, X3, grad);
    for (int k = 0; k < 4; ++k) o.f[k] = (angle - rest_angle) * grad[k];
    float *H = o.H;
    for (int i = 0; i < 144; ++i) H[i] = 0.0f;

    const V3 y1 = X1 - X2, y2 = X0 - X2, y3 = X3 - X2;
    const float ly1 = norm(y1);
    if (ly1 <= 1e-9f) return;
    const V3 t1 = y1 / ly1;
    const M3 sigma = m3_identity() - m3_outer(t1, t1);
    const V3 z0 = y2 - t1 * dot(t1, y2);
    const V3 z1 = y3 - t1 * dot(t1, y3);
    const float l0 = norm(z0), l1 = norm(z1);
    const V3 bcr = cross(z0, z1);
    const float lb = norm(bcr);
    if (l0 <= 1e-9f || l1 <= 1e-9f || lb <= 1e-9f * l0 * l1) return;
    const V3 tb = bcr / lb;
    const float Sgeom = dot(tb, t1) >= 0.0f ? 1.0f : -1.0f;
    const float g = angle - rest_angle;

    // Hp is accumulated in the local (P0=X2, P1=X1, P2=X0, P3=X3) order and permuted at the end.
    float Hp[144];
    for (int i = 0; i < 144; ++i) Hp[i] = 0.0f;
#define HP(i, j) Hp[(j) * 12 + (i)]

    const V3 e0p = cross(z0, tb), e1p = cross(z1, tb);
    const float S3 = -Sgeom;
    const float w = S3 * g;
    const float gamma = (l1 * l1) / (l0 * l0);
    const float gm1 = gamma - 1.0f, gp1 = gamma + 1.0f;
    const float r = sqrtf(4.0f * w * w * (gm1 / gp1) * (gm1 / gp1) + 1.0f);
    const float Rm = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f - r)));
    const float Rp = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f + r)));
    const float Lh = 1.0f / (l0 * l0) + 1.0f / (l1 * l1);
    const float cc = 4.0f * w / gp1;
    const M3 eta2 = (-1.0f / ly1) * (dot(t1, y2) * sigma + m3_outer(t1, z0));
    const M3 eta3 = (-1.0f / ly1) * (dot(t1, y3) * sigma + m3_outer(t1, z1));
    const M3 e2T = transpose(eta2), e3T = transpose(eta3);
    M3 m0, m1;
    for (int i = 0; i < 9; ++i) { m0.m[i] = -e2T.m[i] - sigma.m[i]; m1.m[i] = -e3T.m[i] - sigma.m[i]; }

    // rank-1 accumulation  Hp += lam/|q|^2 * p p^T,  p = B^T [qa; qb]
    auto accF = [&](float lam, const V3 &qa, const V3 &qb) {
        if (lam <= 0.0f) return;
        const float n2 = dot(qa, qa) + dot(qb, qb);
        if (n2 <= 1e-20f) return;
        const V3 b0 = mul(m0, qa) + mul(m1, qb);
        const V3 b1 = mul(e2T, qa) + mul(e3T, qb);
        const V3 b2 = mul(sigma, qa);
        const V3 b3 = mul(sigma, qb);
        const float p12[12] = {b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, b2.x, b2.y, b2.z, b3.x, b3.y, b3.z};
        const float s = lam / n2;
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) HP(i, j) += s * p12[i] * p12[j];
    };
    const int sr[4] = {-1, -1, 1, 1};
    const float RR[4] = {Rm, Rm, Rp, Rp};
    const int sR[4] = {-1, 1, -1, 1};
    for (int mm = 0; mm < 4; ++mm) {
        const float R = RR[mm];
        const float a = w * (gm1 * gm1 / gp1 + sr[mm] * gp1 * r + sR[mm] * gm1 * R);
        const float b = -0.5f * gamma * R * R + 2.0f * w * w - sR[mm] * 0.5f * (gm1 + sr[mm] * gp1 * r) * R;
        const float d = 1.0f + sr[mm] * r + sR[mm] * R;
        accF(0.25f * Lh * d, a * z0 + b * e0p, cc * z1 + d * e1p);
    }
    const float cos_t = dot(z0, z1) / (l0 * l1);
    const float sin_t = lb / (l0 * l1);
    const float beta = (l1 / l0 - l0 / l1) * cos_t;
    const float alpha0 = 0.5f * (-beta + sqrtf(beta * beta + 4.0f));
    const float ginv = w / sin_t;
    accF(ginv * (cos_t / (l1 * l1) - alpha0 / (l0 * l1)), alpha0 * tb, tb);
    accF(ginv * (cos_t / (l1 * l1) + 1.0f / (alpha0 * l0 * l1)), tb, (-alpha0) * tb);

    // W-part: rank-4 correction in the span of pc[0..3], PSD-projected in an orthonormalised basis
    const V3 tau0 = z0 / l0, tau1 = z1 / l1;
    const V3 tau1p = cross(tau1, tb);
    const float a0 = dot(t1, y2) / (ly1 * l0), a1 = dot(t1, y3) / (ly1 * l1);
    const float c0 = 1.0f / l0, c1 = 1.0f / l1;
    const float k0 = -4.0f * dot(tau0, tau1p) / (ly1 * ly1);
    const float k1 = 2.0f * dot(tau0, tau1p) / ly1;
    const V3 s01 = tau0 + tau1, d01 = tau0 - tau1;
    float pc[4][12];
    for (int c = 0; c < 3; ++c) {
        const float sc = comp(s01, c), dc = comp(d01, c), tc = comp(t1, c);
        pc[0][c] = sc;  pc[0][3 + c] = -sc; pc[0][6 + c] = 0.0f; pc[0][9 + c] = 0.0f;
        pc[1][c] = dc;  pc[1][3 + c] = -dc; pc[1][6 + c] = 0.0f; pc[1][9 + c] = 0.0f;
        pc[2][c] = (-a1 - a0 + c1 + c0) * tc; pc[2][3 + c] = (a1 + a0) * tc; pc[2][6 + c] = -c0 * tc; pc[2][9 + c] = -c1 * tc;
        pc[3][c] = (-a1 + a0 + c1 - c0) * tc; pc[3][3 + c] = (a1 - a0) * tc; pc[3][6 + c] = c0 * tc;  pc[3][9 + c] = -c1 * tc;
    }
    float gg[4], g23 = 0.0f;
    for (int i = 0; i < 4; ++i) {
        float s = 0.0f;
        for (int m = 0; m < 12; ++m) s += pc[i][m] * pc[i][m];
        gg[i] = s;
    }
    for (int m = 0; m < 12; ++m) g23 += pc[2][m] * pc[3][m];
    const float det23 = gg[2] * gg[3] - g23 * g23;
    if (gg[0] > 1e-5f && gg[1] > 1e-5f && det23 > 1e-6f * gg[2] * gg[3]) {
        const float we = -g * Sgeom;
        float E[4][4] = {{0.0f}};
        E[0][0] = we * (-k0 * gg[0]);
        E[1][1] = we * (k0 * gg[1]);
        E[0][2] = we * (k1 * gg[2]); E[2][0] = E[0][2];
        E[0][3] = we * (k1 * g23);   E[3][0] = E[0][3];
        E[1][2] = we * (k1 * g23);   E[2][1] = E[1][2];
        E[1][3] = we * (k1 * gg[3]); E[3][1] = E[1][3];
        float Li[4][4] = {{0.0f}};
        Li[0][0] = 1.0f / sqrtf(gg[0]);
        Li[1][1] = 1.0f / sqrtf(gg[1]);
        const float L22 = sqrtf(gg[2]);
        const float L32 = g23 / L22;
        const float L33 = sqrtf(fmaxf(1e-30f, gg[3] - L32 * L32));
        Li[2][2] = 1.0f / L22;
        Li[3][3] = 1.0f / L33;
        Li[3][2] = -L32 / (L22 * L33);
        float LiE[4][4], Ss[4][4];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = 0.0f;
                for (int k = 0; k < 4; ++k) s += Li[i][k] * E[k][j];
                LiE[i][j] = s;
            }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = 0.0f;
                for (int k = 0; k < 4; ++k) s += LiE[i][k] * Li[j][k];
                Ss[i][j] = s;
            }
        psd_project4(Ss);
        float LtS[4][4], Core[4][4];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = 0.0f;
                for (int k = 0; k < 4; ++k) s += Li[k][i] * Ss[k][j];
                LtS[i][j] = s;
            }
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = 0.0f;
                for (int k = 0; k < 4; ++k) s += LtS[i][k] * Li[k][j];
                Core[i][j] = s;
            }
        for (int m = 0; m < 12; ++m) {
            float PCm[4];
            for (int j = 0; j < 4; ++j) {
                float s = 0.0f;
                for (int i = 0; i < 4; ++i) s += pc[i][m] * Core[i][j];
                PCm[j] = s;
            }
            for (int n = 0; n < 12; ++n) {
                float s = 0.0f;
                for (int j = 0; j < 4; ++j) s += PCm[j] * pc[j][n];
                HP(m, n) += s;
            }
        }
    }
    // local order (P0,P1,P2,P3) = (X2,X1,X0,X3)  ->  remapped order (X0,X1,X2,X3)
    const int perm[4] = {2, 1, 0, 3};
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H[(3 * perm[b] + j) * 12 + (3 * perm[a] + i)] = HP(3 * a + i, 3 * b + j);
#undef HP
}

} // namespace ppf

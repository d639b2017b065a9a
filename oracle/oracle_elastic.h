/* oracle_elastic.h -- TEST INFRASTRUCTURE.  CPU restatement (plain C arrays, fp32) of the elastic
 * element math of the reference: Baraff-Witkin membrane and discrete-shell hinge bending.
 * Paths under /root/reference/crates/ppf-cts-solver/src/cpp/.
 *   membrane: energy/energy.cu:304-364, energy/model/baraffwitkin.hpp:15-229, utility/utility.cu:185-282
 *   hinge:    energy/energy.cu:448-553, energy/model/dihedral_angle.hpp:15-17,19-46,81-96,130-375,
 *             linalg/eigsolve.hpp:245-322 (4x4 Jacobi PSD projection)
 * Matrices are column-major flat arrays: a 9x9 / 12x12 element Hessian is H[c * N + r].
 */
#ifndef PPF_ORACLE_ELASTIC_H
#define PPF_ORACLE_ELASTIC_H
#include "oracle_math.h"

static inline float o_norm3(const float *a) { return sqrtf(o_dot(a, a)); }
static inline int o_pick3(float a, float b, float c) { return (a > b && a > c) ? 0 : ((b > a && b > c) ? 1 : 2); }
static inline float o_dot6(const float *a, const float *b) { float s = 0.0f; for (int i = 0; i < 6; ++i) s = s + a[i] * b[i]; return s; }

/* mass-scaled gradient g[3][3] (per vertex) and Hessian H[81] of one shell face */
static inline void o_bw_face(const float *x0, const float *x1, const float *x2, const float *ir /* 2x2 col-major */, float mu, float lmd, float mass,
                             float g[3][3], float *H) {
    o3 d0, d1, fu, fv;
    o_sub(x1, x0, d0); o_sub(x2, x0, d1);
    for (int k = 0; k < 3; ++k) { fu[k] = d0[k] * ir[0] + d1[k] * ir[1]; fv[k] = d0[k] * ir[2] + d1[k] * ir[3]; }
    const float nu = o_norm3(fu), nv = o_norm3(fv), w = o_dot(fu, fv);
    float gu[3], gv[3];
    for (int k = 0; k < 3; ++k) {
        gu[k] = mu * (fu[k] * (nu - 1.0f) / nu) + lmd * (w * fv[k]);
        gv[k] = mu * (fv[k] * (nv - 1.0f) / nv) + lmd * (w * fu[k]);
    }
    /* 6x6 Hessian in F space (col-major h[c*6+r]) */
    float h[36];
    for (int i = 0; i < 36; ++i) h[i] = 0.0f;
    { /* stretch: per column, eigenvalues 1 (along f) and 1 - 1/|f| (two tangents), clamped at 0 */
        const float *f[2] = {fu, fv};
        const float n[2] = {nu, nv};
        for (int c2 = 0; c2 < 2; ++c2) {
            const float lam = fmaxf(0.0f, 1.0f - 1.0f / n[c2]);
            float q[3];
            for (int k = 0; k < 3; ++k) q[k] = f[c2][k] / n[c2];
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) {
                    const float qq = q[r] * q[c];
                    h[(3 * c2 + c) * 6 + 3 * c2 + r] += mu * (qq + lam * ((r == c ? 1.0f : 0.0f) - qq));
                }
        }
    }
    { /* shear: three clamped eigenpairs (baraffwitkin.hpp:117-225) */
        const float I6 = w, I2 = o_dot(fu, fu) + o_dot(fv, fv);
        float gg[6], Hg[6], q0[6], z[6], w1[6] = {0}, w2[6] = {0}, w3[6] = {0};
        for (int k = 0; k < 3; ++k) { gg[k] = fv[k]; gg[3 + k] = fu[k]; Hg[k] = fu[k]; Hg[3 + k] = fv[k]; }
        const float root = sqrtf(I2 * I2 + 12.0f * I6 * I6), lmd0 = I2 + root;
        for (int i = 0; i < 6; ++i) q0[i] = 2.0f * I6 * Hg[i] + lmd0 * gg[i];
        float lmd1;
        if (I6 > 0) { memcpy(z, q0, sizeof(z)); lmd1 = 2.0f * I6; w1[0] = w1[1] = 1; w2[2] = w2[3] = 1; w3[4] = w3[5] = 1; }
        else {
            lmd1 = -2.0f * I6;
            const float tmp = I2 - root;
            for (int i = 0; i < 6; ++i) z[i] = 2.0f * I6 * Hg[i] + tmp * gg[i];
            w1[0] = 1; w1[1] = -1; w2[2] = 1; w2[3] = -1; w3[4] = 1; w3[5] = -1;
        }
        const float a1 = o_dot6(z, w1), a2 = o_dot6(z, w2), a3 = o_dot6(z, w3);
        float c1[3][6], c2[3][6], q1[6], q2[6];
        for (int i = 0; i < 6; ++i) {
            c1[0][i] = a3 * w1[i] + a3 * w2[i] - (a1 - a2) * w3[i];
            c1[1][i] = a2 * w1[i] - (a1 - a3) * w2[i] + a2 * w3[i];
            c1[2][i] = -(a2 - a3) * w1[i] + a1 * w2[i] + a1 * w3[i];
        }
        memcpy(q1, c1[o_pick3(o_dot6(c1[0], c1[0]), o_dot6(c1[1], c1[1]), o_dot6(c1[2], c1[2]))], sizeof(q1));
        const float b1 = o_dot6(q1, w1), b2 = o_dot6(q1, w2), b3 = o_dot6(q1, w3);
        const float dA1 = a2 * b3 - a3 * b2, dA2 = a3 * b1 - a1 * b3, dA3 = a1 * b2 - a2 * b1;
        const float c45_0 = b2 * (-a3) + (-a2) * (-b3), c45_1 = (-b1) * (-a3) + a1 * (-b3);
        const float c56_0 = b3 * (-a1) + (-a3) * (-b1), c56_1 = (-b2) * (-a1) + a2 * (-b1);
        const float c64_0 = b1 * (-a2) + (-a1) * (-b2), c64_1 = (-b3) * (-a2) + a3 * (-b2);
        for (int i = 0; i < 6; ++i) {
            c2[0][i] = c45_0 * w1[i] + c45_1 * w2[i] + dA3 * w3[i];
            c2[1][i] = dA1 * w1[i] + c56_0 * w2[i] + c56_1 * w3[i];
            c2[2][i] = c64_1 * w1[i] + dA2 * w2[i] + c64_0 * w3[i];
        }
        memcpy(q2, c2[o_pick3(o_dot6(c2[0], c2[0]), o_dot6(c2[1], c2[1]), o_dot6(c2[2], c2[2]))], sizeof(q2));
        const float n0 = sqrtf(o_dot6(q0, q0)), n1 = sqrtf(o_dot6(q1, q1)), n2 = sqrtf(o_dot6(q2, q2));
        for (int i = 0; i < 6; ++i) {
            if (n0 > 0.0f) q0[i] /= n0;
            if (n1 > 0.0f) q1[i] /= n1;
            if (n2 > 0.0f) q2[i] /= n2;
        }
        const float l0 = fmaxf(0.0f, lmd0), l1 = fmaxf(0.0f, lmd1), s = 0.5f * lmd;
        for (int c = 0; c < 6; ++c)
            for (int r = 0; r < 6; ++r) h[c * 6 + r] += s * (l0 * q0[r] * q0[c] + l1 * q1[r] * q1[c] + l1 * q2[r] * q2[c]);
    }
    /* chain rule dF/dx_a = G[a] (utility.cu:185-245) */
    const float G[3][2] = {{-ir[0] - ir[1], -ir[2] - ir[3]}, {ir[0], ir[2]}, {ir[1], ir[3]}};
    for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 3; ++k) g[a][k] = mass * (G[a][0] * gu[k] + G[a][1] * gv[k]);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) {
                    float v = 0.0f;
                    for (int d = 0; d < 2; ++d)
                        for (int e = 0; e < 2; ++e) v += (G[a][d] * G[b][e]) * h[(3 * e + c) * 6 + 3 * d + r];
                    H[(3 * b + c) * 9 + 3 * a + r] = mass * v;
                }
}

/* y = M x for a 3x3 column-major matrix */
static inline void o_mv3(const float *M, const float *x, float *y) {
    for (int r = 0; r < 3; ++r) y[r] = M[r] * x[0] + M[3 + r] * x[1] + M[6 + r] * x[2];
}

static inline void o_psd4(float A[4][4]) { /* linalg/eigsolve.hpp:245-322, floor 0 */
    float V[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0f : 0.0f;
    float scale = 0.0f;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) scale = fmaxf(scale, fabsf(A[i][j]));
    if (scale == 0.0f) return;
    const float tiny = 1e-7f * scale;
    for (int sweep = 0; sweep < 24; ++sweep) {
        float off = 0.0f;
        for (int p = 0; p < 4; ++p) for (int q = p + 1; q < 4; ++q) off = fmaxf(off, fabsf(A[p][q]));
        if (off <= tiny) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const float apq = A[p][q];
                if (fabsf(apq) <= tiny) continue;
                const float app = A[p][p], aqq = A[q][q];
                const float tau = (aqq - app) / (2.0f * apq);
                const float t = (tau >= 0.0f ? 1.0f : -1.0f) / (fabsf(tau) + sqrtf(1.0f + tau * tau));
                const float c = 1.0f / sqrtf(1.0f + t * t), s = t * c;
                A[p][p] = c * c * app - 2.0f * s * c * apq + s * s * aqq;
                A[q][q] = s * s * app + 2.0f * s * c * apq + c * c * aqq;
                A[p][q] = A[q][p] = 0.0f;
                for (int k = 0; k < 4; ++k) {
                    if (k == p || k == q) continue;
                    const float akp = A[k][p], akq = A[k][q];
                    const float np = c * akp - s * akq, nq = s * akp + c * akq;
                    A[k][p] = A[p][k] = np; A[k][q] = A[q][k] = nq;
                }
                for (int k = 0; k < 4; ++k) {
                    const float vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    float lam[4];
    for (int i = 0; i < 4; ++i) lam[i] = fmaxf(A[i][i], 0.0f);
    for (int i = 0; i < 4; ++i)
        for (int j = i; j < 4; ++j) {
            float sum = 0.0f;
            for (int k = 0; k < 4; ++k) sum += V[i][k] * lam[k] * V[j][k];
            A[i][j] = A[j][i] = sum;
        }
}

/* X[4][3]: REMAPPED hinge vertices (h2, h1, h0, h3).  f[4][3] gradient, H[144] Hessian (remapped order), unit stiffness */
static inline void o_hinge(const float (*X)[3], float rest_angle, float f[4][3], float *H) {
    o3 e0, e1, e2, e3, e4, n1, n2, t;
    /* angle (dihedral_angle.hpp:81-96) with (v0,v1,v2,v3) = X */
    o3 a, b;
    o_sub(X[1], X[0], a); o_sub(X[2], X[0], b); o_cross(a, b, n1);
    o_sub(X[2], X[3], a); o_sub(X[1], X[3], b); o_cross(a, b, n2);
    o3 ee; o_sub(X[1], X[2], ee);
    const float en = o_norm3(ee);
    float angle = 0.0f;
    if (en > 0.0f) { o_cross(n2, n1, t); angle = atan2f(o_dot(t, ee) / en, o_dot(n1, n2)); }
    /* gradient (19-46): face_dihedral_angle_grad(v2=X0, v0=X1, v1=X2, v3=X3) */
    o_sub(X[2], X[1], e0); o_sub(X[0], X[1], e1); o_sub(X[3], X[1], e2); o_sub(X[0], X[2], e3); o_sub(X[3], X[2], e4);
    o_cross(e0, e1, n1); o_cross(e2, e0, n2);
    const float n1s = o_dot(n1, n1), n2s = o_dot(n2, n2), e0n = o_norm3(e0);
    float gr[4][3];
    for (int k = 0; k < 3; ++k) {
        gr[0][k] = (-e0n / n1s) * n1[k];
        gr[1][k] = (-o_dot(e0, e3) / (e0n * n1s)) * n1[k] - (o_dot(e0, e4) / (e0n * n2s)) * n2[k];
        gr[2][k] = (o_dot(e0, e1) / (e0n * n1s)) * n1[k] + (o_dot(e0, e2) / (e0n * n2s)) * n2[k];
        gr[3][k] = (-e0n / n2s) * n2[k];
    }
    const float g = angle - rest_angle;
    for (int v = 0; v < 4; ++v) for (int k = 0; k < 3; ++k) f[v][k] = g * gr[v][k];
    for (int i = 0; i < 144; ++i) H[i] = 0.0f;
    o3 y1, y2, y3;
    o_sub(X[1], X[2], y1); o_sub(X[0], X[2], y2); o_sub(X[3], X[2], y3);
    const float ly1 = o_norm3(y1);
    if (ly1 <= 1e-9f) return;
    o3 t1, z0, z1, bcr, tb;
    for (int k = 0; k < 3; ++k) t1[k] = y1[k] / ly1;
    float sigma[9];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) sigma[c * 3 + r] = (r == c ? 1.0f : 0.0f) - t1[r] * t1[c];
    const float t1y2 = o_dot(t1, y2), t1y3 = o_dot(t1, y3);
    for (int k = 0; k < 3; ++k) { z0[k] = y2[k] - t1[k] * t1y2; z1[k] = y3[k] - t1[k] * t1y3; }
    const float l0 = o_norm3(z0), l1 = o_norm3(z1);
    o_cross(z0, z1, bcr);
    const float lb = o_norm3(bcr);
    if (l0 <= 1e-9f || l1 <= 1e-9f || lb <= 1e-9f * l0 * l1) return;
    for (int k = 0; k < 3; ++k) tb[k] = bcr[k] / lb;
    const float Sgeom = o_dot(tb, t1) >= 0.0f ? 1.0f : -1.0f;
    float Hp[144];
    for (int i = 0; i < 144; ++i) Hp[i] = 0.0f;
    o3 e0p, e1p;
    o_cross(z0, tb, e0p); o_cross(z1, tb, e1p);
    const float w = -Sgeom * g;
    const float gamma = (l1 * l1) / (l0 * l0), gm1 = gamma - 1.0f, gp1 = gamma + 1.0f;
    const float r = sqrtf(4.0f * w * w * (gm1 / gp1) * (gm1 / gp1) + 1.0f);
    const float Rm = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f - r))), Rp = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f + r)));
    const float Lh = 1.0f / (l0 * l0) + 1.0f / (l1 * l1), cc = 4.0f * w / gp1;
    float eta2[9], eta3[9], e2T[9], e3T[9], m0[9], m1[9];
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) {
            eta2[c * 3 + rr] = (-1.0f / ly1) * (t1y2 * sigma[c * 3 + rr] + t1[rr] * z0[c]);
            eta3[c * 3 + rr] = (-1.0f / ly1) * (t1y3 * sigma[c * 3 + rr] + t1[rr] * z1[c]);
        }
    for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) { e2T[c * 3 + rr] = eta2[rr * 3 + c]; e3T[c * 3 + rr] = eta3[rr * 3 + c]; }
    for (int i = 0; i < 9; ++i) { m0[i] = -e2T[i] - sigma[i]; m1[i] = -e3T[i] - sigma[i]; }
#define O_ACCF(LAM, QA, QB)                                                                                      \
    do {                                                                                                         \
        const float lam_ = (LAM);                                                                                \
        if (lam_ > 0.0f) {                                                                                       \
            const float n2_ = o_dot((QA), (QA)) + o_dot((QB), (QB));                                             \
            if (n2_ > 1e-20f) {                                                                                  \
                float u0[3], u1[3], p12[12];                                                                     \
                o_mv3(m0, (QA), u0); o_mv3(m1, (QB), u1); for (int k = 0; k < 3; ++k) p12[k] = u0[k] + u1[k];    \
                o_mv3(e2T, (QA), u0); o_mv3(e3T, (QB), u1); for (int k = 0; k < 3; ++k) p12[3 + k] = u0[k] + u1[k]; \
                o_mv3(sigma, (QA), u0); for (int k = 0; k < 3; ++k) p12[6 + k] = u0[k];                          \
                o_mv3(sigma, (QB), u0); for (int k = 0; k < 3; ++k) p12[9 + k] = u0[k];                          \
                const float s_ = lam_ / n2_;                                                                     \
                for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) Hp[j * 12 + i] += s_ * p12[i] * p12[j]; \
            }                                                                                                    \
        }                                                                                                        \
    } while (0)
    const int sr[4] = {-1, -1, 1, 1}, sR[4] = {-1, 1, -1, 1};
    const float RR[4] = {Rm, Rm, Rp, Rp};
    for (int mm = 0; mm < 4; ++mm) {
        const float R = RR[mm];
        const float aa = w * (gm1 * gm1 / gp1 + sr[mm] * gp1 * r + sR[mm] * gm1 * R);
        const float bb = -0.5f * gamma * R * R + 2.0f * w * w - sR[mm] * 0.5f * (gm1 + sr[mm] * gp1 * r) * R;
        const float dd = 1.0f + sr[mm] * r + sR[mm] * R;
        float qa[3], qb[3];
        for (int k = 0; k < 3; ++k) { qa[k] = aa * z0[k] + bb * e0p[k]; qb[k] = cc * z1[k] + dd * e1p[k]; }
        O_ACCF(0.25f * Lh * dd, qa, qb);
    }
    const float cos_t = o_dot(z0, z1) / (l0 * l1), sin_t = lb / (l0 * l1);
    const float beta = (l1 / l0 - l0 / l1) * cos_t, alpha0 = 0.5f * (-beta + sqrtf(beta * beta + 4.0f)), ginv = w / sin_t;
    {
        float qa[3], qb[3];
        for (int k = 0; k < 3; ++k) { qa[k] = alpha0 * tb[k]; qb[k] = tb[k]; }
        O_ACCF(ginv * (cos_t / (l1 * l1) - alpha0 / (l0 * l1)), qa, qb);
        for (int k = 0; k < 3; ++k) { qa[k] = tb[k]; qb[k] = (-alpha0) * tb[k]; }
        O_ACCF(ginv * (cos_t / (l1 * l1) + 1.0f / (alpha0 * l0 * l1)), qa, qb);
    }
#undef O_ACCF
    float tau0[3], tau1[3], tau1p[3];
    for (int k = 0; k < 3; ++k) { tau0[k] = z0[k] / l0; tau1[k] = z1[k] / l1; }
    o_cross(tau1, tb, tau1p);
    const float a0 = t1y2 / (ly1 * l0), a1 = t1y3 / (ly1 * l1), c0 = 1.0f / l0, c1 = 1.0f / l1;
    const float k0 = -4.0f * o_dot(tau0, tau1p) / (ly1 * ly1), k1 = 2.0f * o_dot(tau0, tau1p) / ly1;
    float pc[4][12];
    for (int c = 0; c < 3; ++c) {
        const float sc = tau0[c] + tau1[c], dc = tau0[c] - tau1[c], tc = t1[c];
        pc[0][c] = sc; pc[0][3 + c] = -sc; pc[0][6 + c] = 0.0f; pc[0][9 + c] = 0.0f;
        pc[1][c] = dc; pc[1][3 + c] = -dc; pc[1][6 + c] = 0.0f; pc[1][9 + c] = 0.0f;
        pc[2][c] = (-a1 - a0 + c1 + c0) * tc; pc[2][3 + c] = (a1 + a0) * tc; pc[2][6 + c] = -c0 * tc; pc[2][9 + c] = -c1 * tc;
        pc[3][c] = (-a1 + a0 + c1 - c0) * tc; pc[3][3 + c] = (a1 - a0) * tc; pc[3][6 + c] = c0 * tc; pc[3][9 + c] = -c1 * tc;
    }
    float gg[4], g23 = 0.0f;
    for (int i = 0; i < 4; ++i) { float s = 0.0f; for (int m = 0; m < 12; ++m) s += pc[i][m] * pc[i][m]; gg[i] = s; }
    for (int m = 0; m < 12; ++m) g23 += pc[2][m] * pc[3][m];
    const float det23 = gg[2] * gg[3] - g23 * g23;
    if (gg[0] > 1e-5f && gg[1] > 1e-5f && det23 > 1e-6f * gg[2] * gg[3]) {
        const float we = -g * Sgeom;
        float E[4][4] = {{0}}, Li[4][4] = {{0}}, LiE[4][4], Ss[4][4], LtS[4][4], Core[4][4];
        E[0][0] = we * (-k0 * gg[0]); E[1][1] = we * (k0 * gg[1]);
        E[0][2] = E[2][0] = we * (k1 * gg[2]); E[0][3] = E[3][0] = we * (k1 * g23);
        E[1][2] = E[2][1] = we * (k1 * g23); E[1][3] = E[3][1] = we * (k1 * gg[3]);
        Li[0][0] = 1.0f / sqrtf(gg[0]); Li[1][1] = 1.0f / sqrtf(gg[1]);
        const float L22 = sqrtf(gg[2]), L32 = g23 / L22, L33 = sqrtf(fmaxf(1e-30f, gg[3] - L32 * L32));
        Li[2][2] = 1.0f / L22; Li[3][3] = 1.0f / L33; Li[3][2] = -L32 / (L22 * L33);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.0f; for (int k = 0; k < 4; ++k) s += Li[i][k] * E[k][j]; LiE[i][j] = s; }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.0f; for (int k = 0; k < 4; ++k) s += LiE[i][k] * Li[j][k]; Ss[i][j] = s; }
        o_psd4(Ss);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.0f; for (int k = 0; k < 4; ++k) s += Li[k][i] * Ss[k][j]; LtS[i][j] = s; }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.0f; for (int k = 0; k < 4; ++k) s += LtS[i][k] * Li[k][j]; Core[i][j] = s; }
        for (int m = 0; m < 12; ++m) {
            float PCm[4];
            for (int j = 0; j < 4; ++j) { float s = 0.0f; for (int i = 0; i < 4; ++i) s += pc[i][m] * Core[i][j]; PCm[j] = s; }
            for (int n = 0; n < 12; ++n) { float s = 0.0f; for (int j = 0; j < 4; ++j) s += PCm[j] * pc[j][n]; Hp[n * 12 + m] += s; }
        }
    }
    const int perm[4] = {2, 1, 0, 3};
    for (int a2 = 0; a2 < 4; ++a2)
        for (int b2 = 0; b2 < 4; ++b2)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H[(3 * perm[b2] + j) * 12 + 3 * perm[a2] + i] = Hp[(3 * b2 + j) * 12 + 3 * a2 + i];
}


/* ---- strain limiting (strainlimiting/strainlimiting.cu:18-120, 250-307; utility/utility.cu:60-150, 185-250;
 *      barrier/barrier.cu:116-131; eigenanalysis/eigenanalysis.cu:11-85; linalg/eigsolve.hpp:40-94) ---- */
typedef struct { float U[3][2], S[2], Vt[2][2]; } o_svd32;
static inline void o_jacobi_svd32(const float F[3][2], o_svd32 *o) { /* one-sided Jacobi, one sweep, ascending singular values */
    float A[3][2], V[2][2] = {{1.0f, 0.0f}, {0.0f, 1.0f}};
    float a = 0.0f, b = 0.0f, g = 0.0f;
    for (int r = 0; r < 3; ++r) { A[r][0] = F[r][0]; A[r][1] = F[r][1]; }
    for (int r = 0; r < 3; ++r) { a = a + A[r][0] * A[r][0]; b = b + A[r][1] * A[r][1]; g = g + A[r][0] * A[r][1]; }
    const float zeta = (b - a) / (2.0f * g);
    const float t = (g == 0.0f) ? 0.0f : (zeta >= 0.0f ? 1.0f : -1.0f) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
    const float c = 1.0f / sqrtf(1.0f + t * t), s = c * t;
    for (int r = 0; r < 3; ++r) { const float p = A[r][0], q = A[r][1]; A[r][0] = c * p - s * q; A[r][1] = s * p + c * q; }
    for (int r = 0; r < 2; ++r) { const float p = V[r][0], q = V[r][1]; V[r][0] = c * p - s * q; V[r][1] = s * p + c * q; }
    float sv[2], u[3][2];
    for (int k = 0; k < 2; ++k) {
        float n2 = 0.0f;
        for (int r = 0; r < 3; ++r) n2 = n2 + A[r][k] * A[r][k];
        sv[k] = sqrtf(n2);
        const float inv = sv[k] > 0.0f ? 1.0f / sv[k] : 0.0f;
        for (int r = 0; r < 3; ++r) u[r][k] = A[r][k] * inv;
    }
    const int i0 = sv[1] < sv[0] ? 1 : 0, idx[2] = {i0, 1 - i0};
    for (int k = 0; k < 2; ++k) {
        o->S[k] = sv[idx[k]];
        for (int r = 0; r < 3; ++r) o->U[r][k] = u[r][idx[k]];
        for (int r = 0; r < 2; ++r) o->Vt[k][r] = V[r][idx[k]];
    }
}
static inline void o_defgrad32(const float *x0, const float *x1, const float *x2, const float *ir, float F[3][2]) {
    for (int r = 0; r < 3; ++r) {
        const float e0 = x1[r] - x0[r], e1 = x2[r] - x0[r];
        F[r][0] = e0 * ir[0] + e1 * ir[1];
        F[r][1] = e0 * ir[2] + e1 * ir[3];
    }
}
static inline float o_max_strain_at(const float F0[3][2], const float dF[3][2], float t) {
    float F[3][2];
    o_svd32 svd;
    for (int r = 0; r < 3; ++r) { F[r][0] = F0[r][0] + t * dF[r][0]; F[r][1] = F0[r][1] + t * dF[r][1]; }
    o_jacobi_svd32(F, &svd);
    const float s0 = svd.S[0] - 1.0f, s1 = svd.S[1] - 1.0f;
    return s0 < s1 ? s1 : s0;
}
/* xa/xb: face vertex positions at the start / end of the candidate step (3 x 3) */
static inline float o_strain_face_toi(const float (*xa)[3], const float (*xb)[3], const float *ir, float limit, float max_t) {
    float F0[3][2], F1[3][2], dF[3][2];
    o_defgrad32(xa[0], xa[1], xa[2], ir, F0);
    o_defgrad32(xb[0], xb[1], xb[2], ir, F1);
    for (int r = 0; r < 3; ++r) { dF[r][0] = F1[r][0] - F0[r][0]; dF[r][1] = F1[r][1] - F0[r][1]; }
    float t = max_t;
    if (o_max_strain_at(F0, dF, t) >= limit) {
        float hi = t, lo = 0.0f, window = hi - lo;
        for (;;) {
            t = 0.5f * (hi + lo);
            if (o_max_strain_at(F0, dF, t) - limit < 0.0f) lo = t; else hi = t;
            const float nw = hi - lo;
            if (nw == window) break;
            window = nw;
        }
        t = lo;
    }
    return t;
}
static inline void o_symm2x2(float a00, float a01, float a11, float *val, float vec[2][2]) {
    const float scale = fmaxf(fmaxf(fabsf(a00), fabsf(a01)), fabsf(a11));
    if (scale <= 0.0f) { val[0] = val[1] = 0.0f; vec[0][0] = vec[1][1] = 1.0f; vec[0][1] = vec[1][0] = 0.0f; return; }
    const float b00 = a00 / scale, b01 = a01 / scale, b11 = a11 / scale;
    const float d = b00 - b11, D = 0.5f * sqrtf(d * d + 4.0f * b01 * b01), mid = 0.5f * (b00 + b11);
    const float l0 = mid - D, l1 = mid + D;
    const float r0[2] = {b01, -(b00 - l0)}, r1[2] = {b11 - l0, -b01};
    const float n0 = r0[0] * r0[0] + r0[1] * r0[1], n1 = r1[0] * r1[0] + r1[1] * r1[1];
    float u[2] = {1.0f, -0.0f};
    if (n0 > 1e-14f) { const float n = sqrtf(n0); u[0] = r0[0] / n; u[1] = r0[1] / n; }
    else if (n1 > 1e-14f) { const float n = sqrtf(n1); u[0] = r1[0] / n; u[1] = r1[1] / n; }
    val[0] = scale * l0; val[1] = scale * l1;
    vec[0][0] = u[0]; vec[1][0] = u[1]; vec[0][1] = u[1]; vec[1][1] = -u[0];
}
/* One face.  xe: evaluation positions, xc: start-of-step positions, Hel[a][b][9]: elastic-only blocks (column-major).
 * Outputs are already scaled by the stiffness; returns 0 when the face is not stretched. */
static int o_barrier_kind = 0; /* ParamSet.barrier of the running scene (set by the oracle at initialize; 0 = cubic) */
static inline int o_strain_face(const float (*xe)[3], const float (*xc)[3], const float *ir, float strainlimit, float limit, float mass,
                                const float Hel[3][3][9], float eig_eps, float f[3][3], float H[3][3][9]) {
    float F[3][2];
    o_svd32 svd;
    o_defgrad32(xe[0], xe[1], xe[2], ir, F);
    o_jacobi_svd32(F, &svd);
    float S[2] = {svd.S[0] - 1.0f, svd.S[1] - 1.0f};
    const float smax = S[0] < S[1] ? S[1] : S[0];
    if (!(smax > 0.0f)) return 0;
    float deda[2] = {0.0f, 0.0f}, d2[2] = {0.0f, 0.0f};
    for (int i = 0; i < 2; ++i)
        if (S[i] > 0.0f) {
            const float y = strainlimit - S[i];
            deda[i] = -o_barrier_gradient(y, strainlimit, 0.0f, o_barrier_kind);
            d2[i] = o_barrier_curvature(y, strainlimit, 0.0f, o_barrier_kind);
        }
    const float third = 1.0f / 3.0f;
    float w[3][3], wHw = 0.0f;
    for (int k = 0; k < 3; ++k) {
        const float c = third * xc[0][k] + third * xc[1][k] + third * xc[2][k];
        for (int a = 0; a < 3; ++a) w[a][k] = xc[a][k] - c;
    }
    for (int a = 0; a < 3; ++a) {
        float row[3] = {0.0f, 0.0f, 0.0f};
        for (int b = 0; b < 3; ++b)
            for (int r = 0; r < 3; ++r) row[r] += Hel[a][b][r] * w[b][0] + Hel[a][b][3 + r] * w[b][1] + Hel[a][b][6 + r] * w[b][2];
        wHw += w[a][0] * row[0] + w[a][1] * row[1] + w[a][2] * row[2];
    }
    const float dgap = limit - smax, stiffness = wHw + mass / (dgap * dgap);
    S[0] += 1.0f; S[1] += 1.0f;
    float dedF[3][2];
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 2; ++j) dedF[r][j] = (svd.U[r][0] * deda[0]) * svd.Vt[0][j] + (svd.U[r][1] * deda[1]) * svd.Vt[1][j];
    float ev[2], evec[2][2];
    o_symm2x2(d2[0], 0.0f, d2[1], ev, evec);
    const float is2 = 0.7071067811865475f;
    float U3[3][3], Q[6][3][2];
    for (int r = 0; r < 3; ++r) { U3[r][0] = svd.U[r][0]; U3[r][1] = svd.U[r][1]; }
    U3[0][2] = svd.U[1][0] * svd.U[2][1] - svd.U[2][0] * svd.U[1][1];
    U3[1][2] = svd.U[2][0] * svd.U[0][1] - svd.U[0][0] * svd.U[2][1];
    U3[2][2] = svd.U[0][0] * svd.U[1][1] - svd.U[1][0] * svd.U[0][1];
    memset(Q, 0, sizeof(Q));
    Q[0][0][1] = is2; Q[0][1][0] = -is2; Q[1][0][1] = is2; Q[1][1][0] = is2; Q[2][2][0] = 1.0f; Q[3][2][1] = 1.0f;
    for (int i = 0; i < 2; ++i) { Q[4][i][i] = evec[i][0]; Q[5][i][i] = evec[i][1]; }
    const float a0 = S[0], a1 = S[1], den = a0 - a1;
    float lam[6];
    lam[0] = fmaxf(0.0f, (deda[0] + deda[1]) / (a0 + a1));
    lam[1] = fmaxf(0.0f, fabsf(den) > eig_eps ? (deda[0] - deda[1]) / den : 0.5f * (d2[0] + d2[1]));
    lam[2] = fmaxf(0.0f, deda[0] / a0);
    lam[3] = fmaxf(0.0f, deda[1] / a1);
    lam[4] = fmaxf(0.0f, ev[0]);
    lam[5] = fmaxf(0.0f, ev[1]);
    float d2F[6][6];
    memset(d2F, 0, sizeof(d2F));
    for (int i = 0; i < 6; ++i) {
        if (!lam[i]) continue;
        float q[6];
        for (int j = 0; j < 2; ++j)
            for (int r = 0; r < 3; ++r) {
                const float t0 = U3[r][0] * Q[i][0][0] + U3[r][1] * Q[i][1][0] + U3[r][2] * Q[i][2][0];
                const float t1 = U3[r][0] * Q[i][0][1] + U3[r][1] * Q[i][1][1] + U3[r][2] * Q[i][2][1];
                q[3 * j + r] = t0 * svd.Vt[0][j] + t1 * svd.Vt[1][j];
            }
        for (int x = 0; x < 6; ++x) for (int y = 0; y < 6; ++y) d2F[x][y] += lam[i] * q[x] * q[y];
    }
    const float g[3][2] = {{-ir[0] - ir[1], -ir[2] - ir[3]}, {ir[0], ir[2]}, {ir[1], ir[3]}};
    for (int a = 0; a < 3; ++a)
        for (int dim = 0; dim < 3; ++dim) f[a][dim] = stiffness * (g[a][0] * dedF[dim][0] + g[a][1] * dedF[dim][1]);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) {
                    float v = 0.0f;
                    for (int dd = 0; dd < 2; ++dd) for (int ee = 0; ee < 2; ++ee) v += (g[a][dd] * g[b][ee]) * d2F[3 * dd + r][3 * ee + c];
                    H[a][b][c * 3 + r] = stiffness * v;
                }
    return 1;
}

/* ---- rods (yarn): Hookean stretch and interior-vertex bending ---------------------------------------- */
/* hook::make_diff_table (energy/model/hook.hpp:20-37): gradient per end point and the block K of the
 * 6x6 Hessian [[K,-K],[-K,K]] = weight (c0 g g^T + c1 dtdx^T dtdx), g = [-n; n], dtdx = [-I, I] */
static inline void o_rod_hook(const float *x0, const float *x1, float l0, float weight, float f[2][3], float *K /* 3x3 column-major */) {
    float t[3];
    o_sub(x1, x0, t);
    const float l = o_norm3(t);
    const float n[3] = {t[0] / l, t[1] / l, t[2] / l};
    const float s = l / l0 - 1.0f;
    const float r = (l - l0) / l;
    const float c0 = fmaxf(0.0f, 1.0f - r) / l0;
    const float c1 = fmaxf(0.0f, r / l0);
    for (int k = 0; k < 3; ++k) { const float d = s * n[k]; f[0][k] = (-weight) * d; f[1][k] = weight * d; }
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) K[c * 3 + rr] = weight * ((c0 * n[rr]) * n[c] + (rr == c ? c1 : 0.0f));
}
/* dihedral_angle::gradient_theta (energy/model/dihedral_angle.hpp:397-425) */
static inline void o_strand_gradient_theta(const float *e0, const float *e1, float *g0, float *g1) {
    float cr[3];
    o_cross(e0, e1, cr);
    const float cs = o_dot(cr, cr), a = o_dot(e0, e0), b = o_dot(e1, e1);
    if (cs <= 1e-12f * a * b) { for (int k = 0; k < 3; ++k) g0[k] = g1[k] = 0.0f; return; }
    const float sq = sqrtf(cs);
    const float n[3] = {cr[0] / sq, cr[1] / sq, cr[2] / sq};
    float p0[3], p1[3];
    o_cross(e0, n, p0);
    o_cross(e1, n, p1);
    for (int k = 0; k < 3; ++k) { g0[k] = p0[k] / a; g1[k] = (-p1[k]) / b; }
}
static inline void o_strand_accum(float *H, float lam, const float *qa, const float *qb) {
    if (lam <= 0.0f) return;
    const float n2 = o_dot(qa, qa) + o_dot(qb, qb);
    if (n2 <= 1e-20f) return;
    float p9[9];
    for (int c = 0; c < 3; ++c) { p9[c] = qa[c]; p9[3 + c] = -qa[c] - qb[c]; p9[6 + c] = qb[c]; }
    const float s = lam / n2;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) H[j * 9 + i] = H[j * 9 + i] + s * p9[i] * p9[j];
}
/* dihedral_angle::strand_compute_force_hessian (dihedral_angle.hpp:451-541): E = 1/2 (theta - theta0)^2 over the
 * polyline (x0, x1, x2); exact gradient, Hessian projected to PSD through the closed-form eigenpairs
 * (four in-plane, two out-of-plane) of the angle energy in the space of the two segment vectors. H: 9x9 column-major */
static inline void o_strand(const float *x0, const float *x1, const float *x2, float rest_angle, float f[3][3], float *H) {
    float e0[3], e1[3], bcr[3], g0[3], g1[3];
    o_sub(x0, x1, e0);
    o_sub(x2, x1, e1);
    o_cross(e0, e1, bcr);
    const float theta = atan2f(o_norm3(bcr), o_dot(e0, e1));
    o_strand_gradient_theta(e0, e1, g0, g1);
    const float w = theta - rest_angle;
    for (int k = 0; k < 3; ++k) { f[0][k] = w * g0[k]; f[1][k] = w * ((-g0[k]) - g1[k]); f[2][k] = w * g1[k]; }
    for (int i = 0; i < 81; ++i) H[i] = 0.0f;
    const float l0 = o_norm3(e0), l1 = o_norm3(e1), lb = o_norm3(bcr);
    if (lb <= 1e-9f * l0 * l1 || l0 <= 1e-9f || l1 <= 1e-9f) return;
    const float tb[3] = {bcr[0] / lb, bcr[1] / lb, bcr[2] / lb};
    float e0p[3], e1p[3];
    o_cross(e0, tb, e0p);
    o_cross(e1, tb, e1p);
    const float cos_t = o_dot(e0, e1) / (l0 * l1);
    const float sin_t = lb / (l0 * l1);
    const float gamma = (l1 * l1) / (l0 * l0);
    const float gm1 = gamma - 1.0f, gp1 = gamma + 1.0f;
    const float r = sqrtf(4.0f * w * w * (gm1 / gp1) * (gm1 / gp1) + 1.0f);
    const float Rm = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f - r)));
    const float Rp = sqrtf(fmaxf(0.0f, 2.0f * (2.0f * w * w + 1.0f + r)));
    const float Lh = 1.0f / (l0 * l0) + 1.0f / (l1 * l1);
    const float cc = 4.0f * w / gp1;
    const int sr[4] = {-1, -1, 1, 1}, sR[4] = {-1, 1, -1, 1};
    const float RR[4] = {Rm, Rm, Rp, Rp};
    for (int m = 0; m < 4; ++m) {
        const float R = RR[m];
        const float a = w * (gm1 * gm1 / gp1 + sr[m] * gp1 * r + sR[m] * gm1 * R);
        const float b = -0.5f * gamma * R * R + 2.0f * w * w - sR[m] * 0.5f * (gm1 + sr[m] * gp1 * r) * R;
        const float d = 1.0f + sr[m] * r + sR[m] * R;
        const float lam = 0.25f * Lh * d;
        float qa[3], qb[3];
        for (int k = 0; k < 3; ++k) { qa[k] = a * e0[k] + b * e0p[k]; qb[k] = cc * e1[k] + d * e1p[k]; }
        o_strand_accum(H, lam, qa, qb);
    }
    const float beta = (l1 / l0 - l0 / l1) * cos_t;
    const float alpha0 = 0.5f * (-beta + sqrtf(beta * beta + 4.0f));
    const float ginv = w / sin_t;
    const float lam4 = ginv * (cos_t / (l1 * l1) - alpha0 / (l0 * l1));
    const float lam5 = ginv * (cos_t / (l1 * l1) + 1.0f / (alpha0 * l0 * l1));
    float qa[3], qb[3];
    for (int k = 0; k < 3; ++k) { qa[k] = alpha0 * tb[k]; qb[k] = tb[k]; }
    o_strand_accum(H, lam4, qa, qb);
    for (int k = 0; k < 3; ++k) { qa[k] = tb[k]; qb[k] = (-alpha0) * tb[k]; }
    o_strand_accum(H, lam5, qa, qb);
}
/* coefficient of the rod-bend energy at an interior vertex (energy/energy.cu:572-603) */
static inline float o_rod_bend_coefficient(float bend0, float bend1, float len0, float len1, float mass) {
    const float bend = 0.5f * (bend0 + bend1);
    const float REF_LENGTH = 1e-2f;
    const float voronoi = 0.5f * (len0 + len1);
    return voronoi > 0.0f ? bend * mass * (REF_LENGTH * REF_LENGTH) / (voronoi * voronoi) : 0.0f;
}

/* ---- rod strain limiting (strainlimiting/strainlimiting.cu:100-190, 192-245) ---- */
static inline void o_mtv3(const float *M, const float *x, float *y) { /* y = M^T x, M column-major */
    for (int c = 0; c < 3; ++c) y[c] = M[c * 3 + 0] * x[0] + M[c * 3 + 1] * x[1] + M[c * 3 + 2] * x[2];
}
/* barrier on the stretch g = l / l0 - 1 beyond the DRAWN length l0, gap = limit; Hessian [[K,-K],[-K,K]];
 * Hel00 / Hel01 / Hel11: blocks (e0,e0), (e0,e1), (e1,e1) of the elastic-only matrix; c0, c1: start-of-step positions */
static inline int o_rod_strain(const float *x0, const float *x1, const float *c0, const float *c1, float l0, float limit, float mass,
                               const float *Hel00, const float *Hel01, const float *Hel11, float f[2][3], float *K) {
    float d[3];
    o_sub(x1, x0, d);
    const float l = o_norm3(d);
    const float g = l / l0 - 1.0f;
    if (!(g > 0.0f)) return 0;
    const float y = limit - g;
    const float dEdg = -o_barrier_gradient(y, limit, 0.0f, o_barrier_kind);
    const float d2Edg2 = o_barrier_curvature(y, limit, 0.0f, o_barrier_kind);
    const float n[3] = {d[0] / l, d[1] / l, d[2] / l};
    const float inv_ll0 = 1.0f / (l * l0);
    const float a_psd = fmaxf(0.0f, d2Edg2), b_psd = fmaxf(0.0f, dEdg);
    float w0[3], w1[3], y0[3], y1[3], t0[3], t1[3];
    for (int k = 0; k < 3; ++k) { const float ce = 0.5f * c0[k] + 0.5f * c1[k]; w0[k] = c0[k] - ce; w1[k] = c1[k] - ce; }
    o_mv3(Hel00, w0, t0); o_mv3(Hel01, w1, t1);
    for (int k = 0; k < 3; ++k) y0[k] = t0[k] + t1[k];
    o_mtv3(Hel01, w0, t0); o_mv3(Hel11, w1, t1);
    for (int k = 0; k < 3; ++k) y1[k] = t0[k] + t1[k];
    const float dd = limit - g;
    const float stiffness = (o_dot(w0, y0) + o_dot(w1, y1)) + mass / (dd * dd);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) {
            const float P = (r == c ? 1.0f : 0.0f) - n[r] * n[c];
            K[c * 3 + r] = stiffness * (a_psd * ((n[r] / l0) * (n[c] / l0)) + b_psd * (inv_ll0 * P));
        }
    for (int k = 0; k < 3; ++k) { const float j1 = n[k] / l0; f[0][k] = stiffness * (dEdg * (-j1)); f[1][k] = stiffness * (dEdg * j1); }
    return 1;
}
static inline float o_rod_strain_toi(const float *d0, const float *d1, float l0, float limit, float max_t) {
    float t = max_t, dd[3], p[3];
    o_sub(d1, d0, dd);
    for (int k = 0; k < 3; ++k) p[k] = d0[k] + t * dd[k];
    if (o_norm3(p) / l0 - 1.0f >= limit) {
        float upper_t = t, lower_t = 0.0f, window = upper_t - lower_t;
        for (;;) {
            t = 0.5f * (upper_t + lower_t);
            for (int k = 0; k < 3; ++k) p[k] = d0[k] + t * dd[k];
            if ((o_norm3(p) / l0 - 1.0f) - limit < 0.0f) lower_t = t;
            else upper_t = t;
            const float nw = upper_t - lower_t;
            if (nw == window) break;
            window = nw;
        }
        t = lower_t;
    }
    return t;
}

/* ---- isotropic shell models ARAP / StVK / SNHk (energy.cu:337-352; arap.hpp:36-49, stvk.hpp:38-72 + detsqr.hpp:22-33,
 * snhk.hpp:45-56; eigenanalysis.cu:14-85).  model: 0 ARAP, 1 StVK, 3 SNHk.  d2 = (d00, d01, d11). ---- */
static inline void o_iso_table2(int model, const float *a, float mu, float lmd, float *deda, float *d2) {
    if (model == 0) {
        const float tr = (a[0] - 1.0f) + (a[1] - 1.0f);
        deda[0] = 2.0f * mu * (a[0] - 1.0f) + lmd * tr * 1.0f;
        deda[1] = 2.0f * mu * (a[1] - 1.0f) + lmd * tr * 1.0f;
        d2[0] = 2.0f * mu * 1.0f + lmd; d2[2] = 2.0f * mu * 1.0f + lmd; d2[1] = 2.0f * mu * 0.0f + lmd;
    } else if (model == 1) {
        const float a0 = a[0] - 1.0f, a1 = a[1] - 1.0f, J = a[0] * a[1];
        deda[0] = mu * (a0 * a0 * a0) + lmd * a[1] * (J - 1.0f);
        deda[1] = mu * (a1 * a1 * a1) + lmd * a[0] * (J - 1.0f);
        d2[0] = mu * (3.0f * a0 * a0) + lmd * (a[1] * a[1]);
        d2[2] = mu * (3.0f * a1 * a1) + lmd * (a[0] * a[0]);
        d2[1] = mu * 0.0f + lmd * (2.0f * a[0] * a[1] - 1.0f);
    } else {
        const float x = a[0], y = a[1];
        deda[0] = lmd * y * (x * y - 1.0f) + mu * (x - y);
        deda[1] = lmd * x * (x * y - 1.0f) + mu * (y - x);
        d2[0] = lmd * y * y + mu; d2[2] = lmd * x * x + mu; d2[1] = 2.0f * lmd * x * y - lmd - mu;
    }
}
static inline void o_iso_face(const float *x0, const float *x1, const float *x2, const float *ir, int model, float mu, float lmd, float mass,
                              float eig_eps, float f[3][3], float H[3][3][9]) {
    float F[3][2], deda[2], d2[3], dedF[3][2], ev[2], evec[2][2];
    o_svd32 svd;
    o_defgrad32(x0, x1, x2, ir, F);
    o_jacobi_svd32(F, &svd);
    o_iso_table2(model, svd.S, mu, lmd, deda, d2);
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 2; ++j) dedF[r][j] = (svd.U[r][0] * deda[0]) * svd.Vt[0][j] + (svd.U[r][1] * deda[1]) * svd.Vt[1][j];
    o_symm2x2(d2[0], d2[1], d2[2], ev, evec);
    const float is2 = 0.7071067811865475f;
    float U3[3][3], Q[6][3][2];
    for (int r = 0; r < 3; ++r) { U3[r][0] = svd.U[r][0]; U3[r][1] = svd.U[r][1]; }
    U3[0][2] = svd.U[1][0] * svd.U[2][1] - svd.U[2][0] * svd.U[1][1];
    U3[1][2] = svd.U[2][0] * svd.U[0][1] - svd.U[0][0] * svd.U[2][1];
    U3[2][2] = svd.U[0][0] * svd.U[1][1] - svd.U[1][0] * svd.U[0][1];
    memset(Q, 0, sizeof(Q));
    Q[0][0][1] = is2; Q[0][1][0] = -is2; Q[1][0][1] = is2; Q[1][1][0] = is2; Q[2][2][0] = 1.0f; Q[3][2][1] = 1.0f;
    for (int i = 0; i < 2; ++i) { Q[4][i][i] = evec[i][0]; Q[5][i][i] = evec[i][1]; }
    const float a0 = svd.S[0], a1 = svd.S[1], den = a0 - a1;
    float lam[6];
    lam[0] = fmaxf(0.0f, (deda[0] + deda[1]) / (a0 + a1));
    lam[1] = fmaxf(0.0f, fabsf(den) > eig_eps ? (deda[0] - deda[1]) / den : 0.5f * (d2[0] + d2[2]) - 0.5f * (d2[1] + d2[1]));
    lam[2] = fmaxf(0.0f, deda[0] / a0);
    lam[3] = fmaxf(0.0f, deda[1] / a1);
    lam[4] = fmaxf(0.0f, ev[0]);
    lam[5] = fmaxf(0.0f, ev[1]);
    float d2F[6][6];
    memset(d2F, 0, sizeof(d2F));
    for (int i = 0; i < 6; ++i) {
        if (!lam[i]) continue;
        float T[3][2], M[3][2], q[6];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 2; ++j) T[r][j] = U3[r][0] * Q[i][0][j] + U3[r][1] * Q[i][1][j] + U3[r][2] * Q[i][2][j];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 2; ++j) M[r][j] = T[r][0] * svd.Vt[0][j] + T[r][1] * svd.Vt[1][j];
        for (int j = 0; j < 2; ++j) for (int r = 0; r < 3; ++r) q[3 * j + r] = M[r][j];
        for (int x = 0; x < 6; ++x) for (int y = 0; y < 6; ++y) d2F[x][y] += lam[i] * q[x] * q[y];
    }
    const float g[3][2] = {{-ir[0] - ir[1], -ir[2] - ir[3]}, {ir[0], ir[2]}, {ir[1], ir[3]}};
    for (int a = 0; a < 3; ++a)
        for (int dim = 0; dim < 3; ++dim) f[a][dim] = mass * (g[a][0] * dedF[dim][0] + g[a][1] * dedF[dim][1]);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            float blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int dd = 0; dd < 2; ++dd)
                for (int ee = 0; ee < 2; ++ee) {
                    const float sc = g[a][dd] * g[b][ee];
                    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) blk[c * 3 + r] += sc * d2F[3 * dd + r][3 * ee + c];
                }
            for (int e = 0; e < 9; ++e) H[a][b][e] = mass * blk[e];
        }
}

/* ---- stiffness-proportional (Rayleigh) damping of one element: energy.cu:88-180 ----
 * K: dense (3n)x(3n) column-major element Hessian, u: x - x^n per vertex, f: element gradient.  s = beta / dt.
 * plain form: f += s K u, K *= (1 + s);  lagged form (bending): f += s K_lag u, K += s K_lag */
static inline void o_damping_gradient(int n, const float *K, const float (*u)[3], float s, float (*f)[3]) {
    for (int a = 0; a < n; ++a) {
        float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;
        for (int b = 0; b < 3 * n; ++b) {
            const float ub = u[b / 3][b % 3];
            f0 += K[b * 3 * n + 3 * a + 0] * ub;
            f1 += K[b * 3 * n + 3 * a + 1] * ub;
            f2 += K[b * 3 * n + 3 * a + 2] * ub;
        }
        f[a][0] += s * f0; f[a][1] += s * f1; f[a][2] += s * f2;
    }
}
static inline void o_stiffness_damping(int n, float *K, const float (*u)[3], float beta, float dt, float (*f)[3]) {
    if (beta <= 0.0f || dt <= 0.0f) return;
    const float s = beta / dt;
    o_damping_gradient(n, K, u, s, f);
    for (int i = 0; i < 9 * n * n; ++i) K[i] *= (1.0f + s);
}
static inline void o_stiffness_damping_lagged(int n, float *K, const float *K_lag, const float (*u)[3], float beta, float dt, float (*f)[3]) {
    if (beta <= 0.0f || dt <= 0.0f) return;
    const float s = beta / dt;
    o_damping_gradient(n, K_lag, u, s, f);
    for (int i = 0; i < 9 * n * n; ++i) K[i] += s * K_lag[i];
}

/* ---- stitches: seam springs between two barycentric points (energy.cu:796-886) ----
 * x[6][3]: the six slots (source 0..2, target 3..5), ghat / offset: their vertices' contact gap and offset, w: weights.
 * f[6][3] gradient, H 18x18 column-major, both scaled by the stitch stiffness. */
static inline void o_stitch(const float (*x)[3], const float *ghat, const float *offset, const float *w, float stiffness, float length_factor,
                            float f[6][3], float *H) {
    const float l0 = (w[0] * ghat[0] + w[1] * ghat[1] + w[2] * ghat[2] + w[3] * ghat[3] + w[4] * ghat[4] + w[5] * ghat[5]) / 2.0f;
    const float so = fmaxf(fmaxf(offset[0], offset[1]), offset[2]), to = fmaxf(fmaxf(offset[3], offset[4]), offset[5]);
    const float l_cap = length_factor * l0 + so + to;
    const float s = 1.0f / 6.0f;
    float cog[3], z0[3], z1[3], t[3], n[3];
    for (int k = 0; k < 3; ++k) {
        cog[k] = s * x[0][k] + s * x[1][k] + s * x[2][k] + s * x[3][k] + s * x[4][k] + s * x[5][k];
        z0[k] = w[0] * (x[0][k] - cog[k]) + w[1] * (x[1][k] - cog[k]) + w[2] * (x[2][k] - cog[k]);
        z1[k] = w[3] * (x[3][k] - cog[k]) + w[4] * (x[4][k] - cog[k]) + w[5] * (x[5][k] - cog[k]);
        t[k] = z0[k] - z1[k];
    }
    const float l = fminf(l_cap, o_norm3(t));
    for (int k = 0; k < 3; ++k) n[k] = t[k] / l;
    const float sw[6] = {w[0], w[1], w[2], -w[3], -w[4], -w[5]};
    const float e = l / l0 - 1.0f;
    float g[18];
    for (int a = 0; a < 6; ++a) for (int k = 0; k < 3; ++k) g[3 * a + k] = sw[a] * n[k];
    const float r = (l - l0) / l;
    const float c0 = fmaxf(0.0f, 1.0f - r) / l0, c1 = fmaxf(0.0f, r / l0);
    for (int a = 0; a < 6; ++a) for (int k = 0; k < 3; ++k) f[a][k] = stiffness * (sw[a] * (e * n[k]));
    for (int j = 0; j < 18; ++j)
        for (int i = 0; i < 18; ++i) {
            const float dd = (i % 3 == j % 3) ? (c1 * sw[i / 3]) * sw[j / 3] : 0.0f; /* (c1 dtdx^T) dtdx */
            H[j * 18 + i] = stiffness * ((c0 * g[i]) * g[j] + dd);
        }
}

/* ---- plastic creep of the rest shape after a step (plasticity/plasticity.cu:12-86, 147-245) ---- */
static inline float o_plastic_alpha(float rate, float dt) { return 1.0f - expf(-rate * dt); }
static inline float o_dead_zone_creep(float delta, float threshold, float alpha) { return alpha * (delta - ((delta > 0.0f) ? threshold : -threshold)); }
static inline float o_hinge_angle(const float (*X)[3]) { /* dihedral_angle.hpp:81-96 with (v0,v1,v2,v3) = the remapped vertices */
    float a[3], b[3], n1[3], n2[3], ee[3], t[3];
    o_sub(X[1], X[0], a); o_sub(X[2], X[0], b); o_cross(a, b, n1);
    o_sub(X[2], X[3], a); o_sub(X[1], X[3], b); o_cross(a, b, n2);
    o_sub(X[1], X[2], ee);
    const float en = o_norm3(ee);
    if (!(en > 0.0f)) return 0.0f;
    o_cross(n2, n1, t);
    return atan2f(o_dot(t, ee) / en, o_dot(n1, n2));
}
/* ir: inverse rest 2x2 (column-major), updated in place; returns 1 when it changed */
static inline int o_face_plasticity(const float *x0, const float *x1, const float *x2, float *ir, float threshold, float alpha) {
    float F[3][2], Sn[2], Fn[3][2], d0[3], d1[3], c[3], n[3], t[3], b[3];
    o_svd32 svd;
    o_defgrad32(x0, x1, x2, ir, F);
    o_jacobi_svd32(F, &svd);
    int changed = 0;
    for (int k = 0; k < 2; ++k) {
        Sn[k] = svd.S[k];
        if (fabsf(svd.S[k] - 1.0f) > threshold) {
            const float target = (svd.S[k] < 1.0f) ? (1.0f - threshold) : (1.0f + threshold);
            Sn[k] = svd.S[k] + alpha * (target - svd.S[k]);
            changed = 1;
        }
    }
    if (!changed) return 0;
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 2; ++j) Fn[r][j] = (svd.U[r][0] * Sn[0]) * svd.Vt[0][j] + (svd.U[r][1] * Sn[1]) * svd.Vt[1][j];
    o_sub(x1, x0, d0); o_sub(x2, x0, d1);
    o_cross(d0, d1, c);
    float l = o_norm3(c); for (int k = 0; k < 3; ++k) n[k] = l > 0.0f ? c[k] / l : c[k];
    l = o_norm3(d0); for (int k = 0; k < 3; ++k) t[k] = l > 0.0f ? d0[k] / l : d0[k];
    o_cross(n, t, c);
    l = o_norm3(c); for (int k = 0; k < 3; ++k) b[k] = l > 0.0f ? c[k] / l : c[k];
    const float m00 = o_dot(t, d0), m01 = o_dot(t, d1), m10 = o_dot(b, d0), m11 = o_dot(b, d1);
    const float det = m00 * m11 - m01 * m10;
    const float i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
    float PF[2][2];
    for (int j = 0; j < 2; ++j) { const float col[3] = {Fn[0][j], Fn[1][j], Fn[2][j]}; PF[0][j] = o_dot(t, col); PF[1][j] = o_dot(b, col); }
    ir[0] = i00 * PF[0][0] + i01 * PF[1][0]; ir[2] = i00 * PF[0][1] + i01 * PF[1][1];
    ir[1] = i10 * PF[0][0] + i11 * PF[1][0]; ir[3] = i10 * PF[0][1] + i11 * PF[1][1];
    return 1;
}
static inline float o_hinge_plasticity(const float (*X)[3], float rest_angle, float threshold, float alpha) {
    const float delta = o_hinge_angle(X) - rest_angle;
    return fabsf(delta) > threshold ? rest_angle + o_dead_zone_creep(delta, threshold, alpha) : rest_angle;
}
static inline float o_rod_bend_plasticity(const float *xj, const float *xi, const float *xk, float rest_angle, float threshold, float alpha) {
    float e0[3], e1[3];
    o_sub(xj, xi, e0); o_sub(xk, xi, e1);
    const float n0 = o_norm3(e0), n1 = o_norm3(e1);
    if (n0 <= 0.0f || n1 <= 0.0f) return rest_angle;
    const float delta = acosf(fmaxf(-1.0f, fminf(1.0f, o_dot(e0, e1) / (n0 * n1)))) - rest_angle;
    return fabsf(delta) > threshold ? rest_angle + o_dead_zone_creep(delta, threshold, alpha) : rest_angle;
}

#endif

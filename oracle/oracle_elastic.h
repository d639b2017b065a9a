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

#endif

// math_inflate.cuh -- inflation pressure on shell faces: E = -(P / 6) x0 . (x1 x x2) per face (the signed volume of the
// tetrahedron the face spans with the origin, summed over a closed surface = -P * volume), its gradient in the
// cancellation-free form and its analytically PSD-projected 9x9 Hessian (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/energy/model/inflate.hpp:16-150 and the one-sided Jacobi SVD it calls
// (utility/utility.cu:60-183: five sweeps for 3x3, singular values ascending, the reflection moved onto the smallest one).
// The eigen-decomposition of the 3x3 "diagonal subspace" matrix uses a cyclic Jacobi here instead of the reference's closed
// form (linalg/eigsolve.hpp:195-232): the projected Hessian is the same matrix up to rounding.
#pragma once
#include "vecmath.cuh"

namespace ppf {

// F = U diag(S) V^T, S ascending, columns of U / V orthonormal
PPF_HD void jacobi_svd3(const M3 &F, M3 &U, float *S, M3 &V) {
    float A[3][3], Vv[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { A[r][c] = F.m[c * 3 + r]; Vv[r][c] = r == c ? 1.0f : 0.0f; }
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
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (s_[idx[j]] < s_[idx[i]]) { const int tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp; }
    for (int c = 0; c < 3; ++c) {
        S[c] = s_[idx[c]];
        for (int r = 0; r < 3; ++r) { U.m[c * 3 + r] = u_[r][idx[c]]; V.m[c * 3 + r] = Vv[r][idx[c]]; }
    }
}
// rotation variant (utility.cu:167-183): a reflection in exactly one of U, V is moved onto the smallest singular value
PPF_HD void svd3_rv(const M3 &F, M3 &U, float *S, M3 &V) {
    jacobi_svd3(F, U, S, V);
    const float du = det3(U), dv = det3(V);
    int mi = 0;
    for (int i = 1; i < 3; ++i) if (S[i] < S[mi]) mi = i;
    if (du < 0.0f && dv > 0.0f) { for (int r = 0; r < 3; ++r) U.m[mi * 3 + r] = -U.m[mi * 3 + r]; S[mi] = -S[mi]; }
    else if (du > 0.0f && dv < 0.0f) { for (int r = 0; r < 3; ++r) V.m[mi * 3 + r] = -V.m[mi * 3 + r]; S[mi] = -S[mi]; }
}
// eigen-decomposition of a symmetric 3x3 (cyclic Jacobi, 8 sweeps); vec[k] = k-th eigenvector
PPF_HD void jacobi_eig3(float A[3][3], float *val, float vec[3][3]) {
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

// gradient per vertex, conditioned on the edge vectors e1 = x1 - x0, e2 = x2 - x0 (inflate.hpp:44-53)
PPF_HD void inflate_gradient(float pressure, const V3 &v0, const V3 &e1, const V3 &e2, V3 *g) {
    const float coeff = -pressure / 6.0f;
    g[0] = coeff * (cross(e1, e2) + cross(e1 - e2, v0));
    g[1] = coeff * cross(e2, v0);
    g[2] = coeff * cross(v0, e1);
}
// PSD-projected Hessian, 9x9 column-major (inflate.hpp:70-150): with X = [v0 v1 v2] = U diag(s) V^T the Hessian of det X has
// the "swap" eigenpairs +-s_k and the three eigenpairs of A = [[0,s2,s1],[s2,0,s0],[s1,s0,0]]; -(P/6) flips the signs, the
// modes with positive weight are kept (three swap modes for positive s, the negative eigenvalues of A)
PPF_HD void inflate_hessian(float pressure, const V3 &v0, const V3 &v1, const V3 &v2, float *H) {
    M3 X, U, V;
    X.m[0] = v0.x; X.m[1] = v0.y; X.m[2] = v0.z; X.m[3] = v1.x; X.m[4] = v1.y; X.m[5] = v1.z; X.m[6] = v2.x; X.m[7] = v2.y; X.m[8] = v2.z;
    float s[3];
    svd3_rv(X, U, s, V);
    const V3 u[3] = {mk(U.m[0], U.m[1], U.m[2]), mk(U.m[3], U.m[4], U.m[5]), mk(U.m[6], U.m[7], U.m[8])};
    float A[3][3] = {{0.f, s[2], s[1]}, {s[2], 0.f, s[0]}, {s[1], s[0], 0.f}};
    float ev[3], evec[3][3];
    jacobi_eig3(A, ev, evec);
    const float c = pressure / 6.0f, inv_sqrt2 = 0.7071067811865475f;
    float weights[5];
    V3 Q[5][3];
    int kept = 0;
    const int swap_ij[3][2] = {{1, 0}, {2, 0}, {2, 1}};
    const float swap_s[3] = {s[2], s[1], s[0]};
    for (int p = 0; p < 3; ++p) {
        const float w = c * swap_s[p];
        if (!(w > 0.0f)) continue;
        const int i = swap_ij[p][0], j = swap_ij[p][1];
        weights[kept] = w;
        for (int k = 0; k < 3; ++k) Q[kept][k] = inv_sqrt2 * (V.m[j * 3 + k] * u[i] + V.m[i * 3 + k] * u[j]); // V(k, j) u_i + V(k, i) u_j
        ++kept;
    }
    for (int d = 0; d < 3 && kept < 5; ++d) {
        if (!(ev[d] < -1e-12f)) continue;
        const float w = c * fabsf(ev[d]);
        if (!(w > 0.0f)) continue;
        weights[kept] = w;
        for (int k = 0; k < 3; ++k) Q[kept][k] = (evec[d][0] * V.m[0 * 3 + k]) * u[0] + (evec[d][1] * V.m[1 * 3 + k]) * u[1] + (evec[d][2] * V.m[2 * 3 + k]) * u[2];
        ++kept;
    }
    for (int i = 0; i < 81; ++i) H[i] = 0.0f;
    for (int m = 0; m < kept; ++m)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int cc = 0; cc < 3; ++cc)
                    for (int r = 0; r < 3; ++r) H[(3 * b + cc) * 9 + 3 * a + r] += weights[m] * comp(Q[m][a], r) * comp(Q[m][b], cc);
}

} // namespace ppf

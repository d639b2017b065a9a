// math_strain.cuh -- strain limiting of shell faces: a barrier on the singular values of the deformation gradient.
//
// Restates, for one face, strainlimiting::embed_strainlimiting_force_hessian and ::line_search
// (crates/ppf-cts-solver/src/cpp/strainlimiting/strainlimiting.cu:18-120, 250-307) with their helpers:
// the fixed-sweep one-sided Jacobi SVD of a 3x2 matrix (utility/utility.cu:60-150), the barrier table on
// the strains (barrier/barrier.cu:116-131), the eigen-mode force / Hessian of an isotropic energy in
// F space (eigenanalysis/eigenanalysis.cu:11-85, symm2x2 of linalg/eigsolve.hpp:40-94) and the chain rule
// to vertex space (utility/utility.cu:185-250).  Shared by the device kernels and the host-compiled tests.
#pragma once
#include "math_contact.cuh"
#include "vecmath.cuh"

namespace ppf {

struct Svd32 {
    float U[3][2]; // orthonormal columns
    float S[2];    // singular values, ASCENDING
    float Vt[2][2];
};

// utility.cu:60-130 with N = 2, SWEEPS = 1 (one rotation orthogonalises two columns exactly)
PPF_HD void jacobi_svd32(const float F[3][2], Svd32 &o) {
    float A[3][2], Vv[2][2] = {{1.0f, 0.0f}, {0.0f, 1.0f}};
    for (int r = 0; r < 3; ++r) { A[r][0] = F[r][0]; A[r][1] = F[r][1]; }
    float a = 0.0f, b = 0.0f, g = 0.0f;
    for (int r = 0; r < 3; ++r) {
        a = a + A[r][0] * A[r][0];
        b = b + A[r][1] * A[r][1];
        g = g + A[r][0] * A[r][1];
    }
    const float zeta = (b - a) / (2.0f * g);
    const float t = (g == 0.0f) ? 0.0f : (zeta >= 0.0f ? 1.0f : -1.0f) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
    const float c = 1.0f / sqrtf(1.0f + t * t), s = c * t;
    for (int r = 0; r < 3; ++r) {
        const float ap = A[r][0], aq = A[r][1];
        A[r][0] = c * ap - s * aq;
        A[r][1] = s * ap + c * aq;
    }
    for (int r = 0; r < 2; ++r) {
        const float vp = Vv[r][0], vq = Vv[r][1];
        Vv[r][0] = c * vp - s * vq;
        Vv[r][1] = s * vp + c * vq;
    }
    float s_[2], u_[3][2];
    for (int k = 0; k < 2; ++k) {
        float n2 = 0.0f;
        for (int r = 0; r < 3; ++r) n2 = n2 + A[r][k] * A[r][k];
        const float n = sqrtf(n2);
        s_[k] = n;
        const float inv = (n > 0.0f) ? 1.0f / n : 0.0f;
        for (int r = 0; r < 3; ++r) u_[r][k] = A[r][k] * inv;
    }
    int idx[2] = {0, 1};
    if (s_[idx[1]] < s_[idx[0]]) { idx[0] = 1; idx[1] = 0; }
    for (int k = 0; k < 2; ++k) {
        o.S[k] = s_[idx[k]];
        for (int r = 0; r < 3; ++r) o.U[r][k] = u_[r][idx[k]];
        for (int r = 0; r < 2; ++r) o.Vt[k][r] = Vv[r][idx[k]]; // Vt = V^T
    }
}

// F = [x1-x0, x2-x0] * inv_rest  (utility.cu:276-282); inv_rest is the ABI's column-major 2x2
PPF_HD void deformation_grad32(const V3 &x0, const V3 &x1, const V3 &x2, const float *ir, float F[3][2]) {
    const V3 d0 = x1 - x0, d1 = x2 - x0;
    const float e0[3] = {d0.x, d0.y, d0.z}, e1[3] = {d1.x, d1.y, d1.z};
    for (int r = 0; r < 3; ++r) {
        F[r][0] = e0[r] * ir[0] + e1[r] * ir[1];
        F[r][1] = e0[r] * ir[2] + e1[r] * ir[3];
    }
}

// largest strain (singular value - 1) of F0 + t dF  (utility.cu:132-140 + maxCoeff)
PPF_HD float max_strain_at(const float F0[3][2], const float dF[3][2], float t) {
    float F[3][2];
    for (int r = 0; r < 3; ++r) { F[r][0] = F0[r][0] + t * dF[r][0]; F[r][1] = F0[r][1] + t * dF[r][1]; }
    Svd32 svd;
    jacobi_svd32(F, svd);
    const float s0 = svd.S[0] - 1.0f, s1 = svd.S[1] - 1.0f;
    return s0 < s1 ? s1 : s0;
}

// strainlimiting::line_search body for one face (strainlimiting.cu:262-301): largest t in [0, max_t] that keeps
// the largest strain below `limit`, found by bisection to float resolution
PPF_HD float strain_face_toi(const V3 *xa, const V3 *xb, const float *ir, float limit, float max_t) {
    float F0[3][2], F1[3][2], dF[3][2];
    deformation_grad32(xa[0], xa[1], xa[2], ir, F0);
    deformation_grad32(xb[0], xb[1], xb[2], ir, F1);
    for (int r = 0; r < 3; ++r) { dF[r][0] = F1[r][0] - F0[r][0]; dF[r][1] = F1[r][1] - F0[r][1]; }
    float t = max_t;
    if (max_strain_at(F0, dF, t) >= limit) {
        float upper_t = t, lower_t = 0.0f, window = upper_t - lower_t;
        while (true) {
            t = 0.5f * (upper_t + lower_t);
            const float diff = max_strain_at(F0, dF, t) - limit;
            if (diff < 0.0f) lower_t = t; else upper_t = t;
            const float nw = upper_t - lower_t;
            if (nw == window) break;
            window = nw;
        }
        t = lower_t;
    }
    return t;
}

// linalg/eigsolve.hpp:40-94 for a symmetric 2x2 (a00, a01, a11): ascending eigenvalues, eigenvector columns
PPF_HD void symm2x2(float a00, float a01, float a11, float val[2], float vec[2][2]) {
    float scale = fmaxf(fmaxf(fabsf(a00), fabsf(a01)), fmaxf(fabsf(a01), fabsf(a11)));
    if (scale <= 0.0f) { val[0] = val[1] = 0.0f; vec[0][0] = vec[1][1] = 1.0f; vec[0][1] = vec[1][0] = 0.0f; return; }
    const float b00 = a00 / scale, b01 = a01 / scale, b11 = a11 / scale;
    const float tmp = b00 - b11;
    const float D = 0.5f * sqrtf(tmp * tmp + 4.0f * b01 * b01);
    const float mid = 0.5f * (b00 + b11);
    const float l0 = mid - D, l1 = mid + D;
    // eigvectors2: u = rot90 of the first sufficiently long column of (B - l0 I), v = rot90(u)
    const float c0[2] = {b00 - l0, b01}, c1[2] = {b01, b11 - l0};
    const float sqr_eps = 1e-7f * 1e-7f;
    float u[2];
    const float r0[2] = {c0[1], -c0[0]}, r1[2] = {c1[1], -c1[0]};
    const float n0 = r0[0] * r0[0] + r0[1] * r0[1], n1 = r1[0] * r1[0] + r1[1] * r1[1];
    if (n0 > sqr_eps) { const float n = sqrtf(n0); u[0] = r0[0] / n; u[1] = r0[1] / n; }
    else if (n1 > sqr_eps) { const float n = sqrtf(n1); u[0] = r1[0] / n; u[1] = r1[1] / n; }
    else { u[0] = 1.0f; u[1] = -0.0f; } // rot90((0, 1))
    val[0] = scale * l0; val[1] = scale * l1;
    vec[0][0] = u[0]; vec[1][0] = u[1];      // column 0 = u
    vec[0][1] = u[1]; vec[1][1] = -u[0];     // column 1 = rot90(u)
}

struct StrainFace {
    V3 f[3];    // gradient per vertex (already scaled by the stiffness)
    M3 H[3][3]; // Hessian blocks (already scaled)
};

// One face of embed_strainlimiting_force_hessian (strainlimiting.cu:59-103).
//   xe: evaluation positions of the face, xc: start-of-step positions (data.vertex.curr),
//   ir: inverse rest 2x2, strainlimit / limit: the parameter and its shrink-adjusted value,
//   Hel[a][b]: block (a, b) of the elastic-only fixed matrix for the face's vertices.
// Returns false when the face is not stretched (largest strain <= 0): nothing to add.
PPF_HD bool strain_face(const V3 *xe, const V3 *xc, const float *ir, float strainlimit, float limit, float mass, const M3 Hel[3][3],
                        float eig_eps, StrainFace &out, int barrier = 0) {
    float F[3][2];
    deformation_grad32(xe[0], xe[1], xe[2], ir, F);
    Svd32 svd;
    jacobi_svd32(F, svd);
    float S[2] = {svd.S[0] - 1.0f, svd.S[1] - 1.0f}; // svd3x2_shifted
    const float smax = S[0] < S[1] ? S[1] : S[0];
    if (!(smax > 0.0f)) return false;
    // barrier table on the strains (barrier.cu:116-131), barrier gap = the strain-limit parameter
    float deda[2] = {0.0f, 0.0f}, d2[2] = {0.0f, 0.0f};
    for (int i = 0; i < 2; ++i) {
        const float g = S[i];
        if (g > 0.0f) {
            const float y = strainlimit - g;
            deda[i] = -barrier_gradient(y, strainlimit, 0.0f, barrier);
            d2[i] = barrier_curvature(y, strainlimit, 0.0f, barrier);
        }
    }
    // stiffness (strainlimiting.cu:22-44): w = start-of-step positions relative to their centroid
    const float third = 1.0f / 3.0f;
    const V3 center = third * xc[0] + third * xc[1] + third * xc[2];
    V3 w[3];
    for (int a = 0; a < 3; ++a) w[a] = xc[a] - center;
    float wHw = 0.0f;
    for (int a = 0; a < 3; ++a) {
        V3 row = mk(0.f, 0.f, 0.f);
        for (int b = 0; b < 3; ++b) row += mul(Hel[a][b], w[b]);
        wHw += dot(w[a], row);
    }
    const float d = limit - smax;
    const float stiffness = wHw + mass / (d * d);
    // eigenanalysis::compute_force: U diag(deda) V^T with the UNSHIFTED singular values restored
    S[0] += 1.0f; S[1] += 1.0f;
    float dedF[3][2];
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 2; ++j) dedF[r][j] = (svd.U[r][0] * deda[0]) * svd.Vt[0][j] + (svd.U[r][1] * deda[1]) * svd.Vt[1][j];
    // eigenanalysis::compute_hessian (2-singular-value overload)
    float ev[2], evec[2][2];
    symm2x2(d2[0], 0.0f, d2[1], ev, evec);
    const float inv_sqrt2 = 0.7071067811865475f;
    float U3[3][3];
    for (int r = 0; r < 3; ++r) { U3[r][0] = svd.U[r][0]; U3[r][1] = svd.U[r][1]; }
    U3[0][2] = svd.U[1][0] * svd.U[2][1] - svd.U[2][0] * svd.U[1][1];
    U3[1][2] = svd.U[2][0] * svd.U[0][1] - svd.U[0][0] * svd.U[2][1];
    U3[2][2] = svd.U[0][0] * svd.U[1][1] - svd.U[1][0] * svd.U[0][1];
    float Q[6][3][2];
    for (int i = 0; i < 6; ++i) for (int r = 0; r < 3; ++r) Q[i][r][0] = Q[i][r][1] = 0.0f;
    Q[0][0][1] = inv_sqrt2; Q[0][1][0] = -inv_sqrt2;
    Q[1][0][1] = inv_sqrt2; Q[1][1][0] = inv_sqrt2;
    Q[2][2][0] = 1.0f;
    Q[3][2][1] = 1.0f;
    for (int i = 0; i < 2; ++i) { Q[4][i][i] = evec[i][0]; Q[5][i][i] = evec[i][1]; }
    const float a0 = S[0], a1 = S[1], denom = a0 - a1;
    float lambda[6];
    lambda[0] = fmaxf(0.0f, (deda[0] + deda[1]) / (a0 + a1));
    lambda[1] = fmaxf(0.0f, fabsf(denom) > eig_eps ? (deda[0] - deda[1]) / denom : 0.5f * (d2[0] + d2[1]) - 0.5f * (0.0f + 0.0f));
    lambda[2] = fmaxf(0.0f, deda[0] / a0);
    lambda[3] = fmaxf(0.0f, deda[1] / a1);
    lambda[4] = fmaxf(0.0f, ev[0]);
    lambda[5] = fmaxf(0.0f, ev[1]);
    float d2F[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) d2F[i][j] = 0.0f;
    for (int i = 0; i < 6; ++i) {
        if (!lambda[i]) continue;
        float T[3][2], M[3][2];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 2; ++j) T[r][j] = U3[r][0] * Q[i][0][j] + U3[r][1] * Q[i][1][j] + U3[r][2] * Q[i][2][j];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 2; ++j) M[r][j] = T[r][0] * svd.Vt[0][j] + T[r][1] * svd.Vt[1][j];
        float q[6];
        for (int j = 0; j < 2; ++j) for (int r = 0; r < 3; ++r) q[3 * j + r] = M[r][j]; // column-major vec
        for (int x = 0; x < 6; ++x) for (int y = 0; y < 6; ++y) d2F[x][y] += lambda[i] * q[x] * q[y];
    }
    // chain rule to the vertices (utility.cu:185-198, 216-245): g_a = shape gradients
    const float irr[2][2] = {{ir[0], ir[2]}, {ir[1], ir[3]}}; // (row, col)
    const float g[3][2] = {{-irr[0][0] - irr[1][0], -irr[0][1] - irr[1][1]}, {irr[0][0], irr[0][1]}, {irr[1][0], irr[1][1]}};
    for (int a = 0; a < 3; ++a) {
        float v[3];
        for (int dim = 0; dim < 3; ++dim) v[dim] = g[a][0] * dedF[dim][0] + g[a][1] * dedF[dim][1];
        out.f[a] = stiffness * mk(v[0], v[1], v[2]);
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            M3 blk = m3_zero();
            for (int dd = 0; dd < 2; ++dd)
                for (int ee = 0; ee < 2; ++ee) {
                    const float sc = g[a][dd] * g[b][ee];
                    for (int c = 0; c < 3; ++c)
                        for (int r = 0; r < 3; ++r) blk.m[c * 3 + r] += sc * d2F[3 * dd + r][3 * ee + c];
                }
            for (int e = 0; e < 9; ++e) out.H[a][b].m[e] = stiffness * blk.m[e];
        }
    return true;
}

// ---- isotropic shell models (energy.cu:337-352): ARAP / StVK / SNHk as functions of the two singular values ----
// table of first / second derivatives w.r.t. the singular values: energy/model/arap.hpp:36-49, stvk.hpp:38-72 with
// detsqr.hpp:22-33, snhk.hpp:45-56.  model: the ABI's Model enum (0 ARAP, 1 StVK, 3 SNHk).  d2 = (d00, d01, d11).
PPF_HD void iso_table2(int model, const float a[2], float mu, float lmd, float deda[2], float d2[3]) {
    if (model == 0) { // ARAP
        const float trace_strain = (a[0] - 1.0f) + (a[1] - 1.0f);
        deda[0] = 2.0f * mu * (a[0] - 1.0f) + lmd * trace_strain * 1.0f;
        deda[1] = 2.0f * mu * (a[1] - 1.0f) + lmd * trace_strain * 1.0f;
        d2[0] = 2.0f * mu * 1.0f + lmd;
        d2[2] = 2.0f * mu * 1.0f + lmd;
        d2[1] = 2.0f * mu * 0.0f + lmd;
    } else if (model == 1) { // StVK: mu/4 sum (a_i - 1)^4 + lmd/2 (a0 a1 - 1)^2
        const float a0 = a[0] - 1.0f, a1 = a[1] - 1.0f;
        const float J = a[0] * a[1];
        deda[0] = mu * (a0 * a0 * a0) + lmd * a[1] * (J - 1.0f);
        deda[1] = mu * (a1 * a1 * a1) + lmd * a[0] * (J - 1.0f);
        d2[0] = mu * (3.0f * a0 * a0) + lmd * (a[1] * a[1]);
        d2[2] = mu * (3.0f * a1 * a1) + lmd * (a[0] * a[0]);
        d2[1] = mu * 0.0f + lmd * (2.0f * a[0] * a[1] - 1.0f);
    } else { // SNHk
        const float x = a[0], y = a[1];
        deda[0] = lmd * y * (x * y - 1.0f) + mu * (x - y);
        deda[1] = lmd * x * (x * y - 1.0f) + mu * (y - x);
        d2[0] = lmd * y * y + mu;
        d2[2] = lmd * x * x + mu;
        d2[1] = 2.0f * lmd * x * y - lmd - mu;
    }
}
// One face of embed_face_force_hessian for the SVD-based models: force U diag(deda) V^T and the eigen-projected Hessian
// (eigenanalysis.cu:14-85) chained to the vertices and scaled by the face mass.
PPF_HD void iso_face(const V3 &x0, const V3 &x1, const V3 &x2, const float *ir, int model, float mu, float lmd, float mass, float eig_eps,
                     StrainFace &out) {
    float F[3][2];
    deformation_grad32(x0, x1, x2, ir, F);
    Svd32 svd;
    jacobi_svd32(F, svd);
    float deda[2], d2[3];
    iso_table2(model, svd.S, mu, lmd, deda, d2);
    float dedF[3][2];
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 2; ++j) dedF[r][j] = (svd.U[r][0] * deda[0]) * svd.Vt[0][j] + (svd.U[r][1] * deda[1]) * svd.Vt[1][j];
    float ev[2], evec[2][2];
    symm2x2(d2[0], d2[1], d2[2], ev, evec);
    const float inv_sqrt2 = 0.7071067811865475f;
    float U3[3][3];
    for (int r = 0; r < 3; ++r) { U3[r][0] = svd.U[r][0]; U3[r][1] = svd.U[r][1]; }
    U3[0][2] = svd.U[1][0] * svd.U[2][1] - svd.U[2][0] * svd.U[1][1];
    U3[1][2] = svd.U[2][0] * svd.U[0][1] - svd.U[0][0] * svd.U[2][1];
    U3[2][2] = svd.U[0][0] * svd.U[1][1] - svd.U[1][0] * svd.U[0][1];
    float Q[6][3][2];
    for (int i = 0; i < 6; ++i) for (int r = 0; r < 3; ++r) Q[i][r][0] = Q[i][r][1] = 0.0f;
    Q[0][0][1] = inv_sqrt2; Q[0][1][0] = -inv_sqrt2;
    Q[1][0][1] = inv_sqrt2; Q[1][1][0] = inv_sqrt2;
    Q[2][2][0] = 1.0f;
    Q[3][2][1] = 1.0f;
    for (int i = 0; i < 2; ++i) { Q[4][i][i] = evec[i][0]; Q[5][i][i] = evec[i][1]; }
    const float a0 = svd.S[0], a1 = svd.S[1], denom = a0 - a1;
    float lambda[6];
    lambda[0] = fmaxf(0.0f, (deda[0] + deda[1]) / (a0 + a1));
    lambda[1] = fmaxf(0.0f, fabsf(denom) > eig_eps ? (deda[0] - deda[1]) / denom : 0.5f * (d2[0] + d2[2]) - 0.5f * (d2[1] + d2[1]));
    lambda[2] = fmaxf(0.0f, deda[0] / a0);
    lambda[3] = fmaxf(0.0f, deda[1] / a1);
    lambda[4] = fmaxf(0.0f, ev[0]);
    lambda[5] = fmaxf(0.0f, ev[1]);
    float d2F[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) d2F[i][j] = 0.0f;
    for (int i = 0; i < 6; ++i) {
        if (!lambda[i]) continue;
        float T[3][2], M[3][2];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 2; ++j) T[r][j] = U3[r][0] * Q[i][0][j] + U3[r][1] * Q[i][1][j] + U3[r][2] * Q[i][2][j];
        for (int r = 0; r < 3; ++r)
            for (int j = 0; j < 2; ++j) M[r][j] = T[r][0] * svd.Vt[0][j] + T[r][1] * svd.Vt[1][j];
        float q[6];
        for (int j = 0; j < 2; ++j) for (int r = 0; r < 3; ++r) q[3 * j + r] = M[r][j];
        for (int x = 0; x < 6; ++x) for (int y = 0; y < 6; ++y) d2F[x][y] += lambda[i] * q[x] * q[y];
    }
    const float irr[2][2] = {{ir[0], ir[2]}, {ir[1], ir[3]}};
    const float g[3][2] = {{-irr[0][0] - irr[1][0], -irr[0][1] - irr[1][1]}, {irr[0][0], irr[0][1]}, {irr[1][0], irr[1][1]}};
    for (int a = 0; a < 3; ++a) {
        float v[3];
        for (int dim = 0; dim < 3; ++dim) v[dim] = g[a][0] * dedF[dim][0] + g[a][1] * dedF[dim][1];
        out.f[a] = mass * mk(v[0], v[1], v[2]);
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            M3 blk = m3_zero();
            for (int dd = 0; dd < 2; ++dd)
                for (int ee = 0; ee < 2; ++ee) {
                    const float sc = g[a][dd] * g[b][ee];
                    for (int c = 0; c < 3; ++c)
                        for (int r = 0; r < 3; ++r) blk.m[c * 3 + r] += sc * d2F[3 * dd + r][3 * ee + c];
                }
            for (int e = 0; e < 9; ++e) out.H[a][b].m[e] = mass * blk.m[e];
        }
}

} // namespace ppf

// math_plasticity.cuh -- plastic creep of the rest shape after a step (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/plasticity/plasticity.cu: update_face_plasticity 21-86 (singular values of the
// deformation gradient outside the dead zone [1 - threshold, 1 + threshold] creep toward its nearer edge and the inverse
// rest matrix is rebuilt from the relaxed deformation), update_hinge_plasticity 147-188 and update_rod_bend_plasticity
// 190-245 (the rest angle creeps toward the current angle outside its dead zone).  alpha = 1 - exp(-rate * dt) with the
// session's dt (param.dt, not the step's possibly shortened one); __expf on the device like the reference's fmath::exp.
#pragma once
#include "math_hinge.cuh"
#include "math_strain.cuh"

namespace ppf {

PPF_HD float plastic_alpha(float rate, float dt) {
#if defined(__CUDA_ARCH__)
    return 1.0f - __expf(-rate * dt);
#else
    return 1.0f - expf(-rate * dt);
#endif
}
// increment of a rest angle whose deviation delta exceeds the threshold (plasticity.cu:12-19)
PPF_HD float dead_zone_creep(float delta, float threshold, float alpha) {
    const float target = (delta > 0.0f) ? threshold : -threshold;
    return alpha * (delta - target);
}
// face: ir = inverse rest 2x2 (column-major, ABI layout), updated in place; returns true when it changed
PPF_HD bool face_plasticity(const V3 &x0, const V3 &x1, const V3 &x2, float *ir, float threshold, float alpha) {
    float F[3][2];
    deformation_grad32(x0, x1, x2, ir, F);
    Svd32 svd;
    jacobi_svd32(F, svd);
    bool changed = false;
    float S_new[2] = {svd.S[0], svd.S[1]};
    for (int k = 0; k < 2; ++k) {
        const float dev = fabsf(svd.S[k] - 1.0f);
        if (dev > threshold) {
            const float target = (svd.S[k] < 1.0f) ? (1.0f - threshold) : (1.0f + threshold);
            S_new[k] = svd.S[k] + alpha * (target - svd.S[k]);
            changed = true;
        }
    }
    if (!changed) return false;
    float Fn[3][2]; // U diag(S_new) V^T
    for (int r = 0; r < 3; ++r)
        for (int j = 0; j < 2; ++j) Fn[r][j] = (svd.U[r][0] * S_new[0]) * svd.Vt[0][j] + (svd.U[r][1] * S_new[1]) * svd.Vt[1][j];
    const V3 d0 = x1 - x0, d1 = x2 - x0;
    const V3 n = normalized(cross(d0, d1));
    const V3 t = normalized(d0);
    const V3 b = normalized(cross(n, t));
    // P1 = [t; b] (2x3);  P1_dx = P1 [d0 d1];  inv_rest = P1_dx^{-1} (P1 F_new)
    const float m00 = dot(t, d0), m01 = dot(t, d1), m10 = dot(b, d0), m11 = dot(b, d1);
    const float det = m00 * m11 - m01 * m10;
    const float i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
    float PF[2][2];
    for (int j = 0; j < 2; ++j) {
        const V3 col = mk(Fn[0][j], Fn[1][j], Fn[2][j]);
        PF[0][j] = dot(t, col);
        PF[1][j] = dot(b, col);
    }
    const float r00 = i00 * PF[0][0] + i01 * PF[1][0], r01 = i00 * PF[0][1] + i01 * PF[1][1];
    const float r10 = i10 * PF[0][0] + i11 * PF[1][0], r11 = i10 * PF[0][1] + i11 * PF[1][1];
    ir[0] = r00; ir[1] = r10; ir[2] = r01; ir[3] = r11; // column-major
    return true;
}
// hinge: X = remapped vertices (h2, h1, h0, h3) at the new positions; returns the new rest angle
PPF_HD float hinge_plasticity(const V3 *X, float rest_angle, float threshold, float alpha) {
    const float theta = hinge_angle(X[0], X[1], X[2], X[3]);
    const float delta = theta - rest_angle;
    return fabsf(delta) > threshold ? rest_angle + dead_zone_creep(delta, threshold, alpha) : rest_angle;
}
// interior rod vertex i with neighbours j, k; returns the new rest bend angle
PPF_HD float rod_bend_plasticity(const V3 &xj, const V3 &xi, const V3 &xk, float rest_angle, float threshold, float alpha) {
    const V3 e0 = xj - xi, e1 = xk - xi;
    const float n0 = norm(e0), n1 = norm(e1);
    if (n0 <= 0.0f || n1 <= 0.0f) return rest_angle;
    const float cos_theta = fmaxf(-1.0f, fminf(1.0f, dot(e0, e1) / (n0 * n1)));
    const float theta = acosf(cos_theta);
    const float delta = theta - rest_angle;
    return fabsf(delta) > threshold ? rest_angle + dead_zone_creep(delta, threshold, alpha) : rest_angle;
}

} // namespace ppf

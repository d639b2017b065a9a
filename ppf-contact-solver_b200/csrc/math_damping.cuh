// math_damping.cuh -- stiffness-proportional (Rayleigh) damping of one element (fp32).
//
// Restates energy::add_stiffness_damping<N> and add_stiffness_damping_lagged<N>
// (crates/ppf-cts-solver/src/cpp/energy/energy.cu:88-180): with s = beta / dt and u = x - x^n (x^n = start-of-step
// positions), the membrane / rod form adds s K u to the element gradient and scales the element Hessian K by (1 + s);
// the lagged form used for bending adds s K_lag u and s K_lag, K_lag being the element Hessian at x^n.
// K is a dense (3N) x (3N) column-major matrix here; the callers keep their own block layouts and convert.
#pragma once
#include "vecmath.cuh"

namespace ppf {

template <int N> PPF_HD void damping_gradient(const float *K /* (3N)^2 column-major */, const V3 *u, float s, V3 *f) {
    for (int a = 0; a < N; ++a) {
        float f0 = 0.0f, f1 = 0.0f, f2 = 0.0f;
        for (int b = 0; b < 3 * N; ++b) {
            const float ub = comp(u[b / 3], b % 3);
            f0 += K[b * 3 * N + 3 * a + 0] * ub;
            f1 += K[b * 3 * N + 3 * a + 1] * ub;
            f2 += K[b * 3 * N + 3 * a + 2] * ub;
        }
        f[a].x += s * f0;
        f[a].y += s * f1;
        f[a].z += s * f2;
    }
}
// f += s K u;  K *= (1 + s)
template <int N> PPF_HD void stiffness_damping(float *K, const V3 *u, float beta, float dt, V3 *f) {
    if (beta <= 0.0f || dt <= 0.0f) return;
    const float s = beta / dt;
    damping_gradient<N>(K, u, s, f);
    for (int i = 0; i < 9 * N * N; ++i) K[i] *= (1.0f + s);
}
// f += s K_lag u;  K += s K_lag
template <int N> PPF_HD void stiffness_damping_lagged(float *K, const float *K_lag, const V3 *u, float beta, float dt, V3 *f) {
    if (beta <= 0.0f || dt <= 0.0f) return;
    const float s = beta / dt;
    damping_gradient<N>(K_lag, u, s, f);
    for (int i = 0; i < 9 * N * N; ++i) K[i] += s * K_lag[i];
}

} // namespace ppf

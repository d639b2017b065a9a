// math_contact.cuh -- cubic barrier, elasticity-inclusive contact stiffness, lagged Coulomb friction.
//
// Restates (fp32, same evaluation order):
//   cubic barrier          crates/ppf-cts-solver/src/cpp/barrier/cubic.hpp:13-41
//   edge gradient/hessian  barrier/barrier.cu:103-114
//   compute_stiffness<N>   barrier/barrier.cu:50-86   kappa = w^T (H_elastic,local + diag(m/g^2)) w
//   Friction               energy/model/friction.hpp:19-82
// Only the cubic barrier is implemented (Barrier::Cubic is the default and the only one the
// BASELINE configs use); the backend refuses quad/log loudly at initialize.
#pragma once
#include "vecmath.cuh"

namespace ppf {

PPF_HD float cubic_gradient(float g, float ghat, float offset) {
    g -= offset;
    const float y = g - ghat;
    return y < 0.0f ? -2.0f * y * y / ghat : 0.0f;
}
PPF_HD float cubic_curvature(float g, float ghat, float offset) {
    g -= offset;
    return (g - ghat < 0.0f) ? 4.0f * (1.0f - g / ghat) : 0.0f;
}
// the other two barriers of ParamSet.barrier (barrier/quadratic.hpp:13-41, barrier/logarithm.hpp:14-44); kind = the ABI's
// Barrier enum (0 cubic, 1 quad, 2 log).  The log barrier uses the fast intrinsic on the device like the reference's fmath::log.
PPF_HD float fast_log(float x) {
#if defined(__CUDA_ARCH__)
    return __logf(x);
#else
    return logf(x);
#endif
}
PPF_HD float barrier_gradient(float g, float ghat, float offset, int kind) {
    if (kind == 0) return cubic_gradient(g, ghat, offset);
    g -= offset;
    if (kind == 1) {
        const float y = ghat - g;
        return y > 0.0f ? -2.0f * y : 0.0f;
    }
    if (g <= 0.0f) return -INFINITY;
    if (g >= ghat) return 0.0f;
    return (ghat - g) * (2.0f * g * fast_log(g / ghat) + g - ghat) / g;
}
PPF_HD float barrier_curvature(float g, float ghat, float offset, int kind) {
    if (kind == 0) return cubic_curvature(g, ghat, offset);
    g -= offset;
    if (kind == 1) return (ghat - g > 0.0f) ? 2.0f : 0.0f;
    if (g <= 0.0f) return INFINITY;
    if (g >= ghat) return 0.0f;
    return -2.0f * fast_log(g / ghat) + ghat * (ghat + 2.0f * g) / (g * g) - 3.0f;
}
PPF_HD V3 barrier_edge_gradient(const V3 &e, float ghat, float offset, int kind = 0) {
    const float g = norm(e);
    const V3 n = e / g;
    return n * barrier_gradient(g, ghat, offset, kind);
}
PPF_HD M3 barrier_edge_hessian(const V3 &e, float ghat, float offset, int kind = 0) {
    const float c = barrier_curvature(norm(e), ghat, offset, kind);
    // c * e * e^T / |e|^2  evaluated left to right: ((c*e) e^T) / sq
    const V3 ce = c * e;
    const float sq = sqnorm(e);
    M3 o = m3_outer(ce, e);
    for (int i = 0; i < 9; ++i) o.m[i] = o.m[i] / sq;
    return o;
}

struct FrictionOut {
    V3 grad;
    M3 hess;
    float lambda;
};
// friction.hpp:19-82
PPF_HD FrictionOut friction_eval(const V3 &force_contact, const V3 &dx, const V3 &normal, float mu, float min_dx) {
    FrictionOut o;
    const float contact = -dot(normal, force_contact);
    const V3 u = dx - normal * dot(normal, dx);
    const float u_norm = sqrtf(sqnorm(u));
    const bool kinetic = u_norm > min_dx;
    const V3 slip_dir = kinetic ? u / u_norm : mk(0.0f, 0.0f, 0.0f);
    o.lambda = mu > 0.0f ? mu * contact / fmaxf(min_dx, u_norm) : 0.0f;
    o.grad = o.lambda * u;
    if (kinetic) {
        const V3 w = cross(normal, slip_dir);
        o.hess = o.lambda * m3_outer(w, w);
    } else {
        const M3 P = m3_identity() - m3_outer(normal, normal);
        o.hess = o.lambda * P;
    }
    return o;
}

PPF_HD float combine_friction(float a, float b, int mode) { // contact.cu:64-74 (Min=0, Max=1, Mean=2)
    return mode == 1 ? fmaxf(a, b) : (mode == 2 ? 0.5f * (a + b) : fminf(a, b));
}

} // namespace ppf

// math_cmesh.cuh -- contact of the simulated mesh with the STATIC collision mesh (Constraint.mesh): penalty ("push") force,
// Hessian and lagged friction of one dynamic vertex / edge / face against one static face / edge / vertex (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/contact/contact.cu:899-962 (dynamic vertex vs static face), 964-1061 (dynamic face
// vs static vertex) and 1063-1151 (dynamic edge vs static edge).  The static side never moves and owns no unknowns, so a
// contact adds c_i f to the force and c_i c_j H to the blocks of the dynamic element only; the stiffness is the same
// elasticity-inclusive one as everywhere else: the dynamic element's elastic blocks plus m / gap^2 on the diagonal, taken
// along the contact normal spread over the element by the barycentric weights.
#pragma once
#include "math_collider.cuh"
#include "math_distance.cuh"

namespace ppf {

struct CmeshOut {
    V3 f;       // force on the contact point: vertex k of the dynamic element receives c[k] * f
    M3 H;       // block (j, k) of the element receives c[j] * c[k] * H
    float c[3]; // weights of the dynamic element's vertices (1 for a vertex)
};

// n^T A n / |n|^2 of the 3N x 3N matrix made of the blocks B[j][k] with m_j / gap^2 on the diagonal, n = (c_j * normal)_j
template <int N> PPF_HD float cmesh_stiffness(const M3 (*B)[N], const float *vmass, float gap_squared, const float *c, const V3 &normal) {
    float v[3 * N], Av[3 * N];
    for (int j = 0; j < 3 * N; ++j) v[j] = c[j / 3] * comp(normal, j % 3);
    for (int r = 0; r < 3 * N; ++r) {
        float acc = 0.0f;
        for (int k = 0; k < 3 * N; ++k) {
            float a = B[r / 3][k / 3].m[(k % 3) * 3 + (r % 3)];
            if (r == k) a += vmass[r / 3] / gap_squared;
            acc += a * v[k];
        }
        Av[r] = acc;
    }
    float num = 0.0f, den = 0.0f;
    for (int r = 0; r < 3 * N; ++r) { num += Av[r] * v[r]; den += v[r] * v[r]; }
    return num / den;
}

// push force + friction of a contact point at e (from the static closest point), a = e - (offset + ghat) normal
PPF_HD void cmesh_force(const V3 &a, const V3 &normal, const V3 &dx, float ghat, float mu, float stiff_k, float friction_eps, V3 &f, M3 &H) {
    f = stiff_k * push_gradient(a, normal, ghat);
    H = stiff_k * push_hessian(a, normal, ghat);
    const FrictionOut fr = friction_eval(f, dx, normal, mu, friction_eps);
    f += fr.grad;
    H = H + fr.hess;
}

// dynamic vertex p (start-of-step position p_start) against the static triangle (T0, T1, T2); Hii = the vertex's elastic block
PPF_HD bool cmesh_vertex_face(const V3 &p, const V3 &p_start, const V3 &T0, const V3 &T1, const V3 &T2, float offset, float ghat, float mu,
                              float mass, const M3 &Hii, float friction_eps, CmeshOut &o) {
    const V3 t0 = T0 - p, t1 = T1 - p, t2 = T2 - p;
    const C3 c = point_triangle_coeff_unclassified(mk(0.f, 0.f, 0.f), t0, t1, t2);
    const V3 y = c.a * t0 + c.b * t1 + c.c * t2;
    const V3 e = -y;
    const V3 q = p_start - p;
    const float s = offset + ghat;
    if (!(sqnorm(e) < s * s)) return false;
    const V3 normal = normalized(e);
    const V3 proj_x = y + normal * s;
    const float gap = norm(e) - offset;
    const float stiff_k = dot(normal, mul(Hii, normal)) + mass / (gap * gap);
    cmesh_force(-proj_x, normal, -q, ghat, mu, stiff_k, friction_eps, o.f, o.H);
    o.c[0] = 1.0f; o.c[1] = o.c[2] = 0.0f;
    return true;
}

// dynamic face X[0..2] (start-of-step Xs) against the static vertex y; B = the face's 3x3 elastic blocks, vmass its vertex masses
PPF_HD bool cmesh_face_vertex(const V3 &y, const V3 *X, const V3 *Xs, float offset, float ghat, float mu, const float *vmass,
                              const M3 (*B)[3], float friction_eps, CmeshOut &o) {
    const V3 t0 = X[0] - y, t1 = X[1] - y, t2 = X[2] - y;
    const V3 s0 = Xs[0] - y, s1 = Xs[1] - y, s2 = Xs[2] - y;
    const C3 c = point_triangle_coeff_unclassified(mk(0.f, 0.f, 0.f), t0, t1, t2);
    const V3 p = c.a * t0 + c.b * t1 + c.c * t2;
    const V3 q = c.a * s0 + c.b * s1 + c.c * s2;
    const float s = offset + ghat;
    if (!(sqnorm(p) < s * s)) return false;
    const V3 normal = normalized(p);
    const float gap = norm(p) - offset;
    o.c[0] = c.a; o.c[1] = c.b; o.c[2] = c.c;
    const float stiff_k = cmesh_stiffness<3>(B, vmass, gap * gap, o.c, normal);
    const V3 proj_x = normal * s;
    cmesh_force(p - proj_x, normal, p - q, ghat, mu, stiff_k, friction_eps, o.f, o.H);
    return true;
}

// dynamic edge (P0, P1) (start-of-step Ps) against the static edge (Q0, Q1); B = the edge's 2x2 elastic blocks
PPF_HD bool cmesh_edge_edge(const V3 *P, const V3 *Ps, const V3 &Q0, const V3 &Q1, float offset, float ghat, float mu, const float *vmass,
                            const M3 (*B)[2], float friction_eps, CmeshOut &o) {
    const C4 c = edge_edge_coeff(P[0], P[1], Q0, Q1);
    const V3 x = c.a * P[0] + c.b * P[1];
    const V3 y = c.c * Q0 + c.d * Q1;
    const V3 z = c.a * Ps[0] + c.b * Ps[1];
    const V3 e = x - y;
    const float s = offset + ghat;
    if (!(sqnorm(e) < s * s)) return false;
    const V3 normal = normalized(e);
    const V3 proj_x = s * normal;
    const float gap = norm(e) - offset;
    o.c[0] = c.a; o.c[1] = c.b; o.c[2] = 0.0f;
    const float stiff_k = cmesh_stiffness<2>(B, vmass, gap * gap, o.c, normal);
    cmesh_force(e - proj_x, normal, x - z, ghat, mu, stiff_k, friction_eps, o.f, o.H);
    return true;
}

} // namespace ppf

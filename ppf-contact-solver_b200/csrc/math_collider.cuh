// math_collider.cuh -- analytic colliders (invisible spheres / bowls and walls): penalty force and Hessian of one
// free vertex against every collider, and the collider part of the step-size line search (fp32).
//
// Restates crates/ppf-cts-solver/src/cpp/energy/model/push.hpp:15-50 (quadratic push barrier),
// contact/contact.cu:1154-1383 (embed_vertex_constraint_force_hessian: sphere / kinematic sphere / floor branches
// with the elasticity-inclusive stiffness n^T H_ii n + m / gap^2 and lagged Coulomb friction) and
// contact/contact.cu:2318-2470 (vertex_constraint_line_search: first crossing of a static collider by linear
// interpolation of the signed distance, and the "prescribed vertex driven through a collider" detection).
// The fix-pin barrier branch of the reference only applies to PDRD anchors / PPF_DISABLE_PIN_DOF_REMOVAL, which
// this backend refuses.
#pragma once
#include "../../include/ppf_abi.h"
#include "math_contact.cuh"

namespace ppf {

PPF_HD V3 push_gradient(const V3 &e, const V3 &normal, float eps) {
    const float d = dot(e, normal);
    return d < 0.0f ? (-(d * d) / eps) * normal : mk(0.f, 0.f, 0.f);
}
PPF_HD M3 push_hessian(const V3 &e, const V3 &normal, float eps) {
    const float d = dot(e, normal);
    const float c = d < 0.0f ? -2.0f * d / eps : 0.0f;
    return m3_outer(c * normal, normal);
}
PPF_HD V3 v3_of(const PpfVec3f &p) { return mk(p.v[0], p.v[1], p.v[2]); }

struct ColliderOut {
    V3 f;
    M3 H;
    unsigned count;
};

// x: evaluation position, dx: x - start-of-step position, Hii: the vertex's diagonal block of the elastic-only matrix
PPF_HD void collider_vertex(const V3 &x, const V3 &dx, float mass, float vfriction, const M3 &Hii, const PpfSphere *spheres, unsigned n_sphere,
                            const PpfFloor *floors, unsigned n_floor, int friction_mode, float friction_eps, float constraint_tol, ColliderOut &o) {
    o.f = mk(0.f, 0.f, 0.f);
    o.H = m3_zero();
    o.count = 0;
    for (unsigned j = 0; j < n_sphere; ++j) {
        const PpfSphere &s = spheres[j];
        const float ghat = s.ghat;
        const float friction = combine_friction(s.friction, vfriction, friction_mode);
        const bool bowl = s.bowl;
        bool reverse = s.reverse;
        float radius = s.radius;
        V3 center = v3_of(s.center);
        if (bowl && x.y > center.y) center = mk(center.x, x.y, center.z);
        const float d2 = sqnorm(x - center);
        if (s.kinematic) {
            if (d2) {
                const V3 normal = normalized(x - center);
                const float eff_radius = reverse ? (radius - ghat) : (radius + ghat);
                const V3 target = eff_radius * normal;
                const V3 ov = x - center;
                if (bowl) reverse = true;
                const float stiff_k = mass / (ghat * ghat);
                const float r2 = eff_radius * eff_radius;
                const float dist = sqrtf(d2);
                const float depth = reverse ? (dist - s.radius) : (s.radius - dist);
                if (s.thickness > 0.0f && depth > s.thickness) continue;
                if (reverse && d2 > r2) {
                    o.count += 1;
                    o.f += stiff_k * push_gradient(ov - target, -normal, ghat);
                    o.H += stiff_k * push_hessian(ov - target, -normal, ghat);
                } else if (!reverse && d2 < r2) {
                    o.count += 1;
                    o.f += stiff_k * push_gradient(ov - target, normal, ghat);
                    o.H += stiff_k * push_hessian(ov - target, normal, ghat);
                }
            }
        } else {
            if (reverse) radius -= ghat;
            else radius += ghat;
            const float r2 = radius * radius;
            const bool intersected = reverse ? d2 > r2 : d2 < r2;
            if (intersected) {
                const float dist = sqrtf(d2);
                const float depth = reverse ? (dist - s.radius) : (s.radius - dist);
                if (s.thickness > 0.0f && depth > s.thickness) continue;
                o.count += 1;
                V3 normal = normalized(x - center);
                const V3 projected_x = radius * normal;
                const V3 ov = x - center;
                if (reverse) normal = -normal;
                const float gap = reverse ? s.radius - sqrtf(d2) : sqrtf(d2) - s.radius;
                const float stiff_k = dot(normal, mul(Hii, normal)) + mass / (gap * gap);
                const V3 f_push = stiff_k * push_gradient(ov - projected_x, normal, ghat);
                o.f += f_push;
                o.H += stiff_k * push_hessian(ov - projected_x, normal, ghat);
                const FrictionOut fr = friction_eval(f_push, dx, normal, friction, friction_eps);
                o.f += fr.grad;
                o.H += fr.hess;
            }
        }
    }
    for (unsigned j = 0; j < n_floor; ++j) {
        const PpfFloor &fl = floors[j];
        const float ghat = fl.ghat;
        const float friction = combine_friction(fl.friction, vfriction, friction_mode);
        const V3 up = v3_of(fl.up);
        const V3 ground = v3_of(fl.ground) + ghat * up;
        const V3 e = x - ground;
        if (dot(e, up) < 0.0f) {
            const float depth = -dot(e, up);
            if (fl.thickness > 0.0f && depth > fl.thickness) continue;
            o.count += 1;
            const V3 projected_x = (-dot(e, up)) * up;
            float gap = dot(x - v3_of(fl.ground), up);
            if (fl.kinematic) gap = fmaxf(gap, constraint_tol * ghat);
            const float stiff_k = dot(up, mul(Hii, up)) + mass / (gap * gap);
            const V3 f_push = stiff_k * push_gradient(-projected_x, up, ghat);
            o.f += f_push;
            o.H += stiff_k * push_hessian(-projected_x, up, ghat);
            const FrictionOut fr = friction_eval(f_push, dx, up, friction, friction_eps);
            o.f += fr.grad;
            o.H += fr.hess;
        }
    }
}

// Line search against the STATIC colliders over the segment y0 -> y0 + max_t (y1 - y0).  Free vertex: returns the
// largest admissible t in (0, max_t].  Pinned vertex (pinned = true): returns max_t and sets *infeasible when the
// prescribed motion crosses a collider from outside to inside.
PPF_HD float collider_toi(const V3 &y0, const V3 &y1, bool pinned, const PpfSphere *spheres, unsigned n_sphere, const PpfFloor *floors,
                          unsigned n_floor, float max_t, bool *infeasible) {
    const V3 x1 = max_t * (y1 - y0) + y0;
    const V3 &x0 = y0;
    float toi = max_t;
    for (unsigned j = 0; j < n_sphere; ++j) {
        const PpfSphere &s = spheres[j];
        if (s.kinematic) continue;
        const V3 center = v3_of(s.center);
        const V3 center0 = (s.bowl && x0.y > center.y) ? mk(center.x, x0.y, center.z) : center;
        const V3 center1 = (s.bowl && x1.y > center.y) ? mk(center.x, x1.y, center.z) : center;
        const float r = s.radius;
        const float r0 = norm(x0 - center0);
        const float r1 = norm(x1 - center1);
        if (pinned) {
            const float d0 = s.reverse ? (r - r0) : (r0 - r);
            const float d1 = s.reverse ? (r - r1) : (r1 - r);
            if (d0 >= 0.0f && d1 < 0.0f) *infeasible = true;
            continue;
        }
        const float depth0 = s.reverse ? (r0 - r) : (r - r0);
        if (s.thickness > 0.0f && depth0 > s.thickness) continue;
        if ((r0 - r) * (r1 - r) <= 0.0f) {
            const float t = (r - r0) / (r1 - r0);
            toi = fminf(toi, max_t * t);
        }
    }
    for (unsigned j = 0; j < n_floor; ++j) {
        const PpfFloor &fl = floors[j];
        if (fl.kinematic) continue;
        const V3 up = v3_of(fl.up), ground = v3_of(fl.ground);
        const float h0 = dot(up, x0 - ground);
        const float h1 = dot(up, x1 - ground);
        if (pinned) {
            if (h0 >= 0.0f && h1 < 0.0f) *infeasible = true;
            continue;
        }
        if (fl.thickness > 0.0f && -h0 > fl.thickness) continue;
        if (h1 < 0.0f) {
            const float t = -h0 / (h1 - h0);
            toi = fminf(toi, max_t * t);
        }
    }
    return toi;
}

} // namespace ppf

// bvh.cuh -- device view of an LBVH and the stack traversal used by every broad phase.
#pragma once
#include "backend.hpp"

namespace ppf {

struct BvhView {
    unsigned n;               // leaves; internal nodes [0, n-1), leaf k is node (n-1)+k
    const int2 *child;        // internal node -> (left, right)
    const float4 *lo, *hi;    // node boxes
    const unsigned *prim;     // sorted leaf k -> primitive index
    const float4 *ulo, *uhi;  // node displacement intervals (valid after bvh_refit_relative)
};

inline BvhView view_of(const Bvh &b) { return BvhView{b.n, b.child.p, b.lo.p, b.hi.p, b.prim_sorted.p, b.ulo.p, b.uhi.p}; }

struct Box {
    float lx, ly, lz, hx, hy, hz;
};
// inclusive overlap, as the reference's aabb::overlap (contact/aabb.hpp:71-77)
__device__ __forceinline__ bool box_overlap(const Box &q, const float4 lo, const float4 hi) {
    return (q.lx <= hi.x && q.hx >= lo.x) && (q.ly <= hi.y && q.hy >= lo.y) && (q.lz <= hi.z && q.hz >= lo.z);
}

#define PPF_BVH_STACK 96
// Calls f(primitive) for every leaf whose box overlaps q.  Returns false on stack overflow.
template <class F> __device__ __forceinline__ bool bvh_query(const BvhView &b, const Box &q, F &&f) {
    if (b.n == 0) return true;
    if (b.n == 1) {
        if (box_overlap(q, b.lo[0], b.hi[0])) f(b.prim[0]);
        return true;
    }
    int stack[PPF_BVH_STACK];
    int sp = 0;
    if (!box_overlap(q, b.lo[0], b.hi[0])) return true;
    stack[sp++] = 0;
    const int first_leaf = (int)b.n - 1;
    while (sp) {
        const int node = stack[--sp];
        const int2 c = b.child[node];
        const bool hit_l = box_overlap(q, b.lo[c.x], b.hi[c.x]);
        const bool hit_r = box_overlap(q, b.lo[c.y], b.hi[c.y]);
        if (hit_l) {
            if (c.x >= first_leaf) f(b.prim[c.x - first_leaf]);
            else stack[sp++] = c.x;
        }
        if (hit_r) {
            if (c.y >= first_leaf) f(b.prim[c.y - first_leaf]);
            else {
                if (sp >= PPF_BVH_STACK) return false;
                stack[sp++] = c.y;
            }
        }
    }
    return true;
}

// ---- relative-motion sweep query --------------------------------------------------------------
// The reference bounds a swept primitive by the hull of its boxes at t = 0 and t = t_max.  Under
// coherent motion (a cloth that rotates or translates as a whole moves ~20 edge lengths per step in
// the twist scenes) those hulls overlap hundreds of neighbours that can never touch.  Here a node
// keeps its box at t = 0 and the component-wise interval U of its vertices' displacements
// t_max*(x1-x0).  A point a of the query primitive and a point b below the node can come within the
// contact gap during the sweep only if, per axis,  a0 - b0 is in  t*(ub - ua) +- gap  for some t in
// [0,1], ub in U_node, ua in U_query -- i.e. only if the query's START box overlaps the node's START
// box widened by  [min(0, Ulo_node - Uhi_query), max(0, Uhi_node - Ulo_query)].  Displacements of
// points of a primitive are convex combinations of its vertices' displacements, so the interval
// bound holds for every point.  The test is conservative (never drops a pair the swept-hull test
// of contact/aabb.hpp keeps that could reach the gap), which is all the line search needs: pairs
// that cannot touch return t_max from ACCD anyway.
struct DispInterval { float lx, ly, lz, hx, hy, hz; };
__device__ __forceinline__ bool box_overlap_relative(const Box &q, const DispInterval &u, const float4 lo, const float4 hi, const float4 ulo, const float4 uhi) {
    const float dlx = fminf(0.f, ulo.x - u.hx), dly = fminf(0.f, ulo.y - u.hy), dlz = fminf(0.f, ulo.z - u.hz);
    const float dhx = fmaxf(0.f, uhi.x - u.lx), dhy = fmaxf(0.f, uhi.y - u.ly), dhz = fmaxf(0.f, uhi.z - u.lz);
    return (q.lx <= hi.x + dhx && q.hx >= lo.x + dlx) && (q.ly <= hi.y + dhy && q.hy >= lo.y + dly) && (q.lz <= hi.z + dhz && q.hz >= lo.z + dlz);
}
template <class F> __device__ __forceinline__ bool bvh_query_relative(const BvhView &b, const Box &q, const DispInterval &u, F &&f) {
    if (b.n == 0) return true;
    if (b.n == 1) {
        if (box_overlap_relative(q, u, b.lo[0], b.hi[0], b.ulo[0], b.uhi[0])) f(b.prim[0]);
        return true;
    }
    int stack[PPF_BVH_STACK];
    int sp = 0;
    if (!box_overlap_relative(q, u, b.lo[0], b.hi[0], b.ulo[0], b.uhi[0])) return true;
    stack[sp++] = 0;
    const int first_leaf = (int)b.n - 1;
    while (sp) {
        const int node = stack[--sp];
        const int2 c = b.child[node];
        const bool hit_l = box_overlap_relative(q, u, b.lo[c.x], b.hi[c.x], b.ulo[c.x], b.uhi[c.x]);
        const bool hit_r = box_overlap_relative(q, u, b.lo[c.y], b.hi[c.y], b.ulo[c.y], b.uhi[c.y]);
        if (hit_l) {
            if (c.x >= first_leaf) f(b.prim[c.x - first_leaf]);
            else stack[sp++] = c.x;
        }
        if (hit_r) {
            if (c.y >= first_leaf) f(b.prim[c.y - first_leaf]);
            else {
                if (sp >= PPF_BVH_STACK) return false;
                stack[sp++] = c.y;
            }
        }
    }
    return true;
}

} // namespace ppf

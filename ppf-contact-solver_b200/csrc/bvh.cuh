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
    // What a traversal reads: per internal node ONE 128-byte record with both children's boxes -- [0..3] left (lo, hi, ulo, uhi),
    // [4..7] right -- and the children's node ids in the .w of the two lo entries.  A visit is then a single cache line and a single
    // memory latency; with the per-node arrays above it was the child pair first and then up to eight sectors in four arrays
    // (the arrays remain: the bottom-up fit merges through them, and the root box is tested from them).
    const float4 *pack;
};

inline BvhView view_of(const Bvh &b) { return BvhView{b.n, b.child.p, b.lo.p, b.hi.p, b.prim_sorted.p, b.ulo.p, b.uhi.p, b.pack.p}; }
__device__ __forceinline__ int pack_child(const float4 lo) { return (int)__float_as_uint(lo.w); }

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
        const float4 *rec = b.pack + 8 * size_t(node);
        const float4 l_lo = rec[0], l_hi = rec[1], r_lo = rec[4], r_hi = rec[5];
        const int2 c = make_int2(pack_child(l_lo), pack_child(r_lo));
        const bool hit_l = box_overlap(q, l_lo, l_hi);
        const bool hit_r = box_overlap(q, r_lo, r_hi);
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
        const float4 *rec = b.pack + 8 * size_t(node);
        const float4 l_lo = rec[0], l_hi = rec[1], l_ulo = rec[2], l_uhi = rec[3], r_lo = rec[4], r_hi = rec[5], r_ulo = rec[6], r_uhi = rec[7];
        const int2 c = make_int2(pack_child(l_lo), pack_child(r_lo));
        const bool hit_l = box_overlap_relative(q, u, l_lo, l_hi, l_ulo, l_uhi);
        const bool hit_r = box_overlap_relative(q, u, r_lo, r_hi, r_ulo, r_uhi);
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

// bvh.cuh -- device view of an LBVH and the stack traversal used by every broad phase.
#pragma once
#include "backend.hpp"

namespace ppf {

struct BvhView {
    unsigned n;               // leaves; internal nodes [0, n-1), leaf k is node (n-1)+k
    const int2 *child;        // internal node -> (left, right)
    const float4 *lo, *hi;    // node boxes
    const unsigned *prim;     // sorted leaf k -> primitive index
};

inline BvhView view_of(const Bvh &b) { return BvhView{b.n, b.child.p, b.lo.p, b.hi.p, b.prim_sorted.p}; }

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

} // namespace ppf

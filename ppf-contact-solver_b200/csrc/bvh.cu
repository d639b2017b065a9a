// bvh.cu -- LBVH build (once per step) and refit (once per Newton iteration) for faces, edges, vertices.
//
// Replaces lbvh::build_{face,edge,vertex}_bvh / update_*_aabb / propagate_aabbs
// (crates/ppf-cts-solver/src/cpp/lbvh/lbvh.cu:863-1097).  Same tree family (30-bit Morton codes,
// Karras 2012 radix tree, ties broken by sorted index) and the SAME leaf boxes (swept hull of x0 and
// x0 + extrapolate*(x1-x0), inflated by 0.5*ghat + offset, lbvh.cu:623-731), but:
//   * one radix sort (CUB) instead of 8 four-bit passes of 3 kernels each;
//   * internal boxes are produced by ONE bottom-up kernel (second-arriver-continues with an atomic
//     flag per node) instead of one launch per tree level plus a D2H copy of the level table;
//   * no level table / depth pass at all (the reference's compute_depths_kernel is O(n * depth)).
//   * for the line search the leaves carry the box at x0 plus the interval of their vertices' sweep
//     displacements (bvh_refit_relative), which the relative-motion traversal of bvh.cuh needs.
// Only conservativeness of the boxes matters for parity (the narrow phase decides), so the tree
// shape is free to differ from the reference's.
#include "bvh.cuh"

namespace ppf {

__device__ __forceinline__ unsigned expand_bits(unsigned v) {
    v = (v | (v << 16)) & 0x030000FF;
    v = (v | (v << 8)) & 0x0300F00F;
    v = (v | (v << 4)) & 0x030C30C3;
    v = (v | (v << 2)) & 0x09249249;
    return v;
}

// order-preserving float <-> uint mapping so scene bounds can use atomicMin/atomicMax
__device__ __forceinline__ unsigned f2o(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float o2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

__device__ __forceinline__ float3 prim_centroid(int kind, unsigned i, const float4 *x, const uint4 *face, const uint2 *edge) {
    if (kind == 0) {
        const uint4 f = face[i];
        const float4 a = x[f.x], b = x[f.y], c = x[f.z];
        return make_float3((a.x + b.x + c.x) / 3.0f, (a.y + b.y + c.y) / 3.0f, (a.z + b.z + c.z) / 3.0f);
    } else if (kind == 1) {
        const uint2 e = edge[i];
        const float4 a = x[e.x], b = x[e.y];
        return make_float3((a.x + b.x) / 2.0f, (a.y + b.y) / 2.0f, (a.z + b.z) / 2.0f);
    }
    const float4 a = x[i];
    return make_float3(a.x, a.y, a.z);
}

__global__ void k_bounds(int kind, unsigned n, const float4 *x, const uint4 *face, const uint2 *edge, unsigned *bounds) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 c = make_float3(0.f, 0.f, 0.f);
    const bool on = i < n;
    if (on) c = prim_centroid(kind, i, x, face, edge);
    float lo[3] = {on ? c.x : FLT_MAX, on ? c.y : FLT_MAX, on ? c.z : FLT_MAX};
    float hi[3] = {on ? c.x : -FLT_MAX, on ? c.y : -FLT_MAX, on ? c.z : -FLT_MAX};
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; ++k) {
            atomicMin(bounds + k, f2o(lo[k]));
            atomicMax(bounds + 3 + k, f2o(hi[k]));
        }
}

__global__ void k_morton(int kind, unsigned n, const float4 *x, const uint4 *face, const uint2 *edge, const unsigned *bounds,
                         unsigned *code, unsigned *prim) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 c = prim_centroid(kind, i, x, face, edge);
    const float lo[3] = {o2f(bounds[0]), o2f(bounds[1]), o2f(bounds[2])};
    const float hi[3] = {o2f(bounds[3]), o2f(bounds[4]), o2f(bounds[5])};
    const float cc[3] = {c.x, c.y, c.z};
    unsigned q[3];
    for (int k = 0; k < 3; ++k) {
        const float ext = hi[k] - lo[k];
        const float sc = ext > 1e-10f ? 1023.0f / ext : 0.0f;
        q[k] = (unsigned)fminf(1023.0f, fmaxf(0.0f, (cc[k] - lo[k]) * sc));
    }
    code[i] = expand_bits(q[0]) | (expand_bits(q[1]) << 1) | (expand_bits(q[2]) << 2);
    prim[i] = i;
}

// Karras 2012: common-prefix length of sorted keys i and j (index as tie-break)
__device__ __forceinline__ int delta(const unsigned *code, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned a = code[i], b = code[j];
    if (a == b) return 32 + __clz((unsigned)i ^ (unsigned)j);
    return __clz(a ^ b);
}

__global__ void k_radix_tree(int n, const unsigned *code, int2 *child, int *parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(code, n, i, i + 1) - delta(code, n, i, i - 1)) > 0 ? 1 : -1;
    const int dmin = delta(code, n, i, i - d);
    int lmax = 2;
    while (delta(code, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta(code, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(code, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(code, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
    child[i] = make_int2(left, right);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

struct LeafParams {
    const float4 *x0, *x1;
    float extrapolate;
    const uint4 *face;
    const uint2 *edge;
    const float4 *cp; // (ghat, offset, friction, mass) of this primitive kind
};

// leaf boxes (lbvh.cu:623-731) + bottom-up merge: the second thread to reach a node owns it
// With ulo/uhi the leaf box is the box at x0 and the node additionally carries the component-wise
// interval of its vertices' sweep displacements extrapolate*(x1-x0) (see bvh_query_relative).
__global__ void k_fit(int kind, unsigned n, LeafParams P, const unsigned *prim_sorted, const int2 *child, const int *parent,
                      float4 *lo, float4 *hi, unsigned *flag, float4 *ulo, float4 *uhi, float4 *pack) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned prim = prim_sorted[k];
    const float4 cp = P.cp[prim];
    const float margin = 0.5f * cp.x + cp.y;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    unsigned vidx[3];
    int nvx;
    if (kind == 0) { const uint4 f = P.face[prim]; vidx[0] = f.x; vidx[1] = f.y; vidx[2] = f.z; nvx = 3; }
    else if (kind == 1) { const uint2 e = P.edge[prim]; vidx[0] = e.x; vidx[1] = e.y; vidx[2] = 0; nvx = 2; }
    else { vidx[0] = prim; vidx[1] = vidx[2] = 0; nvx = 1; }
    const bool rel = ulo != nullptr;
    float un[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ux[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int a = 0; a < nvx; ++a) {
        const float4 y0 = P.x0[vidx[a]], y1 = P.x1[vidx[a]];
        const float y[3] = {y0.x, y0.y, y0.z};
        if (rel) {
            const float u[3] = {P.extrapolate * (y1.x - y0.x), P.extrapolate * (y1.y - y0.y), P.extrapolate * (y1.z - y0.z)};
            for (int d = 0; d < 3; ++d) {
                mn[d] = fminf(mn[d], y[d]); mx[d] = fmaxf(mx[d], y[d]);
                un[d] = fminf(un[d], u[d]); ux[d] = fmaxf(ux[d], u[d]);
            }
        } else {
            const float z[3] = {P.extrapolate * (y1.x - y0.x) + y0.x, P.extrapolate * (y1.y - y0.y) + y0.y, P.extrapolate * (y1.z - y0.z) + y0.z};
            for (int d = 0; d < 3; ++d) {
                mn[d] = fminf(mn[d], fminf(y[d], z[d]));
                mx[d] = fmaxf(mx[d], fmaxf(y[d], z[d]));
            }
        }
    }
    float4 blo = make_float4(mn[0] - margin, mn[1] - margin, mn[2] - margin, 0.f);
    float4 bhi = make_float4(mx[0] + margin, mx[1] + margin, mx[2] + margin, 0.f);
    float4 dlo = make_float4(un[0], un[1], un[2], 0.f), dhi = make_float4(ux[0], ux[1], ux[2], 0.f);
    int node = (int)(n - 1 + k);
    lo[node] = blo;
    hi[node] = bhi;
    if (rel) { ulo[node] = dlo; uhi[node] = dhi; }
    if (n == 1) return;
    int p = parent[node];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(flag + p, 1u) == 0u) return; // first arrival: the sibling's thread will finish this node
        const int2 c = child[p];
        const int other = (c.x == node) ? c.y : c.x;
        // the sibling subtree's box was published before its atomicAdd; read through L2
        const float4 olo = __ldcg(lo + other), ohi = __ldcg(hi + other);
        // the traversal record of p (bvh.cuh): both children's boxes as they stand, ids in lo.w -- written before the merge below
        float4 *rec = pack + 8 * size_t(p);
        const int mine = (c.x == node) ? 0 : 4, theirs = 4 - mine;
        rec[mine + 0] = make_float4(blo.x, blo.y, blo.z, __uint_as_float((unsigned)node));
        rec[mine + 1] = bhi;
        rec[theirs + 0] = make_float4(olo.x, olo.y, olo.z, __uint_as_float((unsigned)other));
        rec[theirs + 1] = ohi;
        blo = make_float4(fminf(blo.x, olo.x), fminf(blo.y, olo.y), fminf(blo.z, olo.z), 0.f);
        bhi = make_float4(fmaxf(bhi.x, ohi.x), fmaxf(bhi.y, ohi.y), fmaxf(bhi.z, ohi.z), 0.f);
        lo[p] = blo;
        hi[p] = bhi;
        if (rel) {
            const float4 qlo = __ldcg(ulo + other), qhi = __ldcg(uhi + other);
            rec[mine + 2] = dlo; rec[mine + 3] = dhi;
            rec[theirs + 2] = qlo; rec[theirs + 3] = qhi;
            dlo = make_float4(fminf(dlo.x, qlo.x), fminf(dlo.y, qlo.y), fminf(dlo.z, qlo.z), 0.f);
            dhi = make_float4(fmaxf(dhi.x, qhi.x), fmaxf(dhi.y, qhi.y), fmaxf(dhi.z, qhi.z), 0.f);
            ulo[p] = dlo;
            uhi[p] = dhi;
        }
        node = p;
        p = parent[p];
    }
}

static void fit_leaves(State &s, Bvh &b, int kind, const LeafParams &P, bool relative) {
    if (b.n == 0) return;
    if (relative) { b.ulo.reserve(2 * size_t(b.n)); b.uhi.reserve(2 * size_t(b.n)); }
    if (b.n > 1) CUDA_OK(cudaMemsetAsync(b.flag.p, 0, sizeof(unsigned) * (b.n - 1), s.stream));
    b.pack.reserve(8 * size_t(b.n));
    k_fit<<<div_up(b.n, 256), 256, 0, s.stream>>>(kind, b.n, P, b.prim_sorted.p, b.child.p, b.parent.p, b.lo.p, b.hi.p, b.flag.p,
                                                   relative ? b.ulo.p : nullptr, relative ? b.uhi.p : nullptr, b.pack.p);
    LAUNCH_COUNT(1);
}
static void fit(State &s, Bvh &b, int kind, const float4 *x0, const float4 *x1, float extrapolate, bool relative = false) {
    LeafParams P{x0, x1, extrapolate, s.face.p, s.edge.p, kind == 0 ? s.fcp.p : (kind == 1 ? s.ecp.p : s.vcp.p)};
    fit_leaves(s, b, kind, P, relative);
}

// Morton codes, sort, Karras tree and the leaf / node boxes of n primitives of any mesh (the simulated one or the static
// collision mesh): kind 0 faces, 1 edges, 2 vertices; cp = (ghat, offset, ., .) per primitive
static void build_tree(State &s, Bvh &b, int kind, unsigned n, const float4 *x, const uint4 *face, const uint2 *edge, const float4 *cp) {
    b.n = n;
    if (n == 0) return;
    cudaStream_t q = s.stream;
    b.morton.reserve(n); b.morton_sorted.reserve(n); b.prim.reserve(n); b.prim_sorted.reserve(n);
    b.child.reserve(n); b.parent.reserve(2 * size_t(n)); b.lo.reserve(2 * size_t(n)); b.hi.reserve(2 * size_t(n)); b.flag.reserve(n);
    unsigned *bounds = s.counters.p + 40; // 6 words
    const unsigned init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    CUDA_OK(cudaMemcpyAsync(bounds, init, sizeof(init), cudaMemcpyHostToDevice, q));
    k_bounds<<<div_up(n, 256), 256, 0, q>>>(kind, n, x, face, edge, bounds);
    k_morton<<<div_up(n, 256), 256, 0, q>>>(kind, n, x, face, edge, bounds, b.morton.p, b.prim.p);
    radix_sort_pairs_u32(s, b.morton.p, b.morton_sorted.p, b.prim.p, b.prim_sorted.p, n, 30); // prims.cu (4 passes of 8 bits)
    if (n > 1) k_radix_tree<<<div_up(n - 1, 256), 256, 0, q>>>((int)n, b.morton_sorted.p, b.child.p, b.parent.p);
    LAUNCH_COUNT(4);
    LeafParams P{x, x, 1.0f, face, edge, cp};
    fit_leaves(s, b, kind, P, false);
}

void bvh_build(State &s, Bvh &b, int kind, const float4 *x) {
    const unsigned n = kind == 0 ? s.nf : (kind == 1 ? s.ne : s.nsv);
    build_tree(s, b, kind, n, x, s.face.p, s.edge.p, kind == 0 ? s.fcp.p : (kind == 1 ? s.ecp.p : s.vcp.p));
    if (n == 0) return;
    s.bvh_cover_eval = true;
    s.bvh_relative = false;
}
// the static collision mesh (lbvh.cu:1099-1115 build_collision_mesh_bvh, plus a vertex tree so that every query of this
// backend runs from a dynamic element against a static tree): built once at initialize
void bvh_build_static(State &s, Bvh &b, int kind, unsigned n, const float4 *x, const uint4 *face, const uint2 *edge, const float4 *cp) {
    build_tree(s, b, kind, n, x, face, edge, cp);
}

void bvh_refit(State &s, Bvh &b, int kind, const float4 *x0, const float4 *x1, float extrapolate) {
    fit(s, b, kind, x0, x1, extrapolate);
    s.bvh_cover_eval = true; // also true for a swept hull: eval_x lies on the swept segment
    s.bvh_relative = false;
}
void bvh_refit_relative(State &s, Bvh &b, int kind, const float4 *x0, const float4 *x1, float t_max) {
    fit(s, b, kind, x0, x1, t_max, true);
    s.bvh_cover_eval = false;
    s.bvh_relative = true;
}

} // namespace ppf

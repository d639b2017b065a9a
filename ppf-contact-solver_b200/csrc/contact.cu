// contact.cu -- broad phase, contact-pair generation, barrier assembly, ACCD line search,
// intersection check.
//
// Replaces contact::embed_contact_force_hessian (crates/ppf-cts-solver/src/cpp/contact/contact.cu:1398-1787
// with its functors 432-887 and embed_contact_force_hess<N> 263-354), contact::line_search
// (2472-2740, CCD functors 2095-2316) and contact::check_intersection (3056-3175), plus the
// DynCSRMat count/fill/finalize protocol of csrmat/csrmat.cu:19-488.
//
// Execution model (B200): the reference runs the narrow phase INSIDE a per-thread BVH traversal
// (divergent, latency bound) and walks its dynamic matrix through per-row pointer structs.  Here
//   1. a broad-phase kernel per query kind only emits candidate pairs (box overlap + index filters),
//      queries issued in Morton order so a warp walks the same subtree;
//   2. narrow-phase kernels run one candidate per thread (uniform work, coalesced candidate reads)
//      with the reference's fp32 predicates (math_distance.cuh) and compact accepted pairs with a
//      warp scan + one atomic per warp;
//   3. the dynamic matrix pattern is the sorted unique set of (row, col) keys the accepted pairs
//      touch outside the fixed pattern (one radix sort), values are accumulated by bisection into it.
//   4. the line search sweeps with start boxes + displacement intervals (relative-motion test, bvh.cuh)
//      instead of swept hulls, and the next assembly runs its narrow phase on the sweep's candidate
//      lists (the evaluation point lies on the swept segment), so there is one traversal per Newton
//      iteration instead of two.
// Pair acceptance is independent of the candidate superset, so the accepted sets equal the
// reference's for the same positions.
#include "bvh.cuh"
#include "dist.hpp"
#include "math_accd.cuh"
#include "math_collider.cuh"
#include "math_cmesh.cuh"
#include "math_contact.cuh"
#include "math_distance.cuh"
#include "warp_prims.cuh"
#include <algorithm>
#include <chrono>

namespace ppf {

__device__ __forceinline__ V3 ld3c(const float4 *p, unsigned i) { const float4 v = p[i]; return mk(v.x, v.y, v.z); }
__device__ __forceinline__ float sqr(float x) { return x * x; }

// ---------------------------------------------------------------------------------------------
// broad phase
// ---------------------------------------------------------------------------------------------
struct Emit { // per-thread staging of candidate pairs, flushed with one atomic per 8 pairs
    uint2 buf[8];
    unsigned n = 0;
    uint2 *out;
    unsigned *counter;
    unsigned cap;
    __device__ __forceinline__ void flush() {
        if (!n) return;
        const unsigned base = atomicAdd(counter, n);
        for (unsigned k = 0; k < n; ++k)
            if (base + k < cap) out[base + k] = buf[k];
        n = 0;
    }
    __device__ __forceinline__ void push(unsigned a, unsigned b) {
        buf[n++] = make_uint2(a, b);
        if (n == 8) flush();
    }
};

struct BpArgs {
    BvhView face, edge, vert;
    const float4 *x0, *x1; // query positions; sweep to x0 + sweep_t*(x1-x0) when sweep
    float sweep_t;
    int sweep;             // 0: assembly boxes around x0, 1: CCD swept boxes
    const float4 *vcp, *ecp;
    const uint4 *faces;
    const uint2 *edges;
    uint2 *out[4];
    unsigned *cnt; // 4 counters
    unsigned cap[4];
    unsigned *error;
    const unsigned char *active; // collision windows: per-vertex flags, or null (everything takes part)
};
// collision windows (main.cu:48-110, 2266-2330): an edge / face takes part when any of its vertices does
__device__ __forceinline__ bool cw_vertex(const unsigned char *a, unsigned v) { return !a || a[v]; }
__device__ __forceinline__ bool cw_edge(const unsigned char *a, const uint2 e) { return !a || a[e.x] || a[e.y]; }
__device__ __forceinline__ bool cw_face(const unsigned char *a, const uint4 f) { return !a || a[f.x] || a[f.y] || a[f.z]; }

__global__ void __launch_bounds__(128) k_bp_vertex(BpArgs A, unsigned q_lo, unsigned nq) { // Morton-ordered queries [q_lo, nq)
    const unsigned t = q_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const unsigned v = A.vert.prim[t];
    if (!cw_vertex(A.active, v)) return;
    const float4 cp = A.vcp[v];
    const float ext = 0.5f * cp.x + cp.y;
    const float4 p0 = A.x0[v];
    Box q;
    q.lx = p0.x - ext; q.ly = p0.y - ext; q.lz = p0.z - ext;
    q.hx = p0.x + ext; q.hy = p0.y + ext; q.hz = p0.z + ext;
    DispInterval u = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int sweep = A.sweep;
    const bool rel = sweep == 1;
    if (rel) { // relative-motion sweep (bvh.cuh): start box + displacement interval of the query
        const float4 p1 = A.x1[v];
        u.lx = u.hx = A.sweep_t * (p1.x - p0.x); u.ly = u.hy = A.sweep_t * (p1.y - p0.y); u.lz = u.hz = A.sweep_t * (p1.z - p0.z);
    } else if (sweep == 2) { // the reference's swept hull (A/B switch PPF_B200_SWEPT_HULL, lbvh.cu:623-731)
        const float4 p1 = A.x1[v];
        const float zx = A.sweep_t * (p1.x - p0.x) + p0.x, zy = A.sweep_t * (p1.y - p0.y) + p0.y, zz = A.sweep_t * (p1.z - p0.z) + p0.z;
        q.lx = fminf(p0.x, zx) - ext; q.ly = fminf(p0.y, zy) - ext; q.lz = fminf(p0.z, zz) - ext;
        q.hx = fmaxf(p0.x, zx) + ext; q.hy = fmaxf(p0.y, zy) + ext; q.hz = fmaxf(p0.z, zz) + ext;
    }
    bool ok = true;
    {
        Emit em; em.out = A.out[0]; em.counter = A.cnt + 0; em.cap = A.cap[0];
        auto emit = [&](unsigned f) {
            if (!sweep) { const uint4 fv = A.faces[f]; if (fv.x == v || fv.y == v || fv.z == v) return; }
            if (A.active && !cw_face(A.active, A.faces[f])) return;
            em.push(v, f);
        };
        ok &= rel ? bvh_query_relative(A.face, q, u, emit) : bvh_query(A.face, q, emit);
        em.flush();
    }
    { // vertex-edge: only the assembly evaluates these pairs; the sweep lists them so that the next
      // assembly can reuse the sweep's candidates instead of traversing again (detect_contacts)
        Emit em; em.out = A.out[1]; em.counter = A.cnt + 1; em.cap = A.cap[1];
        auto emit = [&](unsigned e) {
            const uint2 ev = A.edges[e];
            if (ev.x == v || ev.y == v || !cw_edge(A.active, ev)) return;
            em.push(v, e);
        };
        ok &= rel ? bvh_query_relative(A.edge, q, u, emit) : bvh_query(A.edge, q, emit);
        em.flush();
    }
    {
        Emit em; em.out = A.out[2]; em.counter = A.cnt + 2; em.cap = A.cap[2];
        auto emit = [&](unsigned w) { if (w < v && cw_vertex(A.active, w)) em.push(v, w); };
        ok &= rel ? bvh_query_relative(A.vert, q, u, emit) : bvh_query(A.vert, q, emit);
        em.flush();
    }
    if (!ok) atomicExch(A.error, 1u);
}

__global__ void __launch_bounds__(128) k_bp_edge(BpArgs A, unsigned q_lo, unsigned nq) {
    const unsigned t = q_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const unsigned e = A.edge.prim[t];
    const uint2 ev = A.edges[e];
    if (!cw_edge(A.active, ev)) return;
    const float4 cp = A.ecp[e];
    const float ext = 0.5f * cp.x + cp.y;
    const float4 a0 = A.x0[ev.x], b0 = A.x0[ev.y];
    Box q;
    q.lx = fminf(a0.x, b0.x) - ext; q.ly = fminf(a0.y, b0.y) - ext; q.lz = fminf(a0.z, b0.z) - ext;
    q.hx = fmaxf(a0.x, b0.x) + ext; q.hy = fmaxf(a0.y, b0.y) + ext; q.hz = fmaxf(a0.z, b0.z) + ext;
    DispInterval u = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int sweep = A.sweep;
    const bool rel = sweep == 1;
    if (sweep) {
        const float4 a1 = A.x1[ev.x], b1 = A.x1[ev.y];
        const float s = A.sweep_t;
        const float ax = s * (a1.x - a0.x), ay = s * (a1.y - a0.y), az = s * (a1.z - a0.z);
        const float bx = s * (b1.x - b0.x), by = s * (b1.y - b0.y), bz = s * (b1.z - b0.z);
        if (rel) {
            u.lx = fminf(ax, bx); u.ly = fminf(ay, by); u.lz = fminf(az, bz);
            u.hx = fmaxf(ax, bx); u.hy = fmaxf(ay, by); u.hz = fmaxf(az, bz);
        } else { // swept hull
            q.lx = fminf(q.lx, fminf(ax + a0.x, bx + b0.x) - ext); q.ly = fminf(q.ly, fminf(ay + a0.y, by + b0.y) - ext); q.lz = fminf(q.lz, fminf(az + a0.z, bz + b0.z) - ext);
            q.hx = fmaxf(q.hx, fmaxf(ax + a0.x, bx + b0.x) + ext); q.hy = fmaxf(q.hy, fmaxf(ay + a0.y, by + b0.y) + ext); q.hz = fmaxf(q.hz, fmaxf(az + a0.z, bz + b0.z) + ext);
        }
    }
    Emit em; em.out = A.out[3]; em.counter = A.cnt + 3; em.cap = A.cap[3];
    auto emit = [&](unsigned f) {
        if (!(e < f)) return;
        if (!sweep) { const uint2 fv = A.edges[f]; if (fv.x == ev.x || fv.x == ev.y || fv.y == ev.x || fv.y == ev.y) return; }
        if (A.active && !cw_edge(A.active, A.edges[f])) return;
        em.push(e, f);
    };
    const bool ok = rel ? bvh_query_relative(A.edge, q, u, emit) : bvh_query(A.edge, q, emit);
    em.flush();
    if (!ok) atomicExch(A.error, 1u);
}

// several ranks: are contact detection and the line search sharded (default) or replicated (PPF_B200_DIST_CONTACT=0)?
static bool contact_sharded() {
    static const bool off = getenv("PPF_B200_DIST_CONTACT") && getenv("PPF_B200_DIST_CONTACT")[0] == '0';
    return g_dist.on() && !off;
}

// runs the broad phase; grows candidate buffers and retries on overflow
static void broad_phase(State &s, const float4 *x0, const float4 *x1, bool sweep, float sweep_t) {
    cudaStream_t q = s.stream;
    unsigned *cnt = s.counters.p; // [0..3] candidate counters, [4] error
    const size_t init_cap[4] = {size_t(24) * s.nsv + 1024, size_t(32) * s.nsv + 1024, size_t(8) * s.nsv + 1024, size_t(24) * s.ne + 1024};
    for (int k = 0; k < 4; ++k)
        if (s.cand[k].pairs.cap < init_cap[k]) s.cand[k].pairs.reserve(init_cap[k]);
    for (int attempt = 0; attempt < 8; ++attempt) {
        CUDA_OK(cudaMemsetAsync(cnt, 0, 8 * sizeof(unsigned), q));
        BpArgs A;
        A.face = view_of(s.bvh_face); A.edge = view_of(s.bvh_edge); A.vert = view_of(s.bvh_vert);
        A.x0 = x0; A.x1 = x1; A.sweep_t = sweep_t; A.sweep = sweep ? (s.bvh_relative ? 1 : 2) : 0;
        A.vcp = s.vcp.p; A.ecp = s.ecp.p; A.faces = s.face.p; A.edges = s.edge.p;
        for (int k = 0; k < 4; ++k) { A.out[k] = s.cand[k].pairs.p; A.cap[k] = (unsigned)std::min<size_t>(s.cand[k].pairs.cap, 0xFFFFFFF0u); }
        A.cnt = cnt; A.error = cnt + 4;
        A.active = s.has_cw ? s.cw_active.p : nullptr;
        // several ranks: this rank's slice of the query elements (dist.hpp); the candidate lists are then this rank's share
        unsigned v_lo = 0, v_hi = s.nsv, e_lo = 0, e_hi = s.ne;
        if (contact_sharded()) {
            const unsigned long long r = (unsigned long long)g_dist.rank, w = (unsigned long long)g_dist.world;
            v_lo = (unsigned)(s.nsv * r / w); v_hi = (unsigned)(s.nsv * (r + 1) / w);
            e_lo = (unsigned)(s.ne * r / w); e_hi = (unsigned)(s.ne * (r + 1) / w);
        }
        if (v_hi > v_lo) k_bp_vertex<<<div_up(v_hi - v_lo, 128), 128, 0, q>>>(A, v_lo, v_hi);
        if (e_hi > e_lo) k_bp_edge<<<div_up(e_hi - e_lo, 128), 128, 0, q>>>(A, e_lo, e_hi);
        LAUNCH_COUNT(2);
        unsigned h[5];
        CUDA_OK(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, q));
        CUDA_OK(cudaStreamSynchronize(q));
        if (h[4]) fatal(4, "BVH traversal stack overflow (depth > %d)", PPF_BVH_STACK);
        bool again = false;
        for (int k = 0; k < 4; ++k) {
            s.cand[k].count = h[k];
            if (h[k] > s.cand[k].pairs.cap) { s.cand[k].pairs.reserve(size_t(h[k]) + h[k] / 4); again = true; }
        }
        if (!again) return;
    }
    fatal(4, "broad phase: candidate buffers kept overflowing");
}

// ---------------------------------------------------------------------------------------------
// narrow phase predicates (shared by detection and fill)
// ---------------------------------------------------------------------------------------------
struct Ctx {
    const float4 *x;    // eval_x
    const float4 *x0;   // start-of-step positions (data.vertex.curr)
    const float4 *vcp, *ecp, *fcp;
    const uint4 *face;
    const uint2 *edge;
    const unsigned char *vfixed, *efixed, *ffixed, *vcollider;
    const unsigned *vf_off, *vf_idx, *ve_off, *ve_idx, *ef_off, *ef_idx;
    int friction_mode;
    float friction_eps;
    int barrier; // ParamSet.barrier: 0 cubic, 1 quad, 2 log
    // matrices
    const unsigned *fx_off, *fx_col;
    const float *fx_el; // elastic-only snapshot
    float *fx_val;      // accumulating fixed matrix
    const unsigned *dyn_off, *dyn_col;
    float *dyn_val;
    float *force;
};

// per-object contact statistics (contact.cu:379-430 of the reference): a contact between objects a and b counts once for each of
// them, once in total when both sides belong to the same object; 0xFFFFFFFF = the vertex belongs to no object
struct StatView {
    const unsigned *obj, *sobj; // per vertex / per static-mesh vertex
    unsigned *count;            // per object, null = statistics off
    unsigned n_obj;
};
__device__ __forceinline__ void stat_record(const StatView &st, unsigned oa, unsigned ob) {
    if (oa != 0xFFFFFFFFu && oa < st.n_obj) atomicAdd(st.count + oa, 1u);
    if (ob != 0xFFFFFFFFu && ob != oa && ob < st.n_obj) atomicAdd(st.count + ob, 1u);
}
static StatView make_stat_view(State &s) {
    StatView st;
    st.obj = s.stat_obj.p; st.sobj = s.stat_sobj.p; st.count = s.n_stat_obj ? s.stat_count.p : nullptr; st.n_obj = s.n_stat_obj;
    return st;
}

template <int N> struct Prox {
    unsigned idx[N];
    float w[N];
    float ghat, offset, friction;
    unsigned count;
    bool include_friction;
};

__device__ __forceinline__ bool eval_pf(const Ctx &c, unsigned v, unsigned fi, Prox<4> &o) {
    const uint4 f = c.face[fi];
    if (f.x == v || f.y == v || f.z == v) return false;
    const bool either_dyn = !c.vfixed[v] || !c.ffixed[fi];
    const bool both_collider = c.vcollider[v] && c.vcollider[f.x];
    if (!either_dyn || both_collider) return false;
    const V3 p = ld3c(c.x, v), t0 = ld3c(c.x, f.x), t1 = ld3c(c.x, f.y), t2 = ld3c(c.x, f.z);
    const C3 cf = point_triangle_coeff(p, t0, t1, t2);
    if (!strictly_interior(cf)) return false;
    const V3 e = lin3(cf.a, p - t0, cf.b, p - t1, cf.c, p - t2);
    const float4 vp = c.vcp[v], fp = c.fcp[fi];
    const float offset = vp.y + fp.y;
    const float ghat = 0.5f * (vp.x + fp.x);
    if (!(sqnorm(e) < sqr(ghat + offset))) return false;
    o.idx[0] = v; o.idx[1] = f.x; o.idx[2] = f.y; o.idx[3] = f.z;
    o.w[0] = 1.0f; o.w[1] = -cf.a; o.w[2] = -cf.b; o.w[3] = -cf.c;
    o.ghat = ghat; o.offset = offset; o.friction = combine_friction(vp.z, fp.z, c.friction_mode);
    o.count = 1; o.include_friction = true;
    return true;
}

__device__ __forceinline__ bool eval_pe(const Ctx &c, unsigned v, unsigned ei, Prox<3> &o) {
    const uint2 f = c.edge[ei];
    if (f.x == v || f.y == v) return false;
    const bool either_dyn = !c.vfixed[v] || !c.efixed[ei];
    const bool both_collider = c.vcollider[v] && c.vcollider[f.x];
    if (!either_dyn || both_collider) return false;
    const V3 p = ld3c(c.x, v), t0 = ld3c(c.x, f.x), t1 = ld3c(c.x, f.y);
    const C2 cf = point_edge_coeff(p, t0, t1);
    if (!strictly_interior(cf)) return false;
    const V3 e = lin2(cf.a, p - t0, cf.b, p - t1);
    const float4 vp = c.vcp[v], ep = c.ecp[ei];
    const float offset = vp.y + ep.y;
    const float ghat = 0.5f * (vp.x + ep.x);
    if (!(sqnorm(e) < sqr(ghat + offset))) return false;
    unsigned count = 0;
    bool include_friction = true;
    const unsigned b0 = c.ef_off[ei], b1 = c.ef_off[ei + 1];
    if (b1 > b0) {
        for (unsigned j = b0; j < b1; ++j) {
            const uint4 ff = c.face[c.ef_idx[j]];
            if (ff.x == v || ff.y == v || ff.z == v) { ++count; include_friction = false; }
            else {
                const C3 cc = point_triangle_coeff(p, ld3c(c.x, ff.x), ld3c(c.x, ff.y), ld3c(c.x, ff.z));
                if (strictly_interior(cc)) continue;
                ++count;
            }
        }
    } else count = 1u;
    if (!count) return false;
    o.idx[0] = v; o.idx[1] = f.x; o.idx[2] = f.y;
    o.w[0] = 1.0f; o.w[1] = -cf.a; o.w[2] = -cf.b;
    o.ghat = ghat; o.offset = offset; o.friction = combine_friction(vp.z, ep.z, c.friction_mode);
    o.count = count; o.include_friction = include_friction;
    return true;
}

__device__ __forceinline__ bool eval_pp(const Ctx &c, unsigned v, unsigned w, Prox<2> &o) {
    if (!(w < v)) return false;
    const bool either_dyn = !c.vfixed[v] || !c.vfixed[w];
    const bool both_collider = c.vcollider[v] && c.vcollider[w];
    if (!either_dyn || both_collider) return false;
    const V3 x = ld3c(c.x, v), y = ld3c(c.x, w);
    const V3 e = x - y;
    const float4 vp = c.vcp[v], wp = c.vcp[w];
    const float offset = vp.y + wp.y;
    const float ghat = 0.5f * (vp.x + wp.x);
    if (!(sqnorm(e) < sqr(ghat + offset))) return false;
    unsigned count = 0;
    bool include_friction = true;
    const unsigned e0 = c.ve_off[w], e1 = c.ve_off[w + 1];
    if (e1 > e0) {
        for (unsigned j = e0; j < e1; ++j) {
            const uint2 f = c.edge[c.ve_idx[j]];
            if (f.x == v || f.y == v) { ++count; include_friction = false; }
            else {
                const C2 cc = point_edge_coeff(x, ld3c(c.x, f.x), ld3c(c.x, f.y));
                if (strictly_interior(cc)) continue;
                ++count;
            }
        }
        const unsigned f0 = c.vf_off[w], f1 = c.vf_off[w + 1];
        for (unsigned j = f0; j < f1; ++j) {
            const uint4 f = c.face[c.vf_idx[j]];
            if (f.x == v || f.y == v || f.z == v) { ++count; include_friction = false; }
            else {
                const C3 cc = point_triangle_coeff(x, ld3c(c.x, f.x), ld3c(c.x, f.y), ld3c(c.x, f.z));
                if (strictly_interior(cc)) continue;
                ++count;
            }
        }
    } else count = 1u;
    if (!count) return false;
    o.idx[0] = v; o.idx[1] = w;
    o.w[0] = 1.0f; o.w[1] = -1.0f;
    o.ghat = ghat; o.offset = offset; o.friction = combine_friction(vp.z, wp.z, c.friction_mode);
    o.count = count; o.include_friction = include_friction;
    return true;
}

__device__ __forceinline__ bool eval_ee(const Ctx &c, unsigned ea, unsigned eb, Prox<4> &o) {
    if (!(ea < eb)) return false;
    const uint2 e0 = c.edge[ea], e1 = c.edge[eb];
    if (e0.x == e1.x || e0.x == e1.y || e0.y == e1.x || e0.y == e1.y) return false;
    const bool either_dyn = !c.efixed[ea] || !c.efixed[eb];
    const bool both_collider = c.vcollider[e0.x] && c.vcollider[e1.x];
    if (!either_dyn || both_collider) return false;
    const V3 p0 = ld3c(c.x, e0.x), p1 = ld3c(c.x, e0.y), q0 = ld3c(c.x, e1.x), q1 = ld3c(c.x, e1.y);
    const C4 cf = edge_edge_coeff(p0, p1, q0, q1);
    if (!strictly_interior(cf)) return false;
    const V3 xa = lin2(cf.a, p0, cf.b, p1);
    const V3 xb = lin2(cf.c, q0, cf.d, q1);
    const V3 e = xa - xb;
    const float4 ap = c.ecp[ea], bp = c.ecp[eb];
    const float offset = ap.y + bp.y;
    const float ghat = 0.5f * (ap.x + bp.x);
    if (!(sqnorm(e) < sqr(ghat + offset))) return false;
    o.idx[0] = e0.x; o.idx[1] = e0.y; o.idx[2] = e1.x; o.idx[3] = e1.y;
    o.w[0] = cf.a; o.w[1] = cf.b; o.w[2] = -cf.c; o.w[3] = -cf.d;
    o.ghat = ghat; o.offset = offset; o.friction = combine_friction(ap.z, bp.z, c.friction_mode);
    o.count = 1; o.include_friction = true;
    return true;
}

template <int KIND> __global__ void __launch_bounds__(256) k_narrow(Ctx c, const uint2 *cand, unsigned n, uint2 *out, unsigned *counter, unsigned cap) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    bool acc = false;
    uint2 pr = make_uint2(0, 0);
    if (t < n) {
        pr = cand[t];
        if (KIND == 0) { Prox<4> o; acc = eval_pf(c, pr.x, pr.y, o); }
        else if (KIND == 1) { Prox<3> o; acc = eval_pe(c, pr.x, pr.y, o); }
        else if (KIND == 2) { Prox<2> o; acc = eval_pp(c, pr.x, pr.y, o); }
        else { Prox<4> o; acc = eval_ee(c, pr.x, pr.y, o); }
    }
    const unsigned slot = warp_alloc(counter, acc ? 1u : 0u);
    if (acc && slot < cap) out[slot] = pr;
}

// ---------------------------------------------------------------------------------------------
// matrix lookups
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned fixed_find(const Ctx &c, unsigned i, unsigned j) { // i <= j
    const unsigned b0 = c.fx_off[i], b1 = c.fx_off[i + 1];
    for (unsigned k = b0; k < b1; ++k) {
        const unsigned col = c.fx_col[k];
        if (col == j) return k;
        if (col > j) break;
    }
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ unsigned dyn_find(const Ctx &c, unsigned i, unsigned j) {
    unsigned lo = c.dyn_off[i], hi = c.dyn_off[i + 1];
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        const unsigned col = c.dyn_col[mid];
        if (col < j) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <int N> __device__ __forceinline__ unsigned emit_keys(const Ctx &c, const unsigned *idx, unsigned long long *keys, unsigned shift) {
    unsigned n = 0;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = a + 1; b < N; ++b) {
            const unsigned i = min(idx[a], idx[b]), j = max(idx[a], idx[b]);
            if (fixed_find(c, i, j) == 0xFFFFFFFFu) keys[n++] = ((unsigned long long)i << shift) | j;
        }
    return n;
}

// one count per accepted pair, attributed through the vertices the reference names (contact.cu:617, 722, 800, 879)
template <int KIND> __global__ void __launch_bounds__(256) k_stats(Ctx c, StatView st, const uint2 *pairs, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint2 pr = pairs[t];
    unsigned va, vb;
    if (KIND == 0) { va = pr.x; vb = c.face[pr.y].x; }
    else if (KIND == 1) { va = pr.x; vb = c.edge[pr.y].x; }
    else if (KIND == 2) { va = pr.x; vb = pr.y; }
    else { va = c.edge[pr.x].x; vb = c.edge[pr.y].x; }
    stat_record(st, st.obj[va], st.obj[vb]);
}

template <int KIND> __global__ void __launch_bounds__(256) k_keys(Ctx c, const uint2 *pairs, unsigned n, unsigned long long *out, unsigned *counter, unsigned shift) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long keys[6];
    unsigned k = 0;
    if (t < n) {
        const uint2 pr = pairs[t];
        if (KIND == 0) { const uint4 f = c.face[pr.y]; const unsigned idx[4] = {pr.x, f.x, f.y, f.z}; k = emit_keys<4>(c, idx, keys, shift); }
        else if (KIND == 1) { const uint2 e = c.edge[pr.y]; const unsigned idx[3] = {pr.x, e.x, e.y}; k = emit_keys<3>(c, idx, keys, shift); }
        else if (KIND == 2) { const unsigned idx[2] = {pr.x, pr.y}; k = emit_keys<2>(c, idx, keys, shift); }
        else { const uint2 a = c.edge[pr.x], b = c.edge[pr.y]; const unsigned idx[4] = {a.x, a.y, b.x, b.y}; k = emit_keys<4>(c, idx, keys, shift); }
    }
    const unsigned base = warp_alloc(counter, k);
    for (unsigned q = 0; q < k; ++q) out[base + q] = keys[q];
}

__global__ void k_dyn_keep(unsigned nnz, const float *__restrict__ val, unsigned char *__restrict__ keep) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    bool zero = true;
#pragma unroll
    for (int e = 0; e < 9; ++e)
        if (!(fabsf(val[size_t(9) * k + e]) <= 1e-5f)) zero = false;
    keep[k] = zero ? 0 : 1;
}

__global__ void k_dyn_structure(unsigned nv, unsigned nnz, const unsigned long long *keys, unsigned shift, unsigned *off, unsigned *col) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nnz) col[t] = (unsigned)(keys[t] & ((1ull << shift) - 1ull));
    if (t <= nv) { // off[t] = first key with row >= t
        const unsigned long long target = (unsigned long long)t << shift;
        unsigned lo = 0, hi = nnz;
        while (lo < hi) {
            const unsigned mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        off[t] = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// fill: barrier force / Hessian of one contact (embed_contact_force_hess<N>, contact.cu:263-354)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ M3 load_block(const float *val, unsigned slot) {
    M3 m;
    const float *p = val + size_t(9) * slot;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.m[i] = p[i];
    return m;
}

template <int N> __device__ __forceinline__ void embed_contact(const Ctx &c, const Prox<N> &pr) {
    float mass[N];
    V3 ex0 = mk(0.f, 0.f, 0.f), ex = mk(0.f, 0.f, 0.f);
    const V3 anchor0 = ld3c(c.x0, pr.idx[0]), anchor = ld3c(c.x, pr.idx[0]);
#pragma unroll
    for (int ii = 0; ii < N; ++ii) {
        mass[ii] = c.vcp[pr.idx[ii]].w;
        ex0 += pr.w[ii] * (ld3c(c.x0, pr.idx[ii]) - anchor0);
        ex += pr.w[ii] * (ld3c(c.x, pr.idx[ii]) - anchor);
    }
    const V3 dxs = ex - ex0;
    const V3 normal = normalized(ex);
    // ---- elasticity-inclusive stiffness (barrier.cu:50-86), upper blocks gathered once ----
    const float nrm = norm(ex);
    const float g = nrm - pr.offset;
    const float sqr_x = g * g;
    float m_max = 0.0f;
#pragma unroll
    for (int ii = 0; ii < N; ++ii) m_max = fmaxf(m_max, mass[ii]);
    V3 wv[N];
    float wsq = 0.0f;
#pragma unroll
    for (int ii = 0; ii < N; ++ii) { wv[ii] = pr.w[ii] * ex; wsq += sqnorm(wv[ii]); }
    const float wn = sqrtf(wsq);
    if (wn > 0.0f) {
#pragma unroll
        for (int ii = 0; ii < N; ++ii) wv[ii] = wv[ii] / wn;
    }
    float kappa = 0.0f;
#pragma unroll
    for (int ii = 0; ii < N; ++ii) {
        const float m_ii = mass[ii] > 0.0f ? mass[ii] : m_max;
        kappa += (m_ii / sqr_x) * sqnorm(wv[ii]);
#pragma unroll
        for (int jj = ii; jj < N; ++jj) {
            const unsigned i = pr.idx[ii], j = pr.idx[jj];
            const bool swap = i > j;
            const unsigned slot = fixed_find(c, swap ? j : i, swap ? i : j);
            if (slot == 0xFFFFFFFFu) continue;
            const M3 B = load_block(c.fx_el, slot);
            // block (ii,jj) of the local Hessian is B (or B^T when stored as (j,i))
            const V3 Bw = swap ? mul_t(B, wv[jj]) : mul(B, wv[jj]);
            const float t = dot(wv[ii], Bw);
            kappa += (ii == jj) ? t : 2.0f * t;
        }
    }
    V3 f = kappa * barrier_edge_gradient(ex, pr.ghat, pr.offset, c.barrier);
    M3 H = kappa * barrier_edge_hessian(ex, pr.ghat, pr.offset, c.barrier);
    if (pr.include_friction) {
        const FrictionOut fr = friction_eval(f, dxs, normal, pr.friction, c.friction_eps);
        f += fr.grad;
        H += fr.hess;
    }
    const float cnt = (float)pr.count;
#pragma unroll
    for (int ii = 0; ii < N; ++ii) {
        const V3 fi = (cnt * pr.w[ii]) * f;
        float *dst = c.force + size_t(3) * pr.idx[ii];
        if (fi.x) atomicAdd(dst + 0, fi.x);
        if (fi.y) atomicAdd(dst + 1, fi.y);
        if (fi.z) atomicAdd(dst + 2, fi.z);
    }
#pragma unroll
    for (int ii = 0; ii < N; ++ii)
#pragma unroll
        for (int jj = 0; jj < N; ++jj) {
            const unsigned i = pr.idx[ii], j = pr.idx[jj];
            if (!(i <= j)) continue;
            const float sc = cnt * (pr.w[ii] * pr.w[jj]);
            float *dst;
            const unsigned slot = fixed_find(c, i, j);
            if (slot != 0xFFFFFFFFu) dst = c.fx_val + size_t(9) * slot;
            else dst = c.dyn_val + size_t(9) * dyn_find(c, i, j);
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                const float y = sc * H.m[e];
                if (y) atomicAdd(dst + e, y);
            }
        }
}

template <int KIND> __global__ void __launch_bounds__(128) k_fill(Ctx c, const uint2 *pairs, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint2 pr = pairs[t];
    if (KIND == 0) { Prox<4> o; if (eval_pf(c, pr.x, pr.y, o)) embed_contact<4>(c, o); }
    else if (KIND == 1) { Prox<3> o; if (eval_pe(c, pr.x, pr.y, o)) embed_contact<3>(c, o); }
    else if (KIND == 2) { Prox<2> o; if (eval_pp(c, pr.x, pr.y, o)) embed_contact<2>(c, o); }
    else { Prox<4> o; if (eval_ee(c, pr.x, pr.y, o)) embed_contact<4>(c, o); }
}

static Ctx make_ctx(State &s, const PpfParamSet &prm) {
    Ctx c;
    c.x = s.eval_x.p; c.x0 = s.curr.p;
    c.vcp = s.vcp.p; c.ecp = s.ecp.p; c.fcp = s.fcp.p;
    c.face = s.face.p; c.edge = s.edge.p;
    c.vfixed = s.vfixed.p; c.efixed = s.efixed.p; c.ffixed = s.ffixed.p; c.vcollider = s.vcollider.p;
    c.vf_off = s.vf_off.p; c.vf_idx = s.vf_idx.p; c.ve_off = s.ve_off.p; c.ve_idx = s.ve_idx.p; c.ef_off = s.ef_off.p; c.ef_idx = s.ef_idx.p;
    c.friction_mode = (int)prm.friction_mode; c.friction_eps = prm.friction_eps; c.barrier = (int)prm.barrier;
    c.fx_off = s.fx_off.p; c.fx_col = s.fx_col.p; c.fx_el = s.fx_el.p; c.fx_val = s.fx_val.p;
    c.dyn_off = s.dyn_off.p; c.dyn_col = s.dyn_col.p; c.dyn_val = s.dyn_val.p;
    c.force = s.force.p;
    return c;
}


// broad + narrow phase at eval_x: fills s.pairs[0..3] (unsorted) and returns the four counts
static std::vector<uint2> sorted_pairs_host(State &s, int k) {
    std::vector<uint2> h(s.pairs[k].count);
    if (!h.empty()) CUDA_OK(cudaMemcpy(h.data(), s.pairs[k].pairs.p, sizeof(uint2) * h.size(), cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const uint2 &a, const uint2 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    return h;
}

static unsigned detect_contacts_impl(State &s, const PpfParamSet &prm, unsigned hc[4], bool allow_reuse);

// PPF_B200_VERIFY_CAND=1 (tests): whenever the sweep's candidates are reused, detect again with a fresh
// static traversal at eval_x and demand the two accepted pair sets to be identical.
static unsigned detect_contacts(State &s, const PpfParamSet &prm, unsigned hc[4]) {
    static const bool verify = getenv("PPF_B200_VERIFY_CAND") != nullptr;
    static const bool no_reuse = getenv("PPF_B200_NO_CAND_REUSE") != nullptr;
    const bool reuse = s.cand_cover_eval && !no_reuse;
    if (!(verify && reuse)) return detect_contacts_impl(s, prm, hc, reuse);
    unsigned h1[4];
    detect_contacts_impl(s, prm, h1, true);
    std::vector<uint2> a[4];
    for (int k = 0; k < 4; ++k) a[k] = sorted_pairs_host(s, k);
    const unsigned total = detect_contacts_impl(s, prm, hc, false);
    s.cand_cover_eval = false; // the candidate lists now are the static ones (still valid for eval_x; keep it simple)
    for (int k = 0; k < 4; ++k) {
        const std::vector<uint2> b = sorted_pairs_host(s, k);
        bool same = a[k].size() == b.size();
        for (size_t i = 0; same && i < b.size(); ++i) same = a[k][i].x == b[i].x && a[k][i].y == b[i].y;
        if (!same) fatal(4, "candidate reuse changed the accepted pair set of kind %d: %zu (reused sweep candidates) vs %zu (fresh traversal)", k, a[k].size(), b.size());
    }
    static unsigned long long n_ok = 0;
    if ((++n_ok % 16) == 1) fprintf(stderr, "[verify-cand] %llu assemblies verified identical (last: %u pairs)\n", n_ok, total);
    return total;
}

// Several ranks: every rank accepted the pairs of its own query slice; all-gather the four lists so that every rank assembles
// the whole contact set (the union over the ranks IS the single-GPU set: each pair has exactly one query side).  hc: in = this
// rank's counts, out = the totals.  Rank-major order; the lists are unordered anyway (slots are handed out by atomics).
static DBuf<uint2> g_pair_stage;
static void gather_pair_lists(State &s, unsigned hc[4]) {
    cudaStream_t q = s.stream;
    const unsigned W = (unsigned)g_dist.world;
    unsigned *d_cnt = s.counters.p + 64; // [64 .. 64 + 4 W): counts of every rank
    CUDA_OK(cudaMemcpyAsync(d_cnt + 4 * g_dist.rank, hc, 4 * sizeof(unsigned), cudaMemcpyHostToDevice, q));
    dist_allgather(d_cnt + 4 * g_dist.rank, d_cnt, 4 * sizeof(unsigned), q);
    std::vector<unsigned> all(4 * W);
    CUDA_OK(cudaMemcpyAsync(all.data(), d_cnt, sizeof(unsigned) * 4 * W, cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    for (int k = 0; k < 4; ++k) {
        unsigned mx = 0;
        unsigned long long tot = 0;
        for (unsigned r = 0; r < W; ++r) { mx = std::max(mx, all[4 * r + k]); tot += all[4 * r + k]; }
        if (tot > 0xFFFFFFF0ull) fatal(4, "contact pair list of kind %d exceeds 2^32 entries over the ranks", k);
        hc[k] = (unsigned)tot;
        if (!mx) continue;
        s.pairs[k].pairs.reserve_keep(std::max<size_t>(mx, tot), q); // send buffer: at least the largest share (the tail is never read as pairs)
        g_pair_stage.reserve(size_t(mx) * W);
        dist_allgather(s.pairs[k].pairs.p, g_pair_stage.p, sizeof(uint2) * size_t(mx), q);
        size_t at = 0;
        for (unsigned r = 0; r < W; ++r) {
            const unsigned c = all[4 * r + k];
            if (c) CUDA_OK(cudaMemcpyAsync(s.pairs[k].pairs.p + at, g_pair_stage.p + size_t(mx) * r, sizeof(uint2) * c, cudaMemcpyDeviceToDevice, q));
            at += c;
        }
    }
}

static unsigned detect_contacts_impl(State &s, const PpfParamSet &prm, unsigned hc[4], bool allow_reuse) {
    cudaStream_t q = s.stream;
    // The candidate lists of the last line-search sweep x_old -> x_old + t_max*(x_new - x_old) cover every
    // pair that is within the contact gap anywhere on that segment (bvh.cuh), and eval_x is a point of
    // the segment (x_old + toi*(x_new - x_old), toi <= 1 <= t_max): no second traversal is needed.
    if (!allow_reuse) {
        if (!s.bvh_cover_eval) { // the trees hold start boxes of the last relative sweep: refit at eval_x first
            bvh_refit(s, s.bvh_face, 0, s.eval_x.p, s.eval_x.p, 1.0f);
            bvh_refit(s, s.bvh_edge, 1, s.eval_x.p, s.eval_x.p, 1.0f);
            bvh_refit(s, s.bvh_vert, 2, s.eval_x.p, s.eval_x.p, 1.0f);
        }
        broad_phase(s, s.eval_x.p, s.eval_x.p, false, 0.0f);
    }
    Ctx c = make_ctx(s, prm);
    unsigned *acc_cnt = s.counters.p + 8; // [8..11]
    CUDA_OK(cudaMemsetAsync(acc_cnt, 0, 8 * sizeof(unsigned), q));
    for (int k = 0; k < 4; ++k) {
        const unsigned n = s.cand[k].count;
        s.pairs[k].pairs.reserve(std::max<size_t>(n, 16));
        if (!n) continue;
        const unsigned cap = (unsigned)std::min<size_t>(s.pairs[k].pairs.cap, 0xFFFFFFF0u);
        const unsigned grid = div_up(n, 256);
        if (k == 0) k_narrow<0><<<grid, 256, 0, q>>>(c, s.cand[k].pairs.p, n, s.pairs[k].pairs.p, acc_cnt + k, cap);
        if (k == 1) k_narrow<1><<<grid, 256, 0, q>>>(c, s.cand[k].pairs.p, n, s.pairs[k].pairs.p, acc_cnt + k, cap);
        if (k == 2) k_narrow<2><<<grid, 256, 0, q>>>(c, s.cand[k].pairs.p, n, s.pairs[k].pairs.p, acc_cnt + k, cap);
        if (k == 3) k_narrow<3><<<grid, 256, 0, q>>>(c, s.cand[k].pairs.p, n, s.pairs[k].pairs.p, acc_cnt + k, cap);
        LAUNCH_COUNT(1);
    }
    CUDA_OK(cudaMemcpyAsync(hc, acc_cnt, 4 * sizeof(unsigned), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    if (contact_sharded()) gather_pair_lists(s, hc);
    unsigned total = 0;
    for (int k = 0; k < 4; ++k) { s.pairs[k].count = hc[k]; s.last_num_contact[k] = hc[k]; total += hc[k]; }
    return total;
}

unsigned assemble_contact(State &s, float dt, const PpfParamSet &prm, unsigned &max_nnz_row, float &dyn_consumed) {
    (void)dt;
    cudaStream_t q = s.stream;
    max_nnz_row = 0;
    dyn_consumed = 0.0f;
    // PPF_B200_CONTACT_PROF=1: wall-clock split of this call (stream drained at the marks), diagnostic only
    static const bool prof = getenv("PPF_B200_CONTACT_PROF") != nullptr;
    auto now = [&] { if (prof) cudaStreamSynchronize(q); return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point t0, std::chrono::steady_clock::time_point t1) { return std::chrono::duration<double, std::milli>(t1 - t0).count(); };
    const auto t0 = now();
    const bool reused = s.cand_cover_eval;
    // ---- detect ----
    unsigned hc[4];
    const unsigned total = detect_contacts(s, prm, hc);
    if (s.n_stat_obj) { // the counts were reset by the step (main.cu:825); collision mesh and colliders add to them afterwards
        const StatView st = make_stat_view(s);
        const Ctx cs = make_ctx(s, prm);
        if (hc[0]) k_stats<0><<<div_up(hc[0], 256), 256, 0, q>>>(cs, st, s.pairs[0].pairs.p, hc[0]);
        if (hc[1]) k_stats<1><<<div_up(hc[1], 256), 256, 0, q>>>(cs, st, s.pairs[1].pairs.p, hc[1]);
        if (hc[2]) k_stats<2><<<div_up(hc[2], 256), 256, 0, q>>>(cs, st, s.pairs[2].pairs.p, hc[2]);
        if (hc[3]) k_stats<3><<<div_up(hc[3], 256), 256, 0, q>>>(cs, st, s.pairs[3].pairs.p, hc[3]);
        LAUNCH_COUNT(4);
    }
    Ctx c = make_ctx(s, prm);
    const auto t1 = now();

    // ---- dynamic pattern: sorted unique keys outside the fixed pattern ----
    unsigned shift = 1;
    while ((1ull << shift) < (unsigned long long)s.nv) ++shift;
    const size_t max_keys = size_t(6) * hc[0] + size_t(3) * hc[1] + size_t(1) * hc[2] + size_t(6) * hc[3] + 32;
    s.dyn_keys.reserve(max_keys); s.dyn_keys_sorted.reserve(max_keys);
    s.dyn_off.reserve(s.nv + 2);
    unsigned *key_cnt = s.counters.p + 16;
    CUDA_OK(cudaMemsetAsync(key_cnt, 0, 2 * sizeof(unsigned), q));
    if (hc[0]) k_keys<0><<<div_up(hc[0], 256), 256, 0, q>>>(c, s.pairs[0].pairs.p, hc[0], s.dyn_keys.p, key_cnt, shift);
    if (hc[1]) k_keys<1><<<div_up(hc[1], 256), 256, 0, q>>>(c, s.pairs[1].pairs.p, hc[1], s.dyn_keys.p, key_cnt, shift);
    if (hc[2]) k_keys<2><<<div_up(hc[2], 256), 256, 0, q>>>(c, s.pairs[2].pairs.p, hc[2], s.dyn_keys.p, key_cnt, shift);
    if (hc[3]) k_keys<3><<<div_up(hc[3], 256), 256, 0, q>>>(c, s.pairs[3].pairs.p, hc[3], s.dyn_keys.p, key_cnt, shift);
    LAUNCH_COUNT(4);
    unsigned nkeys = 0;
    CUDA_OK(cudaMemcpyAsync(&nkeys, key_cnt, sizeof(unsigned), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    unsigned nnz = 0;
    if (nkeys) {
        radix_sort_keys_u64(s, s.dyn_keys.p, s.dyn_keys_sorted.p, nkeys, (int)(2 * shift));   // prims.cu
        unique_sorted_u64(s, s.dyn_keys_sorted.p, s.dyn_keys.p, nkeys, key_cnt + 1);
        CUDA_OK(cudaMemcpyAsync(&nnz, key_cnt + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, q));
        CUDA_OK(cudaStreamSynchronize(q));
    }
    if (nnz >= s.dyn_cap) fatal(4, "dynamic contact matrix needs %u blocks but csrmat-max-nnz is %llu", nnz, s.dyn_cap);
    s.dyn_nnz = nnz;
    s.dyn_col.reserve(std::max<size_t>(nnz, 16)); s.dyn_val.reserve(size_t(9) * std::max<size_t>(nnz, 16));
    k_dyn_structure<<<div_up(std::max<size_t>(nnz, s.nv + 1), 256), 256, 0, q>>>(s.nv, nnz, s.dyn_keys.p, shift, s.dyn_off.p, s.dyn_col.p);
    if (nnz) CUDA_OK(cudaMemsetAsync(s.dyn_val.p, 0, sizeof(float) * 9 * size_t(nnz), q));
    LAUNCH_COUNT(1);
    dyn_consumed = (float)nnz / (float)s.dyn_cap;

    const auto t2 = now();
    // ---- fill ----
    c = make_ctx(s, prm); // buffers may have been re-allocated
    if (hc[0]) k_fill<0><<<div_up(hc[0], 128), 128, 0, q>>>(c, s.pairs[0].pairs.p, hc[0]);
    if (hc[1]) k_fill<1><<<div_up(hc[1], 128), 128, 0, q>>>(c, s.pairs[1].pairs.p, hc[1]);
    if (hc[2]) k_fill<2><<<div_up(hc[2], 128), 128, 0, q>>>(c, s.pairs[2].pairs.p, hc[2]);
    if (hc[3]) k_fill<3><<<div_up(hc[3], 128), 128, 0, q>>>(c, s.pairs[3].pairs.p, hc[3]);
    LAUNCH_COUNT(4);
    // which blocks the reference's finalize() would keep (row_dedupe.hpp:96-150 drops blocks that are zero in SMat::isZero's
    // sense) -- decided HERE, on the freshly accumulated values: the Dirichlet elimination zeroes the blocks of pinned rows and
    // columns afterwards (main.cu:954-1049), and the reference's carried pattern (fetch_dyn) still lists those
    s.dyn_keep.reserve(std::max<size_t>(nnz, 16));
    if (nnz) {
        k_dyn_keep<<<div_up(nnz, 256), 256, 0, q>>>(nnz, s.dyn_val.p, s.dyn_keep.p);
        LAUNCH_COUNT(1);
    }
    s.dyn_valid = true;
    if (prof) {
        const auto t3 = now();
        fprintf(stderr, "[contact-prof] detect %.2f ms (reuse %d; candidates %u %u %u %u -> pairs %u %u %u %u)  pattern %.2f ms (%u keys -> %u blocks)  fill %.2f ms\n",
                ms_since(t0, t1), (int)reused, s.cand[0].count, s.cand[1].count, s.cand[2].count, s.cand[3].count, hc[0], hc[1], hc[2], hc[3], ms_since(t1, t2), nkeys, nnz,
                ms_since(t2, t3));
    }
    return total;
}

// pattern handed out by fetch_dyn: per row, the sorted columns whose accumulated block is not
// "zero" in the reference's sense (SMat::isZero, |m| <= 1e-5 for all entries; row_dedupe.hpp:96-150), see k_dyn_keep
void build_fetch_pattern(State &s, std::vector<unsigned> &index, std::vector<unsigned> &offset) {
    const unsigned nv = s.nv, nnz = s.dyn_nnz;
    offset.assign(nv + 1, 0);
    index.clear();
    if (!nnz) return;
    std::vector<unsigned> off(nv + 1), col(nnz);
    std::vector<unsigned char> keep(nnz);
    CUDA_OK(cudaMemcpy(off.data(), s.dyn_off.p, sizeof(unsigned) * (nv + 1), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(col.data(), s.dyn_col.p, sizeof(unsigned) * nnz, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(keep.data(), s.dyn_keep.p, nnz, cudaMemcpyDeviceToHost));
    for (unsigned i = 0; i < nv; ++i) {
        offset[i] = (unsigned)index.size();
        for (unsigned k = off[i]; k < off[i + 1]; ++k)
            if (keep[k]) index.push_back(col[k]);
    }
    offset[nv] = (unsigned)index.size();
}

// ---------------------------------------------------------------------------------------------
// ACCD line search
// ---------------------------------------------------------------------------------------------
struct CcdCtx {
    const float4 *x0, *x1;
    const float4 *vcp, *ecp, *fcp;
    const uint4 *face;
    const uint2 *edge;
    const unsigned char *vfixed, *efixed, *ffixed, *vcollider;
    float max_t, ccd_eps;
    unsigned *toi_bits; // positive floats order like unsigned ints
    unsigned *overlap;  // [0] flag, [1] v0, [2] v1, [3] kind
};

__device__ __forceinline__ void ccd_commit(const CcdCtx &c, float result, const CcdOverlap &ov, unsigned v0, unsigned v1, unsigned kind) {
    if (ov.flag) atomicExch(c.overlap, 1u);
    if (result < c.max_t) {
        atomicMin(c.toi_bits, __float_as_uint(result));
        if (result == 0.0f && atomicCAS(c.overlap + 1, 0xFFFFFFFFu, v0) == 0xFFFFFFFFu) { c.overlap[2] = v1; c.overlap[3] = kind; }
    }
}

__global__ void __launch_bounds__(128) k_ccd_pf(CcdCtx c, const uint2 *cand, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned v = cand[t].x, fi = cand[t].y;
    const uint4 f = c.face[fi];
    const bool either_dyn = !c.vfixed[v] || !c.ffixed[fi];
    const bool both_collider = c.vcollider[v] && c.vcollider[f.x];
    if (!either_dyn || both_collider) return;
    const unsigned fv[3] = {f.x, f.y, f.z};
    int dup = -1;
    for (int i = 0; i < 3; ++i)
        if (fv[i] == v) { dup = i; break; }
    CcdOverlap ov{0u, -1.0f, -1.0f};
    float result;
    const float4 vp = c.vcp[v];
    if (dup == -1) {
        const float4 fp = c.fcp[fi];
        const float offset = vp.y + fp.y, ghat = 0.5f * (vp.x + fp.x);
        result = point_triangle_ccd(ld3c(c.x0, v), ld3c(c.x1, v), ld3c(c.x0, f.x), ld3c(c.x0, f.y), ld3c(c.x0, f.z), ld3c(c.x1, f.x),
                                    ld3c(c.x1, f.y), ld3c(c.x1, f.z), offset, ghat, c.max_t, c.ccd_eps, ov);
    } else {
        const float offset = 2.0f * vp.y, ghat = vp.x;
        const unsigned i = fv[dup], j = fv[(dup + 1) % 3], k = fv[(dup + 2) % 3];
        result = point_edge_ccd(ld3c(c.x0, i), ld3c(c.x1, i), ld3c(c.x0, j), ld3c(c.x0, k), ld3c(c.x1, j), ld3c(c.x1, k), offset, ghat,
                                c.max_t, c.ccd_eps, ov);
    }
    ccd_commit(c, result, ov, v, f.x, 0u);
}

__global__ void __launch_bounds__(128) k_ccd_pp(CcdCtx c, const uint2 *cand, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned v = cand[t].x, w = cand[t].y;
    const bool either_dyn = !c.vfixed[v] || !c.vfixed[w];
    const bool both_collider = c.vcollider[v] && c.vcollider[w];
    if (!(w < v) || !either_dyn || both_collider) return;
    const float4 a = c.vcp[v], b = c.vcp[w];
    CcdOverlap ov{0u, -1.0f, -1.0f};
    const float result = point_point_ccd(ld3c(c.x0, v), ld3c(c.x1, v), ld3c(c.x0, w), ld3c(c.x1, w), a.y + b.y, 0.5f * (a.x + b.x),
                                         c.max_t, c.ccd_eps, ov);
    ccd_commit(c, result, ov, v, w, 2u);
}

__global__ void __launch_bounds__(128) k_ccd_ee(CcdCtx c, const uint2 *cand, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned ea = cand[t].x, eb = cand[t].y;
    const bool either_dyn = !c.efixed[ea] || !c.efixed[eb];
    const uint2 e0 = c.edge[ea], e1 = c.edge[eb];
    const bool both_collider = c.vcollider[e0.x] && c.vcollider[e1.x];
    if (!(ea < eb) || !either_dyn || both_collider) return;
    const float4 ap = c.ecp[ea], bp = c.ecp[eb];
    const float offset = ap.y + bp.y, ghat = 0.5f * (ap.x + bp.x);
    CcdOverlap ov{0u, -1.0f, -1.0f};
    float result = c.max_t;
    const unsigned a[2] = {e0.x, e0.y}, b[2] = {e1.x, e1.y};
    const bool shared = a[0] == b[0] || a[0] == b[1] || a[1] == b[0] || a[1] == b[1];
    if (!shared) {
        result = edge_edge_ccd(ld3c(c.x0, a[0]), ld3c(c.x0, a[1]), ld3c(c.x0, b[0]), ld3c(c.x0, b[1]), ld3c(c.x1, a[0]), ld3c(c.x1, a[1]),
                               ld3c(c.x1, b[0]), ld3c(c.x1, b[1]), offset, ghat, c.max_t, c.ccd_eps, ov);
    } else {
        const int ij[4][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}};
        for (int k = 0; k < 4; ++k) {
            const int i = ij[k][0], j = ij[k][1];
            if (a[i] == b[j]) {
                const unsigned i0 = a[i], i1 = a[1 - i], i2 = b[1 - j];
                const float t0 = point_edge_ccd(ld3c(c.x0, i1), ld3c(c.x1, i1), ld3c(c.x0, i0), ld3c(c.x0, i2), ld3c(c.x1, i0), ld3c(c.x1, i2),
                                                offset, ghat, c.max_t, c.ccd_eps, ov);
                const float t1 = point_edge_ccd(ld3c(c.x0, i2), ld3c(c.x1, i2), ld3c(c.x0, i0), ld3c(c.x0, i1), ld3c(c.x1, i0), ld3c(c.x1, i1),
                                                offset, ghat, c.max_t, c.ccd_eps, ov);
                result = fminf(t0, t1);
                break;
            }
        }
    }
    ccd_commit(c, result, ov, e0.x, e1.x, 1u);
}


float contact_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm, bool &overlap) {
    cudaStream_t q = s.stream;
    broad_phase(s, x0, x1, true, prm.line_search_max_t);
    s.cand_cover_eval = prm.line_search_max_t >= 1.0f; // valid for eval_x = x0 + toi*(x1 - x0) after the lerp
    CcdCtx c;
    c.x0 = x0; c.x1 = x1; c.vcp = s.vcp.p; c.ecp = s.ecp.p; c.fcp = s.fcp.p; c.face = s.face.p; c.edge = s.edge.p;
    c.vfixed = s.vfixed.p; c.efixed = s.efixed.p; c.ffixed = s.ffixed.p; c.vcollider = s.vcollider.p;
    c.max_t = prm.line_search_max_t; c.ccd_eps = prm.ccd_eps;
    unsigned *w = s.counters.p + 20; // [20] toi bits, [21..24] overlap info
    const float max_t = prm.line_search_max_t;
    unsigned init[5];
    memcpy(&init[0], &max_t, sizeof(float));
    init[1] = 0u; init[2] = 0xFFFFFFFFu; init[3] = 0xFFFFFFFFu; init[4] = 0xFFFFFFFFu;
    CUDA_OK(cudaMemcpyAsync(w, init, sizeof(init), cudaMemcpyHostToDevice, q));
    c.toi_bits = w; c.overlap = w + 1;
    if (s.cand[0].count) k_ccd_pf<<<div_up(s.cand[0].count, 128), 128, 0, q>>>(c, s.cand[0].pairs.p, s.cand[0].count);
    if (s.cand[2].count) k_ccd_pp<<<div_up(s.cand[2].count, 128), 128, 0, q>>>(c, s.cand[2].pairs.p, s.cand[2].count);
    if (s.cand[3].count) k_ccd_ee<<<div_up(s.cand[3].count, 128), 128, 0, q>>>(c, s.cand[3].pairs.p, s.cand[3].count);
    LAUNCH_COUNT(3);
    if (contact_sharded()) { // every rank swept its own slice of the queries: the step fraction is the minimum over the ranks
        dist_allreduce_u32(w, 1, false, q);     // positive floats order like unsigned ints
        dist_allreduce_u32(w + 1, 1, true, q);  // "starts overlapping" flag
    }
    unsigned h[5];
    CUDA_OK(cudaMemcpyAsync(h, w, sizeof(h), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    float toi;
    memcpy(&toi, &h[0], sizeof(float));
    overlap = h[1] != 0u;
    return toi / prm.line_search_max_t;
}

// ---------------------------------------------------------------------------------------------
// analytic colliders: invisible spheres / bowls and walls (contact.cu:1154-1383, 2318-2470; math_collider.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_collider(unsigned nsv, const float4 *__restrict__ eval_x, const float4 *__restrict__ curr,
                                                  const float4 *__restrict__ vcp, const unsigned *__restrict__ fix_index,
                                                  const float *__restrict__ fx_el, const unsigned *__restrict__ diag_slot,
                                                  const PpfSphere *__restrict__ spheres, unsigned n_sphere, const PpfFloor *__restrict__ floors,
                                                  unsigned n_floor, int friction_mode, float friction_eps, float constraint_tol,
                                                  float *__restrict__ force, float *__restrict__ fx_val, unsigned *counter, StatView stat) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned count = 0;
    if (i < nsv && fix_index[i] == 0) {
        const float4 cp = vcp[i]; // (ghat, offset, friction, mass)
        if (cp.w > 0.0f) {       // zero-mass vertices are static solids
            const V3 x = ld3c(eval_x, i);
            const unsigned slot = diag_slot[i];
            M3 Hii;
#pragma unroll
            for (int e = 0; e < 9; ++e) Hii.m[e] = fx_el[size_t(9) * slot + e];
            ColliderOut o;
            collider_vertex(x, x - ld3c(curr, i), cp.w, cp.z, Hii, spheres, n_sphere, floors, n_floor, friction_mode, friction_eps, constraint_tol, o);
            count = o.count;
            if (count && stat.count && stat.obj[i] != 0xFFFFFFFFu && stat.obj[i] < stat.n_obj) atomicAdd(stat.count + stat.obj[i], count); // record_analytic_contact per contact
            if (count) { // one thread owns vertex i in this kernel, but the contact kernels may still be adding to it
                float *fo = force + size_t(3) * i;
                if (o.f.x) atomicAdd(fo + 0, o.f.x);
                if (o.f.y) atomicAdd(fo + 1, o.f.y);
                if (o.f.z) atomicAdd(fo + 2, o.f.z);
                float *dst = fx_val + size_t(9) * slot;
#pragma unroll
                for (int e = 0; e < 9; ++e)
                    if (o.H.m[e]) atomicAdd(dst + e, o.H.m[e]);
            }
        }
    }
    warp_sum_add(counter, count);
}

__global__ void __launch_bounds__(128) k_collider_toi(unsigned nsv, const float4 *__restrict__ x0, const float4 *__restrict__ x1,
                                                      const float4 *__restrict__ vcp, const unsigned *__restrict__ fix_index,
                                                      const PpfSphere *__restrict__ spheres, unsigned n_sphere, const PpfFloor *__restrict__ floors,
                                                      unsigned n_floor, float max_t, unsigned *toi_bits, unsigned *bad_vertex) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    float t = max_t;
    if (i < nsv) {
        const bool pinned = fix_index[i] != 0;
        if (pinned || vcp[i].w > 0.0f) {
            bool infeasible = false;
            t = collider_toi(ld3c(x0, i), ld3c(x1, i), pinned, spheres, n_sphere, floors, n_floor, max_t, &infeasible);
            if (infeasible) atomicMin(bad_vertex, i);
        }
    }
    for (int o = 16; o > 0; o >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if ((threadIdx.x & 31) == 0 && t < max_t) atomicMin(toi_bits, __float_as_uint(fmaxf(t, 0.0f)));
}

unsigned assemble_colliders(State &s, const PpfParamSet &prm) {
    if (!(s.nsphere || s.nfloor) || !s.nsv) return 0;
    cudaStream_t q = s.stream;
    unsigned *cnt = s.counters.p + 30;
    CUDA_OK(cudaMemsetAsync(cnt, 0, sizeof(unsigned), q));
    k_collider<<<div_up(s.nsv, 128), 128, 0, q>>>(s.nsv, s.eval_x.p, s.curr.p, s.vcp.p, s.fix_index.p, s.fx_el.p, s.fx_diag_slot.p, s.spheres.p, s.nsphere,
                                                  s.floors.p, s.nfloor, (int)prm.friction_mode, prm.friction_eps, prm.constraint_tol, s.force.p, s.fx_val.p, cnt, make_stat_view(s));
    LAUNCH_COUNT(1);
    unsigned h = 0;
    CUDA_OK(cudaMemcpyAsync(&h, cnt, sizeof(unsigned), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    return h;
}

float collider_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm, unsigned &bad_vertex) {
    bad_vertex = 0xFFFFFFFFu;
    if (!(s.nsphere || s.nfloor) || !s.nsv) return 1.0f;
    cudaStream_t q = s.stream;
    unsigned *w = s.counters.p + 32; // [32] toi bits, [33] smallest infeasible pinned vertex
    const float max_t = prm.line_search_max_t;
    unsigned init[2];
    memcpy(&init[0], &max_t, sizeof(float));
    init[1] = 0xFFFFFFFFu;
    CUDA_OK(cudaMemcpyAsync(w, init, sizeof(init), cudaMemcpyHostToDevice, q));
    k_collider_toi<<<div_up(s.nsv, 128), 128, 0, q>>>(s.nsv, x0, x1, s.vcp.p, s.fix_index.p, s.spheres.p, s.nsphere, s.floors.p, s.nfloor, max_t, w, w + 1);
    LAUNCH_COUNT(1);
    unsigned h[2];
    CUDA_OK(cudaMemcpyAsync(h, w, sizeof(h), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    float t;
    memcpy(&t, &h[0], sizeof(float));
    bad_vertex = h[1];
    return t / max_t;
}

// ---------------------------------------------------------------------------------------------
// static collision mesh (Constraint.mesh; contact.cu:899-1151 embeds, 1826-1955 assembly driver, 1956-2093 + 2596-2728 line
// search; math_cmesh.cuh).  Every query runs from a dynamic element against a tree of the static mesh built once at
// initialize: dynamic vertex -> static faces, dynamic edge -> static edges, dynamic face -> static vertices (the reference
// runs the last one from the static vertex against the dynamic face tree; box overlap is symmetric, so the pairs are the same).
// ---------------------------------------------------------------------------------------------
struct CmArgs {
    BvhView sface, sedge, svert;
    const float4 *sx, *svcp, *secp, *sfcp;
    const uint4 *sfaces;
    const uint2 *sedges;
    const unsigned *fix_index;
    const unsigned char *active; // collision windows (see BpArgs)
    unsigned *counter;           // [0] contact count, [1] traversal error
    StatView stat;               // per-object contact statistics (count == null: off)
};
static CmArgs make_cm_args(State &s, unsigned *counter) {
    CmArgs A;
    A.stat = make_stat_view(s);
    A.sface = view_of(s.cm_bvh_face); A.sedge = view_of(s.cm_bvh_edge); A.svert = view_of(s.cm_bvh_vert);
    A.sx = s.cm_x.p; A.svcp = s.cm_vcp.p; A.secp = s.cm_ecp.p; A.sfcp = s.cm_fcp.p; A.sfaces = s.cm_face.p; A.sedges = s.cm_edge.p;
    A.fix_index = s.fix_index.p;
    A.active = s.has_cw ? s.cw_active.p : nullptr;
    A.counter = counter;
    return A;
}
// block (i, j) of the elastic-only matrix, transposed when it is stored as (j, i); zero when the pattern lacks it
__device__ __forceinline__ M3 elastic_block(const Ctx &c, unsigned i, unsigned j) {
    const bool swap = i > j;
    const unsigned slot = fixed_find(c, swap ? j : i, swap ? i : j);
    if (slot == 0xFFFFFFFFu) return m3_zero();
    const M3 B = load_block(c.fx_el, slot);
    return swap ? transpose(B) : B;
}
template <int N> __device__ __forceinline__ void cm_scatter(const Ctx &c, const unsigned *idx, const CmeshOut &o) {
#pragma unroll
    for (int a = 0; a < N; ++a) {
        const V3 fa = o.c[a] * o.f;
        float *dst = c.force + size_t(3) * idx[a];
        if (fa.x) atomicAdd(dst + 0, fa.x);
        if (fa.y) atomicAdd(dst + 1, fa.y);
        if (fa.z) atomicAdd(dst + 2, fa.z);
    }
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) {
            const unsigned i = idx[a], j = idx[b];
            if (!(i <= j)) continue;
            const unsigned slot = fixed_find(c, i, j);
            if (slot == 0xFFFFFFFFu) continue; // cannot happen: the blocks of one element are in the fixed pattern
            const float sc = o.c[a] * o.c[b];
            float *dst = c.fx_val + size_t(9) * slot;
#pragma unroll
            for (int e = 0; e < 9; ++e) {
                const float v = sc * o.H.m[e];
                if (v) atomicAdd(dst + e, v);
            }
        }
}
__device__ __forceinline__ void cm_count(const CmArgs &A, unsigned count, bool ok) {
    if (count) atomicAdd(A.counter, count);
    if (!ok) atomicExch(A.counter + 1, 1u);
}

__global__ void __launch_bounds__(128) k_cm_vertex(Ctx c, CmArgs A, unsigned nsv) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsv || A.fix_index[i] != 0 || !cw_vertex(A.active, i)) return;
    const float4 cp = c.vcp[i];
    if (!(cp.w > 0.0f)) return;
    const V3 p = ld3c(c.x, i), ps = ld3c(c.x0, i);
    const float ext = 0.5f * cp.x + cp.y;
    Box q;
    q.lx = p.x - ext; q.ly = p.y - ext; q.lz = p.z - ext; q.hx = p.x + ext; q.hy = p.y + ext; q.hz = p.z + ext;
    const M3 Hii = elastic_block(c, i, i);
    unsigned count = 0;
    const bool ok = bvh_query(A.sface, q, [&](unsigned fi) {
        const uint4 f = A.sfaces[fi];
        const float4 sp = A.sfcp[fi];
        CmeshOut o;
        if (!cmesh_vertex_face(p, ps, ld3c(A.sx, f.x), ld3c(A.sx, f.y), ld3c(A.sx, f.z), cp.y + sp.y, 0.5f * (cp.x + sp.x),
                               combine_friction(cp.z, sp.z, c.friction_mode), cp.w, Hii, c.friction_eps, o)) return;
        const unsigned idx[1] = {i};
        cm_scatter<1>(c, idx, o);
        if (A.stat.count) stat_record(A.stat, A.stat.obj[i], A.stat.sobj ? A.stat.sobj[f.x] : 0xFFFFFFFFu); // contact.cu:956
        ++count;
    });
    cm_count(A, count, ok);
}

__global__ void __launch_bounds__(128) k_cm_edge(Ctx c, CmArgs A, unsigned ne) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne || c.efixed[i]) return;
    const float4 cp = c.ecp[i];
    const uint2 ev = c.edge[i];
    if (!(cp.w > 0.0f) || !cw_edge(A.active, ev)) return;
    const V3 P[2] = {ld3c(c.x, ev.x), ld3c(c.x, ev.y)}, Ps[2] = {ld3c(c.x0, ev.x), ld3c(c.x0, ev.y)};
    const float ext = 0.5f * cp.x + cp.y;
    Box q;
    q.lx = fminf(P[0].x, P[1].x) - ext; q.ly = fminf(P[0].y, P[1].y) - ext; q.lz = fminf(P[0].z, P[1].z) - ext;
    q.hx = fmaxf(P[0].x, P[1].x) + ext; q.hy = fmaxf(P[0].y, P[1].y) + ext; q.hz = fmaxf(P[0].z, P[1].z) + ext;
    const unsigned idx[2] = {ev.x, ev.y};
    const float vmass[2] = {c.vcp[ev.x].w, c.vcp[ev.y].w};
    M3 B[2][2];
    bool have = false;
    unsigned count = 0;
    const bool ok = bvh_query(A.sedge, q, [&](unsigned ei) {
        if (!have) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) B[a][b] = elastic_block(c, idx[a], idx[b]);
            have = true;
        }
        const uint2 se = A.sedges[ei];
        const float4 sp = A.secp[ei];
        CmeshOut o;
        if (!cmesh_edge_edge(P, Ps, ld3c(A.sx, se.x), ld3c(A.sx, se.y), cp.y + sp.y, 0.5f * (cp.x + sp.x), combine_friction(cp.z, sp.z, c.friction_mode),
                             vmass, B, c.friction_eps, o)) return;
        cm_scatter<2>(c, idx, o);
        if (A.stat.count) stat_record(A.stat, A.stat.obj[ev.x], A.stat.sobj ? A.stat.sobj[se.x] : 0xFFFFFFFFu); // contact.cu:1146
        ++count;
    });
    cm_count(A, count, ok);
}

__global__ void __launch_bounds__(128) k_cm_face(Ctx c, CmArgs A, unsigned nf) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || c.ffixed[i]) return;
    const float4 cp = c.fcp[i];
    const uint4 fv = c.face[i];
    if (!(cp.w > 0.0f) || !cw_face(A.active, fv)) return;
    const V3 X[3] = {ld3c(c.x, fv.x), ld3c(c.x, fv.y), ld3c(c.x, fv.z)}, Xs[3] = {ld3c(c.x0, fv.x), ld3c(c.x0, fv.y), ld3c(c.x0, fv.z)};
    const float ext = 0.5f * cp.x + cp.y;
    Box q;
    q.lx = fminf(X[0].x, fminf(X[1].x, X[2].x)) - ext; q.ly = fminf(X[0].y, fminf(X[1].y, X[2].y)) - ext; q.lz = fminf(X[0].z, fminf(X[1].z, X[2].z)) - ext;
    q.hx = fmaxf(X[0].x, fmaxf(X[1].x, X[2].x)) + ext; q.hy = fmaxf(X[0].y, fmaxf(X[1].y, X[2].y)) + ext; q.hz = fmaxf(X[0].z, fmaxf(X[1].z, X[2].z)) + ext;
    const unsigned idx[3] = {fv.x, fv.y, fv.z};
    const float vmass[3] = {c.vcp[fv.x].w, c.vcp[fv.y].w, c.vcp[fv.z].w};
    M3 B[3][3];
    bool have = false;
    unsigned count = 0;
    const bool ok = bvh_query(A.svert, q, [&](unsigned vi) {
        if (!have) {
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) B[a][b] = elastic_block(c, idx[a], idx[b]);
            have = true;
        }
        const float4 sp = A.svcp[vi];
        CmeshOut o;
        if (!cmesh_face_vertex(ld3c(A.sx, vi), X, Xs, cp.y + sp.y, 0.5f * (cp.x + sp.x), combine_friction(cp.z, sp.z, c.friction_mode), vmass, B,
                               c.friction_eps, o)) return;
        cm_scatter<3>(c, idx, o);
        if (A.stat.count) stat_record(A.stat, A.stat.obj[fv.x], A.stat.sobj ? A.stat.sobj[vi] : 0xFFFFFFFFu); // contact.cu:1051
        ++count;
    });
    cm_count(A, count, ok);
}

unsigned assemble_collision_mesh(State &s, const PpfParamSet &prm) {
    if (!s.cm_nv || prm.disable_contact) return 0;
    cudaStream_t q = s.stream;
    unsigned *cnt = s.counters.p + 34; // [34] count, [35] traversal error
    CUDA_OK(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned), q));
    Ctx c = make_ctx(s, prm);
    const CmArgs A = make_cm_args(s, cnt);
    if (s.cm_nf && s.nsv) k_cm_vertex<<<div_up(s.nsv, 128), 128, 0, q>>>(c, A, s.nsv);
    if (s.cm_ne && s.ne) k_cm_edge<<<div_up(s.ne, 128), 128, 0, q>>>(c, A, s.ne);
    if (s.nf) k_cm_face<<<div_up(s.nf, 128), 128, 0, q>>>(c, A, s.nf);
    LAUNCH_COUNT(3);
    unsigned h[2];
    CUDA_OK(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    if (h[1]) fatal(4, "BVH traversal stack overflow against the static collision mesh");
    return h[0];
}

// swept box of a dynamic element over x0 -> x0 + max_t (x1 - x0), grown by ext
__device__ __forceinline__ void cm_sweep(Box &q, const V3 &a0, const V3 &a1, float max_t, float ext, bool first) {
    const V3 z = max_t * (a1 - a0) + a0;
    const float lx = fminf(a0.x, z.x) - ext, ly = fminf(a0.y, z.y) - ext, lz = fminf(a0.z, z.z) - ext;
    const float hx = fmaxf(a0.x, z.x) + ext, hy = fmaxf(a0.y, z.y) + ext, hz = fmaxf(a0.z, z.z) + ext;
    if (first) { q.lx = lx; q.ly = ly; q.lz = lz; q.hx = hx; q.hy = hy; q.hz = hz; }
    else { q.lx = fminf(q.lx, lx); q.ly = fminf(q.ly, ly); q.lz = fminf(q.lz, lz); q.hx = fmaxf(q.hx, hx); q.hy = fmaxf(q.hy, hy); q.hz = fmaxf(q.hz, hz); }
}
__global__ void __launch_bounds__(128) k_cm_ccd_vertex(CcdCtx c, CmArgs A, unsigned nsv) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsv || A.fix_index[i] != 0 || !cw_vertex(A.active, i)) return;
    const float4 cp = c.vcp[i];
    if (cp.w == 0.0f) return;
    const V3 p0 = ld3c(c.x0, i), p1 = ld3c(c.x1, i);
    Box q;
    cm_sweep(q, p0, p1, c.max_t, 0.5f * cp.x + cp.y, true);
    const bool ok = bvh_query(A.sface, q, [&](unsigned fi) {
        const uint4 f = A.sfaces[fi];
        const float4 sp = A.sfcp[fi];
        const V3 t0 = ld3c(A.sx, f.x), t1 = ld3c(A.sx, f.y), t2 = ld3c(A.sx, f.z);
        CcdOverlap ov{0u, -1.0f, -1.0f};
        const float result = point_triangle_ccd(p0, p1, t0, t1, t2, t0, t1, t2, cp.y + sp.y, 0.5f * (cp.x + sp.x), c.max_t, c.ccd_eps, ov);
        ccd_commit(c, result, ov, i, f.x, 3u);
    });
    if (!ok) atomicExch(A.counter + 1, 1u);
}
__global__ void __launch_bounds__(128) k_cm_ccd_edge(CcdCtx c, CmArgs A, unsigned ne) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne || c.efixed[i]) return;
    const float4 cp = c.ecp[i];
    const uint2 ev = c.edge[i];
    if (!(cp.w > 0.0f) || !cw_edge(A.active, ev)) return;
    const V3 a0 = ld3c(c.x0, ev.x), b0 = ld3c(c.x0, ev.y), a1 = ld3c(c.x1, ev.x), b1 = ld3c(c.x1, ev.y);
    const float ext = 0.5f * cp.x + cp.y;
    Box q;
    cm_sweep(q, a0, a1, c.max_t, ext, true);
    cm_sweep(q, b0, b1, c.max_t, ext, false);
    const bool ok = bvh_query(A.sedge, q, [&](unsigned ei) {
        const uint2 se = A.sedges[ei];
        const float4 sp = A.secp[ei];
        const V3 q0 = ld3c(A.sx, se.x), q1 = ld3c(A.sx, se.y);
        CcdOverlap ov{0u, -1.0f, -1.0f};
        const float result = edge_edge_ccd(a0, b0, q0, q1, a1, b1, q0, q1, cp.y + sp.y, 0.5f * (cp.x + sp.x), c.max_t, c.ccd_eps, ov);
        ccd_commit(c, result, ov, ev.x, se.x, 5u);
    });
    if (!ok) atomicExch(A.counter + 1, 1u);
}
__global__ void __launch_bounds__(128) k_cm_ccd_face(CcdCtx c, CmArgs A, unsigned nf) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || c.ffixed[i]) return;
    const float4 cp = c.fcp[i];
    const uint4 fv = c.face[i];
    if (!(cp.w > 0.0f) || !cw_face(A.active, fv)) return;
    const V3 t00 = ld3c(c.x0, fv.x), t01 = ld3c(c.x0, fv.y), t02 = ld3c(c.x0, fv.z), t10 = ld3c(c.x1, fv.x), t11 = ld3c(c.x1, fv.y), t12 = ld3c(c.x1, fv.z);
    const float ext = 0.5f * cp.x + cp.y;
    Box q;
    cm_sweep(q, t00, t10, c.max_t, ext, true);
    cm_sweep(q, t01, t11, c.max_t, ext, false);
    cm_sweep(q, t02, t12, c.max_t, ext, false);
    const bool ok = bvh_query(A.svert, q, [&](unsigned vi) {
        const float4 sp = A.svcp[vi];
        const V3 p = ld3c(A.sx, vi);
        CcdOverlap ov{0u, -1.0f, -1.0f};
        const float result = point_triangle_ccd(p, p, t00, t01, t02, t10, t11, t12, cp.y + sp.y, 0.5f * (cp.x + sp.x), c.max_t, c.ccd_eps, ov);
        ccd_commit(c, result, ov, fv.x, vi, 4u);
    });
    if (!ok) atomicExch(A.counter + 1, 1u);
}

// fraction of the step x0 -> x1 that stays clear of the static collision mesh (1 = the whole step)
float collision_mesh_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm, bool &overlap) {
    overlap = false;
    if (!s.cm_nv || prm.disable_contact) return 1.0f;
    cudaStream_t q = s.stream;
    CcdCtx c;
    c.x0 = x0; c.x1 = x1; c.vcp = s.vcp.p; c.ecp = s.ecp.p; c.fcp = s.fcp.p; c.face = s.face.p; c.edge = s.edge.p;
    c.vfixed = s.vfixed.p; c.efixed = s.efixed.p; c.ffixed = s.ffixed.p; c.vcollider = s.vcollider.p;
    c.max_t = prm.line_search_max_t; c.ccd_eps = prm.ccd_eps;
    unsigned *w = s.counters.p + 48; // [48] toi bits, [49..52] overlap info, [53] unused, [54] traversal error
    const float max_t = prm.line_search_max_t;
    unsigned init[7];
    memcpy(&init[0], &max_t, sizeof(float));
    init[1] = 0u; init[2] = 0xFFFFFFFFu; init[3] = 0xFFFFFFFFu; init[4] = 0xFFFFFFFFu; init[5] = 0u; init[6] = 0u;
    CUDA_OK(cudaMemcpyAsync(w, init, sizeof(init), cudaMemcpyHostToDevice, q));
    c.toi_bits = w; c.overlap = w + 1;
    const CmArgs A = make_cm_args(s, w + 5);
    if (s.cm_nf && s.nsv) k_cm_ccd_vertex<<<div_up(s.nsv, 128), 128, 0, q>>>(c, A, s.nsv);
    if (s.cm_ne && s.ne) k_cm_ccd_edge<<<div_up(s.ne, 128), 128, 0, q>>>(c, A, s.ne);
    if (s.nf) k_cm_ccd_face<<<div_up(s.nf, 128), 128, 0, q>>>(c, A, s.nf);
    LAUNCH_COUNT(3);
    unsigned h[7];
    CUDA_OK(cudaMemcpyAsync(h, w, sizeof(h), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    if (h[6]) fatal(4, "BVH traversal stack overflow against the static collision mesh");
    float toi;
    memcpy(&toi, &h[0], sizeof(float));
    overlap = h[1] != 0u;
    return toi / prm.line_search_max_t;
}

// ---------------------------------------------------------------------------------------------
// intersection check (contact.cu:3056-3175, intersect_core.hpp:21-79)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool point_triangle_inside(const float *p, const float *d1, const float *d2) {
    const float a00 = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
    const float a01 = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2];
    const float a11 = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
    const float b0 = d1[0] * p[0] + d1[1] * p[1] + d1[2] * p[2];
    const float b1 = d2[0] * p[0] + d2[1] * p[1] + d2[2] * p[2];
    const float det = a00 * a11 - a01 * a01;
    if (det == 0.0f) return false;
    const float w0 = (a11 * b0 - a01 * b1) / det;
    const float w1 = (a00 * b1 - a01 * b0) / det;
    const float w2 = 1.0f - w0 - w1;
    float wmin = w0 < w1 ? w0 : w1;
    wmin = wmin < w2 ? wmin : w2;
    float wmax = w0 > w1 ? w0 : w1;
    wmax = wmax > w2 ? wmax : w2;
    return wmin >= 0.0f && wmax <= 1.0f;
}
__device__ __forceinline__ bool edge_triangle_intersect(const V3 &a0, const V3 &a1, const V3 &b0, const V3 &b1, const V3 &b2) {
    // wrapper of contact.cu:2745-2759: everything relative to b0, then the shared predicate with v0 = 0
    const V3 d1v = b1 - b0, d2v = b2 - b0, e0v = a0 - b0, e1v = a1 - b0;
    const float d1[3] = {d1v.x - 0.0f, d1v.y - 0.0f, d1v.z - 0.0f}, d2[3] = {d2v.x - 0.0f, d2v.y - 0.0f, d2v.z - 0.0f};
    const float A0[3] = {e0v.x - 0.0f, e0v.y - 0.0f, e0v.z - 0.0f}, A1[3] = {e1v.x - 0.0f, e1v.y - 0.0f, e1v.z - 0.0f};
    const float nx = d1[1] * d2[2] - d1[2] * d2[1], ny = d1[2] * d2[0] - d1[0] * d2[2], nz = d1[0] * d2[1] - d1[1] * d2[0];
    const float s1 = A0[0] * nx + A0[1] * ny + A0[2] * nz;
    const float s2 = A1[0] * nx + A1[1] * ny + A1[2] * nz;
    if (s1 * s2 < 0.0f) {
        const float t = s1 / (s1 - s2);
        const float r[3] = {(1.0f - t) * A0[0] + t * A1[0], (1.0f - t) * A0[1] + t * A1[1], (1.0f - t) * A0[2] + t * A1[2]};
        return point_triangle_inside(r, d1, d2);
    }
    return false;
}

struct IsectArgs {
    BvhView face, edge, vert;
    const float4 *x;
    const uint4 *faces;
    const uint2 *edges;
    const float4 *vcp, *ecp, *fcp;
    const unsigned char *vfixed, *efixed, *ffixed, *vcollider;
    PpfIntersectionRecord *records;
    unsigned *counter; // [0] record counter, [1] error
    int check_offsets;
    const unsigned char *active; // collision windows (see BpArgs)
    BvhView sface;               // static collision mesh (n = 0 when the scene has none)
    const float4 *sx;
    const uint4 *sfaces;
};

__device__ void record_isect(const IsectArgs &A, unsigned type, unsigned e0, unsigned e1, const V3 *v0, unsigned n0, const V3 *v1, unsigned n1) {
    const unsigned slot = atomicAdd(A.counter, 1u);
    if (slot < PPF_MAX_INTERSECTION_RECORDS) {
        PpfIntersectionRecord &r = A.records[slot];
        r.type = type; r.elem0 = e0; r.elem1 = e1; r.num_verts0 = n0; r.num_verts1 = n1;
        unsigned k = 0;
        for (unsigned i = 0; i < n0; ++i) { r.positions[k++] = v0[i].x; r.positions[k++] = v0[i].y; r.positions[k++] = v0[i].z; }
        for (unsigned i = 0; i < n1; ++i) { r.positions[k++] = v1[i].x; r.positions[k++] = v1[i].y; r.positions[k++] = v1[i].z; }
    }
}

__global__ void __launch_bounds__(128) k_isect_edge(IsectArgs A, unsigned nq) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const unsigned e = A.edge.prim[t];
    const uint2 ev = A.edges[e];
    if (!cw_edge(A.active, ev)) return;
    const V3 y0 = ld3c(A.x, ev.x), y1 = ld3c(A.x, ev.y);
    Box q;
    q.lx = fminf(y0.x, y1.x); q.ly = fminf(y0.y, y1.y); q.lz = fminf(y0.z, y1.z);
    q.hx = fmaxf(y0.x, y1.x); q.hy = fmaxf(y0.y, y1.y); q.hz = fmaxf(y0.z, y1.z);
    const float emass = A.ecp[e].w;
    bool ok = bvh_query(A.face, q, [&](unsigned fi) {
        const bool either_dyn = !A.ffixed[fi] || !A.efixed[e];
        const bool either_nonzero = A.fcp[fi].w > 0.0f || emass > 0.0f;
        const uint4 f = A.faces[fi];
        const bool both_collider = A.vcollider[ev.x] && A.vcollider[f.x];
        if (!either_dyn || !either_nonzero || both_collider || !cw_face(A.active, f)) return;
        if (ev.x == f.x || ev.x == f.y || ev.x == f.z || ev.y == f.x || ev.y == f.y || ev.y == f.z) return;
        const V3 x0 = ld3c(A.x, f.x), x1 = ld3c(A.x, f.y), x2 = ld3c(A.x, f.z);
        if (edge_triangle_intersect(y0, y1, x0, x1, x2)) {
            const V3 fv[3] = {x0, x1, x2}, evv[2] = {y0, y1};
            record_isect(A, 0, fi, e, fv, 3, evv, 2);
        }
    });
    if (A.sface.n && emass > 0.0f) // zero-mass edges are static solids themselves (contact.cu:3114-3126)
        ok &= bvh_query(A.sface, q, [&](unsigned fi) {
            const uint4 f = A.sfaces[fi];
            const V3 x0 = ld3c(A.sx, f.x), x1 = ld3c(A.sx, f.y), x2 = ld3c(A.sx, f.z);
            if (edge_triangle_intersect(y0, y1, x0, x1, x2)) {
                const V3 fv[3] = {x0, x1, x2}, evv[2] = {y0, y1};
                record_isect(A, 2, fi, e, fv, 3, evv, 2);
            }
        });
    if (A.check_offsets) {
        const float4 ep = A.ecp[e];
        ok &= bvh_query(A.edge, q, [&](unsigned e2) {
            if (!(e2 < e)) return;
            const uint2 fv = A.edges[e2];
            const bool either_dyn = !A.efixed[e] || !A.efixed[e2];
            const bool either_nonzero = emass > 0.0f || A.ecp[e2].w > 0.0f;
            const bool both_collider = A.vcollider[ev.x] && A.vcollider[fv.x];
            if (!either_dyn || !either_nonzero || both_collider || !cw_edge(A.active, fv)) return;
            const float offset = ep.y + A.ecp[e2].y;
            const bool shared = fv.x == ev.x || fv.x == ev.y || fv.y == ev.x || fv.y == ev.y;
            bool hit = false;
            const V3 q0 = ld3c(A.x, fv.x), q1 = ld3c(A.x, fv.y);
            if (!shared) {
                const C4 c = edge_edge_coeff(y0, y1, q0, q1);
                const V3 d = lin2(c.a, y0, c.b, y1) - lin2(c.c, q0, c.d, q1);
                hit = dot(d, d) < offset * offset;
            } else {
                const unsigned a[2] = {ev.x, ev.y}, b[2] = {fv.x, fv.y};
                for (int i = 0; i < 2 && !hit; ++i)
                    for (int j = 0; j < 2 && !hit; ++j)
                        if (a[i] == b[j]) {
                            const V3 p0 = ld3c(A.x, a[i]), p1 = ld3c(A.x, a[1 - i]), p2 = ld3c(A.x, b[1 - j]);
                            const C2 c0 = point_edge_coeff_unclassified(p1, p0, p2), c1 = point_edge_coeff_unclassified(p2, p0, p1);
                            const V3 d0 = lin2(c0.a, p0, c0.b, p2) - p1, d1 = lin2(c1.a, p0, c1.b, p1) - p2;
                            hit = fminf(dot(d0, d0), dot(d1, d1)) < offset * offset;
                        }
            }
            if (hit) {
                const V3 a0[2] = {y0, y1}, a1[2] = {q0, q1};
                record_isect(A, 1, e, e2, a0, 2, a1, 2);
            }
        });
    }
    if (!ok) atomicExch(A.counter + 1, 1u);
}

__global__ void __launch_bounds__(128) k_isect_vertex(IsectArgs A, unsigned nq) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const unsigned v = A.vert.prim[t];
    if (!cw_vertex(A.active, v)) return;
    const float4 vp = A.vcp[v];
    const V3 p = ld3c(A.x, v);
    Box q;
    q.lx = p.x - vp.y; q.ly = p.y - vp.y; q.lz = p.z - vp.y; q.hx = p.x + vp.y; q.hy = p.y + vp.y; q.hz = p.z + vp.y;
    const bool ok = bvh_query(A.vert, q, [&](unsigned w) {
        if (!(w < v)) return;
        const float4 wp = A.vcp[w];
        const bool either_dyn = !A.vfixed[v] || !A.vfixed[w];
        const bool either_nonzero = vp.w > 0.0f || wp.w > 0.0f;
        const bool both_collider = A.vcollider[v] && A.vcollider[w];
        if (!either_dyn || !either_nonzero || both_collider || !cw_vertex(A.active, w)) return;
        const float offset = vp.y + wp.y;
        const V3 qv = ld3c(A.x, w);
        const V3 e = p - qv;
        if (dot(e, e) < offset * offset) record_isect(A, 3, v, w, &p, 1, &qv, 1);
    });
    if (!ok) atomicExch(A.counter + 1, 1u);
}

bool check_intersection(State &s, const float4 *x, const PpfParamSet &prm) {
    (void)prm;
    cudaStream_t q = s.stream;
    IsectArgs A;
    A.face = view_of(s.bvh_face); A.edge = view_of(s.bvh_edge); A.vert = view_of(s.bvh_vert);
    A.x = x; A.faces = s.face.p; A.edges = s.edge.p; A.vcp = s.vcp.p; A.ecp = s.ecp.p; A.fcp = s.fcp.p;
    A.vfixed = s.vfixed.p; A.efixed = s.efixed.p; A.ffixed = s.ffixed.p; A.vcollider = s.vcollider.p;
    A.records = s.isect_records.p;
    A.counter = s.counters.p + 26;
    A.active = s.has_cw ? s.cw_active.p : nullptr;
    A.sface = view_of(s.cm_bvh_face); A.sx = s.cm_x.p; A.sfaces = s.cm_face.p;
    if (!s.cm_nf) A.sface.n = 0;
    A.check_offsets = s.has_offsets ? 1 : 0; // no contact offset anywhere: the "closer than the offsets" tests cannot fire
    CUDA_OK(cudaMemsetAsync(A.counter, 0, 2 * sizeof(unsigned), q));
    if (s.ne) k_isect_edge<<<div_up(s.ne, 128), 128, 0, q>>>(A, s.ne);
    if (s.has_offsets && s.nsv) k_isect_vertex<<<div_up(s.nsv, 128), 128, 0, q>>>(A, s.nsv);
    LAUNCH_COUNT(2);
    unsigned h[2];
    CUDA_OK(cudaMemcpyAsync(h, A.counter, sizeof(h), cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    if (h[1]) fatal(4, "BVH traversal stack overflow in the intersection check");
    s.h_isect_count = std::min<unsigned>(h[0], PPF_MAX_INTERSECTION_RECORDS);
    if (s.h_isect_count) CUDA_OK(cudaMemcpy(s.h_isect.data(), s.isect_records.p, sizeof(PpfIntersectionRecord) * s.h_isect_count, cudaMemcpyDeviceToHost));
    return h[0] == 0;
}

} // namespace ppf

// contact detection alone at caller-provided positions (3 floats per vertex): uploads them as the
// evaluation positions, refits the three BVHs and runs the broad + narrow phase.  Extension used by
// the parity tests (pair sets against the CPU oracle); returns the total number of accepted pairs.
__global__ void k_upload_positions(unsigned n, const float *__restrict__ x, float4 *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.0f);
}
extern "C" __attribute__((visibility("default"))) unsigned ppf_b200_detect_pairs(const float *x, unsigned n_vert) {
    using namespace ppf;
    State &s = g;
    if (!x || n_vert != s.nv) return 0xFFFFFFFFu;
    float *tmp = nullptr;
    CUDA_OK(cudaMalloc(&tmp, sizeof(float) * 3 * size_t(n_vert)));
    CUDA_OK(cudaMemcpyAsync(tmp, x, sizeof(float) * 3 * size_t(n_vert), cudaMemcpyHostToDevice, s.stream));
    k_upload_positions<<<div_up(n_vert, 256), 256, 0, s.stream>>>(n_vert, tmp, s.eval_x.p);
    bvh_refit(s, s.bvh_face, 0, s.eval_x.p, s.eval_x.p, 1.0f);
    bvh_refit(s, s.bvh_edge, 1, s.eval_x.p, s.eval_x.p, 1.0f);
    bvh_refit(s, s.bvh_vert, 2, s.eval_x.p, s.eval_x.p, 1.0f);
    unsigned hc[4];
    s.cand_cover_eval = false; // arbitrary positions: traverse
    const unsigned total = detect_contacts(s, *s.param, hc);
    CUDA_OK(cudaFree(tmp));
    return total;
}

// sorted contact-pair lists of the last assembly (extension used by the parity tests)
extern "C" __attribute__((visibility("default"))) unsigned ppf_b200_fetch_pairs(unsigned kind, unsigned *out, unsigned cap) {
    using namespace ppf;
    State &s = g;
    if (kind >= 4) return 0;
    const unsigned n = s.pairs[kind].count;
    if (!out || !cap || !n) return n;
    std::vector<uint2> h(n);
    CUDA_OK(cudaMemcpy(h.data(), s.pairs[kind].pairs.p, sizeof(uint2) * n, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end(), [](const uint2 &a, const uint2 &b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    for (unsigned i = 0; i < std::min(n, cap); ++i) { out[2 * i] = h[i].x; out[2 * i + 1] = h[i].y; }
    return n;
}

// backend.hpp -- device-side state of the B200 backend and the interfaces between its phases.
//
// One process owns one dataset (like the reference, main/main.cu:40-44,117-124).  Everything
// the per-step path touches lives in HBM in layouts chosen for the access pattern, converted
// once at initialize():
//   * positions are float4 (x,y,z,0): every random gather is one aligned 16-byte sector-local
//     load instead of a 12-byte load straddling sectors;
//   * faces are padded to uint4, edges uint2, hinges uint4;
//   * per-primitive contact parameters are pre-resolved into float4 (ghat, offset, friction, mass)
//     so narrow-phase kernels do one gather instead of prop -> param_index -> param chains;
//   * the topology-fixed matrix keeps the reference's upper-triangle block-CSR (so the slot
//     tables of the DataSet address it directly); the PCG works on a packed, permuted copy
//     (pcg.cu) built once per Newton iteration.
#pragma once
#include <algorithm>
#include "../../include/ppf_abi.h"
#include "vecmath.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace ppf {

// ---- fatal / error handling (reference: main/fatal.hpp:48-89, main/cuda_utils.hpp:54-109) ----
[[noreturn]] void fatal(unsigned char code, const char *fmt, ...);
void log_info(const char *fmt, ...);
#define CUDA_OK(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) ::ppf::fatal(_e == cudaErrorMemoryAllocation ? 2 : 3, "CUDA error %s at %s:%d (%s)", \
                                            cudaGetErrorString(_e), __FILE__, __LINE__, #expr);        \
    } while (0)

extern unsigned long long g_launches; // kernels launched by this backend (PpfB200Stats.n_kernel_launches)
extern unsigned long long g_dev_alloc, g_dev_free; // cudaMalloc / cudaFree calls of the grow-only buffers ("device-alloc" / "device-free" channels)
#define LAUNCH_COUNT(n) (::ppf::g_launches += (n))

inline unsigned div_up(size_t a, unsigned b) { return (unsigned)((a + b - 1) / b); }

// Grow-only device buffer.  Buffers that follow the contact set (candidates, pairs, keys, the dynamic matrix, the packed block
// list) keep growing while contacts build up, and plain cudaFree + cudaMalloc of a few hundred MB stalls for milliseconds --
// once for half a second in the twist scene -- with a device-wide synchronisation each time.  Growth therefore goes through the
// stream-ordered allocator on the backend stream (cudaMallocAsync from the default pool with an unlimited release threshold:
// freed blocks are reused, nothing goes back to the driver, no implicit synchronisation) and at least doubles the capacity.
// `plain` buffers (exported through CUDA IPC to the peer ranks) stay on cudaMalloc.
extern cudaStream_t g_alloc_stream; // the backend stream once it exists (abi.cu); null: cudaMalloc
template <class T> struct DBuf {
    T *p = nullptr;
    size_t cap = 0;
    bool plain = false;      // force cudaMalloc (IPC-exported memory)
    bool from_pool = false;  // p came from cudaMallocAsync
    static T *alloc(size_t count, bool plain_, bool &pool) {
        T *q = nullptr;
        pool = !plain_ && g_alloc_stream != nullptr;
        if (pool) CUDA_OK(cudaMallocAsync(&q, count * sizeof(T), g_alloc_stream));
        else CUDA_OK(cudaMalloc(&q, count * sizeof(T)));
        ++g_dev_alloc;
        return q;
    }
    void release() {
        if (!p) return;
        if (from_pool && g_alloc_stream) CUDA_OK(cudaFreeAsync(p, g_alloc_stream));
        else CUDA_OK(cudaFree(p));
        ++g_dev_free;
        p = nullptr;
    }
    void reserve(size_t n) { // contents are lost when it grows
        if (n > cap) {
            release();
            const size_t want = std::max(n + n / 4 + 16, cap ? 2 * cap : size_t(0));
            p = alloc(want, plain, from_pool);
            cap = want;
        }
    }
    void reserve_keep(size_t n, cudaStream_t s) { // grow preserving contents
        if (n > cap) {
            const size_t want = std::max(n + n / 2 + 16, 2 * cap);
            bool pool = false;
            T *q = alloc(want, plain, pool);
            if (p) {
                CUDA_OK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
                if (!(from_pool && g_alloc_stream == s)) CUDA_OK(cudaStreamSynchronize(s)); // a stream-ordered free on s needs no wait
                release();
            }
            p = q;
            from_pool = pool;
            cap = want;
        }
    }
    void upload(const T *h, size_t n, cudaStream_t s) {
        reserve(n);
        if (n) CUDA_OK(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void zero(size_t n, cudaStream_t s) { if (n) CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void free() { release(); cap = 0; }
};

struct Bvh {
    unsigned n = 0;             // primitives (leaves); nodes: [0,n-1) internal, [n-1, 2n-1) leaves
    DBuf<unsigned> morton, morton_sorted, prim, prim_sorted; // prim_sorted[k] = primitive of sorted leaf k
    DBuf<int2> child;           // per internal node: (left, right) node ids
    DBuf<int> parent;           // per node
    DBuf<float4> lo, hi;        // per node AABB (xyz)
    DBuf<float4> ulo, uhi;      // per node interval of sweep displacements t_max*(x1-x0) (relative-motion sweeps only)
    DBuf<float4> pack;          // per internal node, 128 bytes: its two children's (lo, hi, ulo, uhi); lo.w carries the child's node id
    DBuf<unsigned> flag;        // bottom-up arrival counters
    int root = 0;
};

struct PairList { // candidate or accepted primitive pairs
    DBuf<uint2> pairs;
    unsigned *d_count = nullptr; // device counter
    unsigned count = 0;          // host copy after a sync
};

struct Timers {
    double ms_advance = 0, ms_lbvh_build = 0, ms_assembly = 0, ms_contact = 0, ms_linsolve = 0, ms_spmv = 0,
           ms_aabb_update = 0, ms_line_search = 0;
    unsigned long long n_advance = 0, n_newton = 0, n_cg_iter = 0, n_spmv = 0;
};

struct State {
    bool initialized = false;
    cudaStream_t stream = nullptr;
    PpfDataSet host;       // by-value copy of the caller's DataSet (host pointers inside)
    PpfParamSet *param = nullptr;
    unsigned nv = 0, nsv = 0, nf = 0, ne = 0, nh = 0, nfix = 0, npull = 0;
    unsigned fixed_nnz = 0;

    // geometry / topology
    DBuf<float4> curr, prev, eval_x, target, x_old, velocity;
    DBuf<uint4> face, hinge;
    DBuf<uint2> edge;
    DBuf<unsigned> vf_off, vf_idx, ve_off, ve_idx, ef_off, ef_idx; // vertex->faces, vertex->edges, edge->faces
    // per-primitive resolved parameters
    DBuf<float4> vcp, ecp, fcp;                  // (ghat, offset, friction, mass)
    DBuf<float> varea;                            // vertex area
    DBuf<unsigned> fix_index, pull_index;         // per vertex (1-based, 0 = none)
    DBuf<unsigned char> vfixed, efixed, ffixed, hfixed, vcollider, fskip;
    DBuf<float> fpressure;                        // per face inflation pressure (FaceParam.pressure); has_pressure when any is set
    bool has_pressure = false;
    DBuf<float4> fmat;                            // per face (mu, lambda, mass, model as a float: the ABI's Model enum)
    bool has_iso_faces = false;                   // some face uses an SVD-based model (ARAP / StVK / SNHk)
    // Rayleigh damping coefficients (seconds), 0 = none: faces / rods use deform_damping, hinges / rod bending bend_damping
    DBuf<float2> fplast, hplast, rb_plast;        // plastic creep (rate per second, dead-zone threshold) per face / hinge / bending vertex
    bool has_fplast = false, has_hplast = false, has_rb_plast = false;
    DBuf<float> fdamp, hdamp, rod_damp, rb_damp;
    bool has_fdamp = false, has_hdamp = false, has_rod_damp = false, has_rb_damp = false;
    DBuf<float2> fstrain;                         // per face (strain-limit parameter, shrink-adjusted limit); 0 = off
    DBuf<float> fshrink;                          // per face min(shrink_x, shrink_y) (max_sigma diagnostic)
    bool has_strainlimit = false;
    bool has_offsets = false;                     // some vertex / edge / face parameter set carries a contact offset
    DBuf<float4> inv_rest;                        // per face 2x2
    DBuf<float4> hmat;                            // per hinge (length, area, rest_angle, bend_dir)
    DBuf<float> hinge_stiff;                      // per hinge stiffness (energy.cu:510-517), 0 = skip
    // rods (yarn): edges [0, nrod) are rod edges; bending vertices are the interior rod vertices (two edges, no face)
    unsigned nrod = 0, nrodbend = 0;
    DBuf<float2> rod_par;                         // per rod edge (rest length, stiffness * mass); weight 0 = skip
    DBuf<unsigned> rod_slots;                     // 4 per rod edge (DataSet.edge_hess_slots)
    DBuf<float2> rod_sl;                          // per rod edge (strain limit, drawn length); limit 0 = off
    bool has_rod_strainlimit = false;
    DBuf<uint4> rb_elem;                          // per bending vertex i: (j, i, k, -), its stencil in scatter order
    DBuf<float2> rb_par;                          // per bending vertex (bending coefficient, rest angle)
    DBuf<unsigned> rb_slots;                      // 9 per bending vertex (DataSet.rod_bend_hess_slots rows of those vertices)
    DBuf<PpfFixPair> fix;
    DBuf<PpfPullPair> pull;
    // torque pins (Constraint.torque_groups / torque_vertices, re-sent with every update_constraint)
    DBuf<PpfTorqueGroup> tgroups;
    DBuf<PpfTorqueVertex> tverts;
    DBuf<float> tresult; // 7 floats per group: centre, axis, 1 / sum |r_perp|^2
    unsigned ntgroup = 0, ntvert = 0;
    DBuf<PpfStitch> stitches;                     // seams of the last update_constraint
    DBuf<unsigned> stitch_slots;                  // 36 per seam (DataSet.stitch_hess_slots)
    unsigned nstitch = 0, nstitch_slots = 0;
    DBuf<PpfSphere> spheres;                      // analytic colliders of the last update_constraint
    DBuf<PpfFloor> floors;
    unsigned nsphere = 0, nfloor = 0;
    // ABI-layout copies kept for fetch_* (plasticity is out of scope, so they never change)
    // fixed (topology) matrix: upper-triangle block CSR, AoS 3x3 column-major blocks
    DBuf<unsigned> fx_off, fx_col;
    DBuf<float> fx_val, fx_el;                    // 9 floats per block: assembled / elastic-only snapshot
    DBuf<unsigned> face_slots, hinge_slots;
    DBuf<unsigned> fx_diag_slot;                  // slot of (i,i) per row
    DBuf<float> diag;                             // C: 9 floats per vertex
    DBuf<float> force, dx;                        // 3 floats per vertex (ABI order)
    // dynamic (contact) matrix of the last assembly: CSR over upper-triangle pairs not in the fixed pattern
    DBuf<unsigned long long> dyn_keys, dyn_keys_sorted;
    DBuf<unsigned> dyn_off, dyn_col;
    DBuf<float> dyn_val;
    DBuf<unsigned char> dyn_keep;                 // per block: non-zero when freshly assembled, i.e. before the Dirichlet elimination (fetch_dyn)
    unsigned dyn_nnz = 0;
    unsigned long long dyn_cap = 0;
    // carried pattern handed through fetch_dyn / update_dyn
    std::vector<unsigned> carried_index, carried_offset;
    bool dyn_valid = false;

    Bvh bvh_face, bvh_edge, bvh_vert;
    PairList cand[4];   // candidates: 0 vertex-face, 1 vertex-edge, 2 vertex-vertex, 3 edge-edge
    bool bvh_relative = false;    // the trees carry displacement intervals (last refit was bvh_refit_relative)
    bool bvh_cover_eval = false;  // the BVH boxes bound the primitives at the current eval_x (false after a relative refit)
    bool cand_cover_eval = false; // the candidate lists of the last line-search sweep cover the current eval_x (see contact.cu)
    PairList pairs[4];  // accepted contacts of the last assembly (pf, pe, pp, ee)
    DBuf<unsigned> counters; // device scalars (pair counters etc.)
    DBuf<float> scalars;     // device float scalars (reductions)
    float *h_pinned = nullptr; // pinned host scratch

    // intersection records (fetch_intersection_records)
    DBuf<PpfIntersectionRecord> isect_records;
    // static collision mesh (Constraint.mesh at initialize): positions, elements, (ghat, offset, friction, 0) per vertex /
    // edge / face and one tree per element kind, built once
    unsigned cm_nv = 0, cm_nf = 0, cm_ne = 0;
    DBuf<float4> cm_x, cm_vcp, cm_ecp, cm_fcp;
    DBuf<uint4> cm_face;
    DBuf<uint2> cm_edge;
    Bvh cm_bvh_face, cm_bvh_edge, cm_bvh_vert;
    // collision windows (init_collision_windows / refresh_collision_active): per-vertex "takes part in contact now" flags;
    // has_cw stays false for scenes without windows and the contact kernels then see a null pointer
    DBuf<unsigned> cw_dmap, cw_counts;
    DBuf<float> cw_windows;
    DBuf<unsigned char> cw_active;
    std::vector<unsigned> h_cw_dmap, h_cw_counts;
    std::vector<float> h_cw_windows;
    bool has_cw = false, cw_uploaded = false;
    // per-object contact statistics (DataSet.statistics_*; recorded per contact assembly, copied back by fetch())
    DBuf<unsigned> stat_obj, stat_sobj, stat_count;   // object of every vertex / of every static-mesh vertex (0xFFFFFFFF = none), contacts per object
    unsigned n_stat_obj = 0;
    std::vector<PpfIntersectionRecord> h_isect;
    unsigned h_isect_count = 0;

    Timers tm;
    unsigned last_num_contact[4] = {0, 0, 0, 0};
    unsigned long long spmv_n_row = 0, spmv_n_off = 0;
    std::string log_dir;
};
extern State g;

// ---- logging (simplelog.cpp): the reference's SimpleLog side channel (SimpleLog.cpp:86-178) ----
struct ScopeLog {
    explicit ScopeLog(const char *name);
    ~ScopeLog();
    void push(const char *name);
    void pop();
    void mark(const char *name, double value);
    void message(const char *fmt, ...);
    struct Impl;
    Impl *impl;
};
void simplelog_set_path(const char *dir);
void simplelog_set_time(double t);

// ---- phases ----
void upload_dataset(State &s);                               // abi.cu
void apply_constraint(State &s, const PpfConstraint *c);     // abi.cu
void step_advance(State &s, PpfStepResult *result);          // step.cu

// elastic.cu
void assemble_momentum(State &s, float dt, const PpfParamSet &prm);
void assemble_elastic(State &s, float dt, const PpfParamSet &prm);
void apply_plasticity(State &s, const PpfParamSet &prm);                                      // after the step: rest shapes creep (plasticity.cu)
void assemble_stitches(State &s, const PpfParamSet &prm);                                     // seam springs, before the elastic snapshot
void assemble_damping(State &s, float dt, const PpfParamSet &prm);                             // Rayleigh damping, part of "asm elastic"
void assemble_strainlimit(State &s, const PpfParamSet &prm);                                  // after the elastic snapshot
float max_stretch_ratio(State &s);                                                            // "max_sigma" channel (main.cu:424-488)
float strain_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm); // fraction in (0, 1]
float rod_strain_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm); // same for rod edges
void recompute_fixed_flags(State &s);

// bvh.cu
void bvh_build(State &s, Bvh &b, int kind, const float4 *x);      // kind: 0 face, 1 edge, 2 vertex
void bvh_build_static(State &s, Bvh &b, int kind, unsigned n, const float4 *x, const uint4 *face, const uint2 *edge, const float4 *cp);
void bvh_refit(State &s, Bvh &b, int kind, const float4 *x0, const float4 *x1, float extrapolate);
// boxes at x0 plus per-node displacement intervals t_max*(x1-x0) for relative-motion sweep queries
void bvh_refit_relative(State &s, Bvh &b, int kind, const float4 *x0, const float4 *x1, float t_max);

// contact.cu
unsigned assemble_contact(State &s, float dt, const PpfParamSet &prm, unsigned &max_nnz_row, float &dyn_consumed);
float contact_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm, bool &overlap);
bool check_intersection(State &s, const float4 *x, const PpfParamSet &prm);
unsigned assemble_collision_mesh(State &s, const PpfParamSet &prm); // static collision mesh; returns the contact count
float collision_mesh_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm, bool &overlap);
unsigned assemble_colliders(State &s, const PpfParamSet &prm);   // analytic spheres / walls; returns the contact count
// collider part of the line search: step fraction in (0, 1]; bad_vertex != 0xFFFFFFFF: a pinned vertex is driven through a collider
float collider_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm, unsigned &bad_vertex);
void build_fetch_pattern(State &s, std::vector<unsigned> &index, std::vector<unsigned> &offset);

// pcg.cu
struct SolveResult { bool ok; unsigned iter; float resid; };
SolveResult pcg_solve(State &s, const PpfParamSet &prm, const unsigned char *dof_mask);
void apply_dirichlet(State &s);
bool pcg_schwarz_served();   // ParamSet.precond == Schwarz was honoured by the last solve (multilevel additive preconditioner, pcg.cu)

// prims.cu -- reductions on the backend stream, result returned to the host
float reduce_max_f(State &s, const float *d, size_t n, float init);
float reduce_min_f(State &s, const float *d, size_t n, float init);
unsigned reduce_sum_u(State &s, const unsigned *d, size_t n);
// prims.cu -- LSD radix sort of the low `bits` bits (stable; the first buffers are scratch afterwards, the result is in *_sorted)
void radix_sort_pairs_u32(State &s, unsigned *keys, unsigned *keys_sorted, unsigned *vals, unsigned *vals_sorted, unsigned n, int bits);
void radix_sort_keys_u64(State &s, unsigned long long *keys, unsigned long long *keys_sorted, unsigned n, int bits);
// first key of every run of a sorted list -> out (may not alias), number of runs -> *d_count (device)
void unique_sorted_u64(State &s, const unsigned long long *keys_sorted, unsigned long long *out, unsigned n, unsigned *d_count);

} // namespace ppf

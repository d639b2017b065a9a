/* ppf_abi.h -- the drop-in C ABI of the ppf-contact-solver GPU backend.
 *
 * This header is the boundary of the B200-native backend in this repository:
 * a shared library named libsimbackend_cuda.so that exports exactly the
 * extern "C" symbols the reference's Rust host binds
 * (crates/ppf-cts-solver/src/backend.rs:28-62 and main.rs:44-53), taking and
 * returning exactly the repr(C) structures the reference defines in
 * crates/ppf-cts-solver/src/cpp/data.hpp:85-748 (Rust mirror: src/data.rs).
 *
 * Everything here is plain C: pointers, sizes, 4-byte enums.  All pointers
 * inside DataSet / Constraint / RestShapeUpdate are HOST memory owned by the
 * caller; `initialize` deep-copies them to the device and keeps the by-value
 * host DataSet (with its host pointers) and the ParamSet* for the lifetime of
 * the process: `fetch*` write back into the caller's arrays and `advance`
 * updates param->prev_dt / param->time (reference: main/main.cu:140-142,
 * 1572-1573, 1926-1968).
 *
 * Layout is pinned by the static assertions at the bottom (sizes/offsets were
 * measured by compiling the reference data.hpp with g++ on x86-64).
 */
#ifndef PPF_ABI_H
#define PPF_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- small fixed-size algebra types (column-major, align 4) ------------- */
/* reference: linalg/smat.hpp:6-16 (SMat<T,R,C> is a POD array, column major) */
typedef struct { float v[3]; } PpfVec3f;
typedef struct { unsigned v[2]; } PpfVec2u;
typedef struct { unsigned v[3]; } PpfVec3u;
typedef struct { unsigned v[4]; } PpfVec4u;
typedef struct { unsigned v[6]; } PpfVec6u;
typedef struct { float v[6]; } PpfVec6f;
typedef struct { float m[4]; } PpfMat2x2f; /* m[col*2+row] */
typedef struct { float m[9]; } PpfMat3x3f; /* m[col*3+row] */

/* reference: vec/vec.hpp:79-83  Vec<T> = {T* data; u32 size; u32 allocated} */
#define PPF_VEC(T, NAME) typedef struct { T *data; unsigned size; unsigned allocated; } NAME
/* reference: vec/vec.hpp:15-22  VecVec<T> */
#define PPF_VECVEC(T, NAME)                                                    \
    typedef struct {                                                           \
        T *data;                                                               \
        unsigned *offset;                                                      \
        unsigned size, nnz, nnz_allocated, offset_allocated;                   \
    } NAME

/* ---- enums (4-byte, order fixed by data.hpp:77-83) ---------------------- */
typedef enum { PPF_MODEL_ARAP = 0, PPF_MODEL_STVK, PPF_MODEL_BARAFFWITKIN, PPF_MODEL_SNHK, PPF_MODEL_PDRD } PpfModel;
typedef enum { PPF_BARRIER_CUBIC = 0, PPF_BARRIER_QUAD, PPF_BARRIER_LOG } PpfBarrier;
typedef enum { PPF_FRICTION_MIN = 0, PPF_FRICTION_MAX, PPF_FRICTION_MEAN } PpfFrictionMode;
typedef enum { PPF_PRECOND_BLOCKJACOBI = 0, PPF_PRECOND_SCHWARZ } PpfPrecondMode;

/* ---- per-element parameter / property records (data.hpp:119-349) -------- */
typedef struct { float ghat, offset, friction; } PpfVertexParam;
typedef struct {
    float stiffness, bend, ghat, offset, friction, strainlimit, plasticity, plasticity_threshold;
    bool bend_rest_from_geometry;
    float deform_damping, bend_damping;
} PpfEdgeParam;
typedef struct {
    PpfModel model;
    float mu, lambda, friction, ghat, offset, bend, strainlimit, shrink_x, shrink_y, pressure;
    float plasticity, plasticity_threshold, bend_plasticity, bend_plasticity_threshold;
    bool bend_rest_from_geometry;
    float deform_damping, bend_damping, bend_warp, bend_weft;
} PpfFaceParam;
typedef struct { float bend, ghat, offset, plasticity, plasticity_threshold, bend_damping, bend_warp, bend_weft; } PpfHingeParam;
typedef struct { PpfModel model; float mu, lambda, shrink, plasticity, plasticity_threshold, deform_damping; } PpfTetParam;

typedef struct {
    float area, volume, mass, rest_bend_angle;
    unsigned fix_index, pull_index, param_index, pdrd_body_index;
    bool collider;
} PpfVertexProp;
typedef struct { float length, initial_length, mass; bool fixed; unsigned param_index; } PpfEdgeProp;
typedef struct { float area, mass; bool fixed, rest_excluded, collider; unsigned param_index; } PpfFaceProp;
typedef struct { float length, area, rest_angle, uv_edge_sin2; bool fixed, collider; unsigned param_index; } PpfHingeProp;
typedef struct { float mass, volume; bool fixed, rest_excluded; unsigned param_index; } PpfTetProp;
typedef struct {
    PpfVec3f rest_centroid;
    PpfMat3x3f rest_gram_inv;
    float volume;
    unsigned vertex_start, vertex_count;
    float mass_per_vertex;
    unsigned joint_mode;
    PpfVec3f joint_axis, joint_pin;
} PpfPdrdBodyProp;

/* ---- containers --------------------------------------------------------- */
PPF_VEC(PpfVec3f, PpfVecVec3f);
PPF_VEC(PpfVec2u, PpfVecVec2u);
PPF_VEC(PpfVec3u, PpfVecVec3u);
PPF_VEC(PpfVec4u, PpfVecVec4u);
PPF_VEC(char, PpfVecChar);
PPF_VEC(unsigned char, PpfVecU8);
PPF_VEC(unsigned, PpfVecU32);
PPF_VEC(float, PpfVecF32);
PPF_VEC(PpfMat2x2f, PpfVecMat2x2f);
PPF_VEC(PpfMat3x3f, PpfVecMat3x3f);
PPF_VEC(PpfVertexParam, PpfVecVertexParam);
PPF_VEC(PpfEdgeParam, PpfVecEdgeParam);
PPF_VEC(PpfFaceParam, PpfVecFaceParam);
PPF_VEC(PpfHingeParam, PpfVecHingeParam);
PPF_VEC(PpfTetParam, PpfVecTetParam);
PPF_VEC(PpfVertexProp, PpfVecVertexProp);
PPF_VEC(PpfEdgeProp, PpfVecEdgeProp);
PPF_VEC(PpfFaceProp, PpfVecFaceProp);
PPF_VEC(PpfHingeProp, PpfVecHingeProp);
PPF_VEC(PpfTetProp, PpfVecTetProp);
PPF_VEC(PpfPdrdBodyProp, PpfVecPdrdBodyProp);
PPF_VECVEC(unsigned, PpfVecVecU32);
PPF_VECVEC(PpfVec2u, PpfVecVecVec2u);

/* ---- mesh (data.hpp:85-117) --------------------------------------------- */
typedef struct { PpfVecVecU32 face, hinge, edge, rod; } PpfVertexNeighbor;
typedef struct { PpfVecVecU32 face; } PpfHingeNeighbor;
typedef struct { PpfVecVecU32 face; } PpfEdgeNeighbor;
typedef struct {
    struct { PpfVecVec3u face; PpfVecVec4u hinge; PpfVecVec2u edge; PpfVecVec4u tet; } mesh;
    struct { PpfVertexNeighbor vertex; PpfHingeNeighbor hinge; PpfEdgeNeighbor edge; } neighbor;
    struct { PpfVecChar face, vertex, hinge; } type;
} PpfMeshInfo;

typedef struct {
    PpfVecVertexProp vertex; PpfVecEdgeProp edge; PpfVecFaceProp face;
    PpfVecHingeProp hinge; PpfVecTetProp tet; PpfVecPdrdBodyProp pdrd_body;
} PpfPropSet;
typedef struct {
    PpfVecVertexParam vertex; PpfVecEdgeParam edge; PpfVecFaceParam face;
    PpfVecHingeParam hinge; PpfVecTetParam tet;
} PpfParamArrays;

/* ---- constraints (data.hpp:397-499) ------------------------------------- */
typedef struct { PpfVec3f position, step_delta; float ghat; unsigned index; bool kinematic; } PpfFixPair;
typedef struct { PpfVec3f position; float weight; unsigned index; } PpfPullPair;
typedef struct { unsigned axis_component, vertex_start, vertex_count, hint_vertex; } PpfTorqueGroup;
typedef struct { float magnitude; unsigned index, group_id; } PpfTorqueVertex;
typedef struct { PpfVec6u index; PpfVec6f weight; float stiffness; } PpfStitch;
typedef struct { PpfVec3f center; float ghat, friction, radius, thickness; bool bowl, reverse, kinematic; } PpfSphere;
typedef struct { PpfVec3f ground; float ghat, friction, thickness; PpfVec3f up; bool kinematic; } PpfFloor;
PPF_VEC(PpfFixPair, PpfVecFixPair);
PPF_VEC(PpfPullPair, PpfVecPullPair);
PPF_VEC(PpfTorqueGroup, PpfVecTorqueGroup);
PPF_VEC(PpfTorqueVertex, PpfVecTorqueVertex);
PPF_VEC(PpfStitch, PpfVecStitch);
PPF_VEC(PpfSphere, PpfVecSphere);
PPF_VEC(PpfFloor, PpfVecFloor);
typedef struct {
    PpfVecVec3f vertex; PpfVecVec3u face; PpfVecVec2u edge;
    struct { PpfVecVertexProp vertex; PpfVecFaceProp face; PpfVecEdgeProp edge; } prop;
    struct { PpfVecVertexParam vertex; PpfVecFaceParam face; PpfVecEdgeParam edge; } param_arrays;
    struct { PpfVertexNeighbor vertex; PpfHingeNeighbor hinge; PpfEdgeNeighbor edge; } neighbor;
} PpfCollisionMesh;
typedef struct {
    PpfVecFixPair fix; PpfVecPullPair pull; PpfVecTorqueGroup torque_groups; PpfVecTorqueVertex torque_vertices;
    PpfVecSphere sphere; PpfVecFloor floor; PpfVecStitch stitch; PpfCollisionMesh mesh;
} PpfConstraint;

typedef struct { PpfVecMat2x2f inv_rest2x2; PpfVecMat3x3f inv_rest3x3; PpfVecU8 exclude_face, exclude_tet; } PpfRestShapeUpdate;

/* ---- parameters / results (data.hpp:515-595) ---------------------------- */
typedef struct {
    double time;
    float air_friction, air_density, constraint_tol, prev_dt, dt, playback;
    unsigned min_newton_steps;
    float target_toi, stitch_length_factor;
    unsigned cg_max_iter;
    float cg_tol, line_search_max_t, ccd_eps, max_dx, eiganalysis_eps, friction_eps, isotropic_air_friction;
    PpfVec3f gravity, wind;
    PpfBarrier barrier;
    PpfFrictionMode friction_mode;
    unsigned csrmat_max_nnz;
    float fix_xz;
    bool disable_contact, inactive_momentum;
    PpfPrecondMode precond;
    unsigned schwarz_levels, max_newton_steps;
    bool disable_pin_dof_removal;
} PpfParamSet;

typedef struct {
    double time;
    bool ccd_success, pcg_success, intersection_free, newton_progress, pin_feasible, contact_separated;
} PpfStepResult;

typedef struct { PpfVecVec3f prev, curr; } PpfVertexSet;
typedef struct {
    PpfVec3f axis; float total_mass; unsigned pdrd_body_index, dmap_index;
    PpfVec3f rotation_axis; unsigned rotation_mode; PpfVec3f anchor;
} PpfTranslationLock;
PPF_VEC(PpfTranslationLock, PpfVecTranslationLock);

/* data.hpp:624-720 */
typedef struct {
    PpfVertexSet vertex;
    PpfMeshInfo mesh;
    PpfPropSet prop;
    PpfParamArrays param_arrays;
    PpfVecMat2x2f inv_rest2x2;
    PpfVecMat3x3f inv_rest3x3;
    PpfConstraint constraint;
    PpfVecVecU32 fixed_index_table;
    PpfVecVecVec2u transpose_table;
    unsigned rod_count, shell_face_count, surface_vert_count;
    PpfVecU32 pdrd_vert_list;
    PpfVecVec3f pdrd_rest_centered;
    PpfVecVec3f grain_omega;
    PpfVecF32 grain_inv_inertia;
    PpfVecVec3f grain_torque;
    PpfVecF32 grain_ang_stiff;
    PpfVecVec3f grain_contact_normal;
    PpfVecF32 grain_inv_inertia_center;
    PpfVecVec3f grain_omega_prev;
    PpfVecMat3x3f grain_A, grain_B;
    PpfVecVec3f grain_grot;
    PpfVecU32 tet_hess_slots, face_hess_slots, edge_hess_slots, hinge_hess_slots, rod_bend_hess_slots, stitch_hess_slots;
    PpfVecTranslationLock translation_lock;
    PpfVecU32 translation_lock_index;
    PpfVecVec3f translation_lock_initial;
    PpfVecU32 statistics_object_index, statistics_static_object_index, statistics_contact_count;
} PpfDataSet;

#define PPF_MAX_INTERSECTION_RECORDS 256 /* data.hpp:739 */
typedef struct {
    unsigned type, elem0, elem1, num_verts0, num_verts1;
    float positions[15];
} PpfIntersectionRecord;

/* fatal codes (main/fatal.hpp:48-89) */
enum {
    PPF_FATAL_NONE = 0, PPF_FATAL_INIT_INTERSECTION = 1, PPF_FATAL_OOM = 2, PPF_FATAL_CUDA_DRIVER = 3,
    PPF_FATAL_SOLVER_INVARIANT = 4, PPF_FATAL_DEVICE_ASSERT = 5, PPF_FATAL_WATCHDOG_TIMEOUT = 6
};

/* ---- exported entry points ---------------------------------------------- */
/* Each replaces the same-named symbol of the reference backend; line numbers
 * are into crates/ppf-cts-solver/src/cpp/main/main.cu.                        */
bool initialize(PpfDataSet *dataset, PpfParamSet *param);                 /* :1877 */
void advance(PpfStepResult *result);                                      /* :1922 */
void fetch(void);                                                         /* :1926 */
void fetch_inv_rest(void);                                                /* :1939 */
void fetch_rest_angles(void);                                             /* :1950 */
void fetch_dyn_counts(unsigned *n_value, unsigned *n_offset);             /* :1970 */
void fetch_dyn(unsigned *index, PpfMat3x3f *value, unsigned *offset);     /* :1979 */
void update_dyn(unsigned *index, unsigned *offset);                       /* :1984 */
void update_constraint(const PpfConstraint *constraint);                  /* :1988 */
void update_rest_shape(const PpfRestShapeUpdate *update);                 /* :2079 */
void override_velocity(const unsigned *indices, unsigned count, float vx, float vy, float vz, float dt); /* :2119 */
void gather_current_positions(const unsigned *indices, unsigned count, float *out);                      /* :2155 */
void override_angular_velocity(const unsigned *indices, unsigned count, float wx, float wy, float wz,
                               float cx, float cy, float cz, float dt);                                   /* :2193 */
void init_collision_windows(const unsigned *vert_dmap, unsigned vert_count, const float *windows,
                            const unsigned *window_counts, unsigned n_groups);                            /* :2230 */
void refresh_collision_active(float time);                                /* :2266 */
unsigned fetch_intersection_records(PpfIntersectionRecord *out, unsigned max_count); /* :2341 */
void set_log_path(const char *data_dir);                                  /* :1681 */
unsigned char ppf_fatal_code(void);                                       /* :1673 */
const char *ppf_fatal_detail(void);                                       /* :1677 */
extern unsigned char g_ppf_fatal_code;                                    /* :1658 */
extern char g_ppf_fatal_detail[512];                                      /* :1659 */
/* Imported from the host (#[no_mangle] in main.rs:47); the library carries a
 * weak no-op fallback like the reference's stub.cpp:8.                        */
void print_rust(const char *msg);

/* ---- extensions of this backend (not part of the reference ABI) ---------
 * Read-only instrumentation used by tests/ and bench.py.  A host that does
 * not know them never calls them.                                            */
typedef struct {
    /* device-side (CUDA event) milliseconds accumulated since last reset */
    double ms_advance, ms_lbvh_build, ms_assembly, ms_contact, ms_linsolve, ms_spmv, ms_aabb_update, ms_line_search;
    unsigned long long n_advance, n_newton, n_cg_iter, n_spmv, n_kernel_launches;
    unsigned long long spmv_n_row, spmv_n_off; /* of the last SpMV */
    unsigned last_num_contact[4];               /* pf, pe, pp, ee of the last assembly */
} PpfB200Stats;
void ppf_b200_get_stats(PpfB200Stats *out);
void ppf_b200_reset_stats(void);
/* Copies the four sorted contact-pair lists of the LAST contact assembly.
 * kind: 0 point-face, 1 point-edge, 2 point-point, 3 edge-edge.  Returns the
 * pair count; copies min(count, cap) pairs (a,b) into out[2*i], out[2*i+1]. */
unsigned ppf_b200_fetch_pairs(unsigned kind, unsigned *out, unsigned cap);
/* Contact detection alone (BVH refit + broad + narrow phase) at caller-provided positions, 3 floats
 * per vertex; the lists are then read with ppf_b200_fetch_pairs.  Returns the total pair count, or
 * 0xFFFFFFFF when n_vert does not match the initialized scene.  Clobbers the backend's evaluation
 * positions: call it between steps only. */
unsigned ppf_b200_detect_pairs(const float *x, unsigned n_vert);
/* Multi-GPU (one process per GPU, csrc/dist.hpp): rank 0 creates a 128-byte NCCL id, the launcher
 * distributes it, every rank calls ppf_b200_dist_init after initialize() and before the first
 * advance().  All ranks must hold the same scene; the PCG rows are then sharded over the ranks.
 * Return 0 on success (unique_id: the id length). */
int ppf_b200_dist_unique_id(unsigned char *out, unsigned cap);
int ppf_b200_dist_init(int rank, int world, const unsigned char *id_bytes, unsigned len);
/* Optional peer-memory fast path: every rank exports 128 bytes of CUDA IPC handles, the launcher
 * all-gathers them in rank order and hands the concatenation to every rank.  connect returns 1 when the
 * per-iteration collectives now run over NVLink peer memory, 0/negative when NCCL stays in use. */
int ppf_b200_dist_export(unsigned char *out, unsigned cap);
int ppf_b200_dist_connect(const unsigned char *all, unsigned len);
/* Self-test of the backend's own radix sort / unique kernels (csrc/prims.cu) on n pseudo-random keys against the host's
 * std::stable_sort / std::unique; 0 = every check passed.  Needs no initialize(). */
int ppf_b200_selftest_sort(unsigned n, int bits, unsigned seed);
/* Average microseconds of the SpMV kernel alone on the last packed system (CUDA events, `reps` launches). */
double ppf_b200_measure_spmv(unsigned reps, unsigned long long *n_row, unsigned long long *n_off);

#ifdef __cplusplus
}
#endif

/* ---- layout pins (measured from the reference data.hpp) ---- */
#ifdef __cplusplus
#define PPF_SA(c) static_assert(c, #c)
#else
#define PPF_SA(c) _Static_assert(c, #c)
#endif
PPF_SA(sizeof(PpfVec3f) == 12);
PPF_SA(sizeof(PpfMat3x3f) == 36);
PPF_SA(sizeof(PpfMat2x2f) == 16);
PPF_SA(sizeof(PpfVecU32) == 16);
PPF_SA(sizeof(PpfVecVecU32) == 32);
PPF_SA(sizeof(PpfVertexParam) == 12);
PPF_SA(sizeof(PpfEdgeParam) == 44);
PPF_SA(offsetof(PpfEdgeParam, bend_rest_from_geometry) == 32);
PPF_SA(offsetof(PpfEdgeParam, deform_damping) == 36);
PPF_SA(sizeof(PpfFaceParam) == 80);
PPF_SA(offsetof(PpfFaceParam, bend_rest_from_geometry) == 60);
PPF_SA(offsetof(PpfFaceParam, deform_damping) == 64);
PPF_SA(sizeof(PpfHingeParam) == 32);
PPF_SA(sizeof(PpfTetParam) == 28);
PPF_SA(sizeof(PpfVertexProp) == 36);
PPF_SA(offsetof(PpfVertexProp, collider) == 32);
PPF_SA(sizeof(PpfEdgeProp) == 20);
PPF_SA(offsetof(PpfEdgeProp, param_index) == 16);
PPF_SA(sizeof(PpfFaceProp) == 16);
PPF_SA(offsetof(PpfFaceProp, param_index) == 12);
PPF_SA(sizeof(PpfHingeProp) == 24);
PPF_SA(offsetof(PpfHingeProp, param_index) == 20);
PPF_SA(sizeof(PpfTetProp) == 16);
PPF_SA(sizeof(PpfPdrdBodyProp) == 92);
PPF_SA(sizeof(PpfFixPair) == 36);
PPF_SA(offsetof(PpfFixPair, kinematic) == 32);
PPF_SA(sizeof(PpfPullPair) == 20);
PPF_SA(sizeof(PpfTorqueGroup) == 16);
PPF_SA(sizeof(PpfTorqueVertex) == 12);
PPF_SA(sizeof(PpfStitch) == 52);
PPF_SA(sizeof(PpfSphere) == 32);
PPF_SA(sizeof(PpfFloor) == 40);
PPF_SA(sizeof(PpfCollisionMesh) == 336);
PPF_SA(sizeof(PpfConstraint) == 448);
PPF_SA(sizeof(PpfMeshInfo) == 304);
PPF_SA(sizeof(PpfPropSet) == 96);
PPF_SA(sizeof(PpfParamArrays) == 80);
PPF_SA(sizeof(PpfVertexSet) == 32);
PPF_SA(sizeof(PpfTranslationLock) == 52);
PPF_SA(sizeof(PpfRestShapeUpdate) == 64);
PPF_SA(sizeof(PpfIntersectionRecord) == 80);
PPF_SA(sizeof(PpfParamSet) == 136);
PPF_SA(offsetof(PpfParamSet, dt) == 24);
PPF_SA(offsetof(PpfParamSet, gravity) == 76);
PPF_SA(offsetof(PpfParamSet, barrier) == 100);
PPF_SA(offsetof(PpfParamSet, csrmat_max_nnz) == 108);
PPF_SA(offsetof(PpfParamSet, fix_xz) == 112);
PPF_SA(offsetof(PpfParamSet, disable_contact) == 116);
PPF_SA(offsetof(PpfParamSet, precond) == 120);
PPF_SA(offsetof(PpfParamSet, max_newton_steps) == 128);
PPF_SA(offsetof(PpfParamSet, disable_pin_dof_removal) == 132);
PPF_SA(sizeof(PpfStepResult) == 16);
PPF_SA(sizeof(PpfDataSet) == 1456);
PPF_SA(offsetof(PpfDataSet, mesh) == 32);
PPF_SA(offsetof(PpfDataSet, prop) == 336);
PPF_SA(offsetof(PpfDataSet, param_arrays) == 432);
PPF_SA(offsetof(PpfDataSet, inv_rest2x2) == 512);
PPF_SA(offsetof(PpfDataSet, constraint) == 544);
PPF_SA(offsetof(PpfDataSet, fixed_index_table) == 992);
PPF_SA(offsetof(PpfDataSet, transpose_table) == 1024);
PPF_SA(offsetof(PpfDataSet, rod_count) == 1056);
PPF_SA(offsetof(PpfDataSet, pdrd_vert_list) == 1072);
PPF_SA(offsetof(PpfDataSet, tet_hess_slots) == 1264);
PPF_SA(offsetof(PpfDataSet, translation_lock) == 1360);
PPF_SA(offsetof(PpfDataSet, statistics_contact_count) == 1440);

#endif /* PPF_ABI_H */

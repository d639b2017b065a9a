// ppf_host.hpp -- C++ mirror of the reference's Rust solver host for the hot path.
//
// The reference host (crates/ppf-cts-solver/src/{main,scene,mesh,builder,backend}.rs)
// is Rust; no Rust toolchain exists in this build image, so the part of it that
// stands between a scene description and the backend's C ABI is restated here in
// C++17, function by function, with the reference file:line each one follows:
//
//   mesh generators      crates/ppf-cts-core/src/datamodel/mesh.rs:181-295,606-670
//   index map            crates/ppf-cts-core/src/kernels/scene_build/index_map.rs:68-153
//   topology             crates/ppf-cts-solver/src/mesh.rs:66-293
//   props                crates/ppf-cts-solver/src/scene.rs:1538-1930, main.rs:204-258
//   dataset builder      crates/ppf-cts-solver/src/builder.rs:339-1627, 1629-1681
//   pin operations       crates/ppf-cts-solver/src/scene.rs:1944-2130, 2747-2870
//                        crates/ppf-cts-core/src/datamodel/pin_apply.rs:27-160
//   run loop             crates/ppf-cts-solver/src/backend.rs:626-1025
//
// The host never touches the GPU; it talks to a backend only through the 20
// symbols of include/ppf_abi.h resolved with dlopen/dlsym, so the same host
// drives the reference build (oracle/_ref) and the B200 backend.
#pragma once
#include "../../include/ppf_abi.h"
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ppfh {

using Vec3d = std::array<double, 3>;
using Vec3f = std::array<float, 3>;

// ---- scene description (what the frontend would write into a session dir) ----
struct PinOp {
    enum Kind { MoveBy, Spin, Scale, Torque } kind;
    // Torque (pin.rs:131-140): a force, not a motion -- the pinned vertices stay free and are driven about a principal axis
    // of the group (0 = largest spread), oriented toward the hint vertex (local index), at full power inside [t_start, t_end]
    unsigned axis_component = 0, hint_vertex = 0;
    double magnitude = 0.0;
    Vec3d delta{0, 0, 0};      // MoveBy
    Vec3d center{0, 0, 0};     // Spin / Scale
    Vec3d axis{0, 1, 0};       // Spin
    double angular_velocity = 360.0; // deg/s
    double factor = 1.0;       // Scale
    double t_start = 0.0, t_end = 1e300;
    bool smooth = false;       // transition == "smooth"
};
struct Pin {
    std::vector<unsigned> index; // local vertex indices of the object
    std::vector<PinOp> ops;
    double pull_w = 0.0;
    double unpin_time = -1.0;
};
struct ShellObject {
    std::vector<Vec3d> vert;
    std::vector<std::array<unsigned, 3>> tri;
    std::map<std::string, double> param; // tri params (crates/ppf-cts-core/src/datamodel/params.rs:295-420)
    // rod (yarn) object: edges between the object's vertices and the rod parameter set (ObjectKind::Rod defaults);
    // an object is either a shell (tri) or a rod (rod), like app.asset.add.tri / add.rod of the frontend
    std::vector<std::array<unsigned, 2>> rod;
    std::map<std::string, double> rod_param;
    std::vector<Pin> pins;
    Vec3d displacement{0, 0, 0};
    Vec3d velocity{0, 0, 0};
    // collision windows: [start, end) intervals of simulated time during which the object takes part in contact; none = always
    // (scene.rs:2471-2542 build_collision_window_table, at most 8 per object)
    std::vector<std::pair<double, double>> collision_windows;
};
struct SimArgs { // crates/ppf-cts-solver/src/args.rs:22-94 (defaults params.rs:207-290)
    bool disable_contact = false, inactive_momentum = false, include_face_mass = false;
    double dt = 1e-3, playback = 1.0, fps = 60.0;
    unsigned min_newton_steps = 0, max_newton_steps = 2048;
    double target_toi = 0.25, air_friction = 0.2, line_search_max_t = 1.25, constraint_ghat = 1e-3,
           constraint_tol = 0.01;
    unsigned cg_max_iter = 10000;
    double cg_tol = 1e-3, ccd_eps = 1e-7, max_dx = 1.0, eiganalysis_eps = 1e-2, friction_eps = 1e-5;
    unsigned csrmat_max_nnz = 10000000;
    int frames = 300;
    std::string barrier = "cubic", friction_mode = "min", precond = "block-jacobi";
    unsigned schwarz_levels = 2;
    double stitch_length_factor = 10.0, air_density = 1e-3, isotropic_air_friction = 0.0;
    Vec3d gravity{0, -9.8, 0}, wind{0, 0, 0};
    double fix_xz = 0.0, world_scaling = 1.0;
    // time-varying gravity keyframes (scene.rs dyn_param; used by `hang`): (time, gravity-y)
    std::vector<std::pair<double, double>> gravity_y_keys;
};
// invisible analytic colliders (frontend scene.add.invisible.wall / .sphere; scene.rs InvisibleWall / InvisibleSphere).
// `keys` = further keyframes (time, position[, radius]) after the one at t = 0 (frontend .move_to / .move_by / .radius): a
// collider with keyframes is kinematic for the whole run and blended linearly (or smoothstep) between them (scene.rs:2014-2049).
struct WallCollider {
    Vec3d position{0, 0, 0}, normal{0, 1, 0};
    double contact_gap = 1e-3, friction = 0.0, thickness = 1.0;
    std::vector<std::pair<double, Vec3d>> keys;
    bool smooth = false;
};
struct SphereCollider {
    Vec3d center{0, 0, 0};
    double radius = 1.0;
    bool inverted = false, hemisphere = false;
    double contact_gap = 1e-3, friction = 0.0, thickness = 1.0;
    struct Key { double time; Vec3d center; double radius; };
    std::vector<Key> keys;
    bool smooth = false;
};
// one seam spring between a barycentric point of object `src` and one of object `dst` (local vertex indices; a plain
// vertex is {v, v, v} with weights {1, 0, 0}); scene.rs stitch_ind / stitch_w / stitch_stiffness
struct StitchSpec {
    unsigned src = 0, dst = 0;
    std::array<unsigned, 3> src_index{0, 0, 0}, dst_index{0, 0, 0};
    std::array<double, 3> src_weight{1, 0, 0}, dst_weight{1, 0, 0};
    double stiffness = 1.0; // "stitch-stiffness"
};
// a static triangle mesh the simulated objects collide with (scene.rs:1945-2013 static_vert / static_tri -> Constraint.mesh)
struct StaticObject {
    std::vector<Vec3d> vert;
    std::vector<std::array<unsigned, 3>> tri;
    Vec3d displacement{0, 0, 0};
    double contact_gap = 1e-3, contact_offset = 0.0, friction = 0.0;
};
struct SceneSpec {
    std::string name;
    std::vector<ShellObject> objects;
    std::vector<StaticObject> statics;
    std::vector<StitchSpec> stitches;
    std::vector<WallCollider> walls;
    std::vector<SphereCollider> spheres;
    SimArgs args;
};

// mesh generators
void gen_cylinder(double r, double min_x, double max_x, unsigned n, std::vector<Vec3d> &V,
                  std::vector<std::array<unsigned, 3>> &F);
void gen_square(unsigned res, double size, Vec3d ex, Vec3d ey, std::vector<Vec3d> &V,
                std::vector<std::array<unsigned, 3>> &F);
void gen_icosphere(double r, unsigned subdiv_count, std::vector<Vec3d> &V, std::vector<std::array<unsigned, 3>> &F);
std::vector<unsigned> grab(const std::vector<Vec3d> &V, Vec3d dir, double eps = 1e-3);
// make_woven_cylinder of examples/woven.ipynb / large-woven.ipynb: axial strands (open) and woven rings (closed)
void gen_woven_cylinder(unsigned n, double offset, double scale, std::vector<std::vector<Vec3d>> &strands, std::vector<bool> &closed);

// named scenes of BASELINE.json; `n` overrides the resolution (0 = the config's own)
SceneSpec make_scene(const std::string &name, unsigned n);

// ---- built scene: owns every array the DataSet points into ----
template <class T> struct OwnedVecVec {
    std::vector<T> data;
    std::vector<unsigned> offset;
    void from(const std::vector<std::vector<T>> &rows);
};
struct Built {
    SceneSpec spec;
    // global (index-mapped) geometry
    std::vector<Vec3f> vert;             // initial positions (f32), global order
    std::vector<unsigned> vert_dmap;     // object id per vertex
    std::vector<std::pair<unsigned, unsigned>> vert_origin; // (object, local index)
    std::vector<PpfVec3u> face;
    std::vector<PpfVec2u> edge;
    std::vector<PpfVec4u> hinge;
    std::vector<std::vector<unsigned>> edge_face, hinge_face, v_face, v_hinge, v_edge, v_rod;
    // props / params
    std::vector<PpfVertexProp> vprop; std::vector<PpfVertexParam> vparam;
    std::vector<PpfEdgeProp> eprop;   std::vector<PpfEdgeParam> eparam;
    std::vector<PpfFaceProp> fprop;   std::vector<PpfFaceParam> fparam;
    std::vector<PpfHingeProp> hprop;  std::vector<PpfHingeParam> hparam;
    std::vector<PpfMat2x2f> inv_rest2x2;
    std::vector<PpfVec3f> curr, prev;
    std::vector<std::vector<unsigned>> fixed_index_table;
    std::vector<std::vector<PpfVec2u>> transpose_table;
    std::vector<unsigned> face_slots, edge_slots, hinge_slots, rod_bend_slots;
    std::vector<char> type_face, type_vertex, type_hinge;
    // zero-filled per-vertex grain arrays (builder.rs:748-762) etc.
    std::vector<PpfVec3f> zero3; std::vector<float> zero1; std::vector<PpfMat3x3f> zero9;
    // per-object contact statistics (S/statistics.rs configure_dataset): filled by enable_statistics(), empty otherwise
    std::vector<unsigned> stat_obj, stat_sobj, stat_count;
    std::vector<unsigned> lock_index;
    // flattened VecVec storage
    OwnedVecVec<unsigned> f_v_face, f_v_hinge, f_v_edge, f_v_rod, f_hinge_face, f_edge_face, f_fixed;
    OwnedVecVec<PpfVec2u> f_transpose;
    // pins resolved to global indices
    struct GPin { std::vector<unsigned> index; std::vector<Vec3f> rest; std::vector<PinOp> ops; double pull_w, unpin_time; unsigned object; };
    std::vector<GPin> pins;
    // constraint storage (rebuilt every step)
    std::vector<PpfFixPair> fix; std::vector<PpfPullPair> pull;
    std::vector<PpfSphere> sphere; std::vector<PpfFloor> floor;
    std::vector<PpfTorqueGroup> torque_groups; std::vector<PpfTorqueVertex> torque_vertices;
    // static collision mesh (builder.rs:1823-1935 make_collision_mesh): all static objects merged
    std::vector<PpfVec3f> cm_vert; std::vector<PpfVec3u> cm_face; std::vector<PpfVec2u> cm_edge;
    std::vector<PpfVertexProp> cm_vprop; std::vector<PpfEdgeProp> cm_eprop; std::vector<PpfFaceProp> cm_fprop;
    std::vector<PpfVertexParam> cm_vparam; std::vector<PpfEdgeParam> cm_eparam; std::vector<PpfFaceParam> cm_fparam;
    PpfCollisionMesh cmesh{};
    std::vector<PpfStitch> stitch; std::vector<unsigned> stitch_slots;
    PpfDataSet dataset;
    PpfParamSet param;
    PpfConstraint constraint;
    unsigned n_vert = 0, n_rod = 0;
};

void build_scene(const SceneSpec &spec, Built &out);            // mesh.rs + scene.rs::make_props + builder.rs::build
void make_constraint(Built &b, double time, double prev_time);  // scene.rs:1944 ; fills b.constraint
void update_param(Built &b, double time);                       // scene.rs update_param (gravity keyframes)

// ---- backend handle: the 20 symbols, resolved from one shared library ----
struct Backend {
    void *handle = nullptr;
    std::string path;
    bool (*initialize)(PpfDataSet *, PpfParamSet *) = nullptr;
    void (*advance)(PpfStepResult *) = nullptr;
    void (*fetch)() = nullptr;
    void (*fetch_inv_rest)() = nullptr;
    void (*fetch_rest_angles)() = nullptr;
    void (*fetch_dyn_counts)(unsigned *, unsigned *) = nullptr;
    void (*fetch_dyn)(unsigned *, PpfMat3x3f *, unsigned *) = nullptr;
    void (*update_dyn)(unsigned *, unsigned *) = nullptr;
    void (*update_constraint)(const PpfConstraint *) = nullptr;
    void (*update_rest_shape)(const PpfRestShapeUpdate *) = nullptr;
    void (*override_velocity)(const unsigned *, unsigned, float, float, float, float) = nullptr;
    void (*gather_current_positions)(const unsigned *, unsigned, float *) = nullptr;
    void (*override_angular_velocity)(const unsigned *, unsigned, float, float, float, float, float, float, float) = nullptr;
    void (*init_collision_windows)(const unsigned *, unsigned, const float *, const unsigned *, unsigned) = nullptr;
    void (*refresh_collision_active)(float) = nullptr;
    unsigned (*fetch_intersection_records)(PpfIntersectionRecord *, unsigned) = nullptr;
    void (*set_log_path)(const char *) = nullptr;
    unsigned char (*ppf_fatal_code)() = nullptr;
    const char *(*ppf_fatal_detail)() = nullptr;
    // optional extensions of the B200 backend
    void (*b200_get_stats)(PpfB200Stats *) = nullptr;
    void (*b200_reset_stats)() = nullptr;
    unsigned (*b200_fetch_pairs)(unsigned, unsigned *, unsigned) = nullptr;
    bool open(const std::string &lib_path, std::string &err);
};

// ---- run loop (backend.rs:626-1025, reduced to what the hot path needs) ----
struct StepRecord {
    double time = 0, wall_ms = 0;
    PpfStepResult result{};
};
struct Runner {
    Built *b = nullptr;
    Backend *be = nullptr;
    double time = 0.0;
    int curr_frame = 0;
    std::vector<StepRecord> steps;
    bool init(const std::string &log_dir);
    // restore a saved state before init (backend.rs:687-694 `--load`)
    bool step();                       // one update_param/update_constraint/refresh/advance
    void fetch_positions(std::vector<float> &curr, std::vector<float> &prev);
    void fetch_dyn_pattern(std::vector<unsigned> &index, std::vector<unsigned> &offset);
};

// state files used to re-seed a run from a reference state (backend.rs:169-178 State)
bool save_state(const std::string &path, const Runner &r, const std::vector<float> &curr,
                const std::vector<float> &prev, const std::vector<unsigned> &dyn_index,
                const std::vector<unsigned> &dyn_offset, float prev_dt);
bool load_state(const std::string &path, double &time, float &prev_dt, std::vector<float> &curr,
                std::vector<float> &prev, std::vector<unsigned> &dyn_index, std::vector<unsigned> &dyn_offset);

} // namespace ppfh

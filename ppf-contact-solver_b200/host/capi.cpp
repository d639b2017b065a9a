// capi.cpp -- flat C API over the C++ host (libppf_host.so) so that Python (ctypes) can drive a
// backend in-process: tests, bench.py and __graft_entry__.smoke() use it.  Every call below maps
// onto the run loop of crates/ppf-cts-solver/src/backend.rs:626-1025 as restated in backend.cpp.
#include "ppf_host.hpp"
#include <chrono>
#include <cstring>

using namespace ppfh;

struct Session {
    Built b;
    Backend be;
    Runner r;
    std::string err;
    std::vector<float> curr, prev;
    std::vector<unsigned> dyn_index, dyn_offset;
};

extern "C" {

__attribute__((visibility("default"))) void print_rust(const char *msg) {
    if (getenv("PPF_HOST_QUIET") == nullptr) fprintf(stderr, "[backend] %s\n", msg);
}

Session *ppfh_create(const char *scene, unsigned n) {
    Session *s = new Session;
    try {
        build_scene(make_scene(scene, n), s->b);
    } catch (const std::exception &e) {
        fprintf(stderr, "ppfh_create: %s\n", e.what());
        delete s;
        return nullptr;
    }
    s->r.b = &s->b;
    s->r.be = &s->be;
    return s;
}
void ppfh_destroy(Session *s) { delete s; }
const char *ppfh_error(Session *s) { return s->err.c_str(); }
// counts: [n_vert, n_face, n_edge, n_hinge, n_fix, fixed_nnz]
void ppfh_counts(Session *s, unsigned *out) {
    out[0] = s->b.n_vert; out[1] = (unsigned)s->b.face.size(); out[2] = (unsigned)s->b.edge.size();
    out[3] = (unsigned)s->b.hinge.size(); out[4] = (unsigned)s->b.fix.size(); out[5] = (unsigned)s->b.f_fixed.data.size();
}
void ppfh_set_args(Session *s, unsigned cg_max_iter, float cg_tol, int disable_contact) {
    if (cg_max_iter) s->b.param.cg_max_iter = cg_max_iter;
    if (cg_tol > 0) s->b.param.cg_tol = cg_tol;
    if (disable_contact >= 0) s->b.param.disable_contact = disable_contact != 0;
}
// ParamSet switches a test needs to flip before initialize(): "disable_pin_dof_removal" (barrier-held fix pins, scene.rs make_param)
int ppfh_set_param_flag(Session *s, const char *name, int value) {
    if (std::string(name) == "disable_pin_dof_removal") { s->b.param.disable_pin_dof_removal = value != 0; return 1; }
    if (std::string(name) == "precond") { s->b.param.precond = value ? PPF_PRECOND_SCHWARZ : PPF_PRECOND_BLOCKJACOBI; return 1; } // 0 block-Jacobi, 1 Schwarz (data.hpp:83)
    return 0;
}
// Per-object contact statistics, the way S/statistics.rs configure_dataset wires them from a session's statistics input: one
// object per scene object (vert_dmap), the static collision mesh as one more object; call before ppfh_init.  Returns the
// number of objects.  ppfh_statistics copies the counts fetch() brought back.
unsigned ppfh_enable_statistics(Session *s) {
    Built &b = s->b;
    const unsigned n_dyn = (unsigned)b.spec.objects.size();
    const unsigned n_static = b.dataset.constraint.mesh.vertex.size;
    b.stat_obj = b.vert_dmap;
    b.stat_sobj.assign(n_static, n_dyn);
    b.stat_count.assign(n_dyn + (n_static ? 1u : 0u), 0u);
    b.dataset.statistics_object_index.data = b.stat_obj.data(); b.dataset.statistics_object_index.size = b.dataset.statistics_object_index.allocated = (unsigned)b.stat_obj.size();
    b.dataset.statistics_static_object_index.data = b.stat_sobj.data(); b.dataset.statistics_static_object_index.size = b.dataset.statistics_static_object_index.allocated = (unsigned)b.stat_sobj.size();
    b.dataset.statistics_contact_count.data = b.stat_count.data(); b.dataset.statistics_contact_count.size = b.dataset.statistics_contact_count.allocated = (unsigned)b.stat_count.size();
    return (unsigned)b.stat_count.size();
}
unsigned ppfh_statistics(Session *s, unsigned *out) {
    if (out && !s->b.stat_count.empty()) memcpy(out, s->b.stat_count.data(), s->b.stat_count.size() * sizeof(unsigned));
    return (unsigned)s->b.stat_count.size();
}
// the DataSet's inv_rest2x2 (4 floats per face) -- starting point of an update_rest_shape in the tests
void ppfh_inv_rest(Session *s, float *out) {
    if (!s->b.inv_rest2x2.empty()) memcpy(out, s->b.inv_rest2x2.data(), s->b.inv_rest2x2.size() * sizeof(s->b.inv_rest2x2[0]));
}
int ppfh_load_state(Session *s, const char *path) {
    double t; float prev_dt;
    std::vector<float> c, p;
    if (!load_state(path, t, prev_dt, c, p, s->dyn_index, s->dyn_offset) || c.size() != size_t(3) * s->b.n_vert) return 0;
    memcpy(s->b.curr.data(), c.data(), c.size() * sizeof(float));
    memcpy(s->b.prev.data(), p.data(), p.size() * sizeof(float));
    s->b.param.time = t;
    s->b.param.prev_dt = prev_dt;
    return 1;
}
int ppfh_open(Session *s, const char *lib) { return s->be.open(lib, s->err) ? 1 : 0; }
int ppfh_init(Session *s, const char *log_dir) {
    const bool ok = s->r.init(log_dir);
    if (ok && !s->dyn_offset.empty() && s->dyn_offset.size() == size_t(s->b.n_vert) + 1 && !s->b.param.disable_contact)
        s->be.update_dyn(s->dyn_index.data(), s->dyn_offset.data());
    return ok ? 1 : 0;
}
// one step through the ABI; returns 1 on success.  wall_ms covers update_constraint + advance.
int ppfh_step(Session *s, double *wall_ms) {
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = s->r.step();
    const auto t1 = std::chrono::steady_clock::now();
    if (wall_ms) *wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    return ok ? 1 : 0;
}
double ppfh_time(Session *s) { return s->r.time; }
int ppfh_frame(Session *s) { return s->r.curr_frame; }
unsigned ppfh_last_flags(Session *s) {
    if (s->r.steps.empty()) return 0;
    const PpfStepResult &r = s->r.steps.back().result;
    return (unsigned)r.ccd_success | ((unsigned)r.pcg_success << 1) | ((unsigned)r.intersection_free << 2) |
           ((unsigned)r.newton_progress << 3) | ((unsigned)r.pin_feasible << 4) | ((unsigned)r.contact_separated << 5);
}
// D2H of the positions through fetch(); out arrays of 3*n_vert floats (may be null to only time the transfer)
void ppfh_fetch(Session *s, float *curr, float *prev) {
    s->r.fetch_positions(s->curr, s->prev);
    if (curr) memcpy(curr, s->curr.data(), s->curr.size() * sizeof(float));
    if (prev) memcpy(prev, s->prev.data(), s->prev.size() * sizeof(float));
}
unsigned ppfh_fetch_dyn(Session *s) {
    s->r.fetch_dyn_pattern(s->dyn_index, s->dyn_offset);
    return (unsigned)s->dyn_index.size();
}
void ppfh_copy_dyn(Session *s, unsigned *index, unsigned *offset) {
    memcpy(index, s->dyn_index.data(), s->dyn_index.size() * sizeof(unsigned));
    memcpy(offset, s->dyn_offset.data(), s->dyn_offset.size() * sizeof(unsigned));
}
int ppfh_save_state(Session *s, const char *path) {
    s->r.fetch_positions(s->curr, s->prev);
    s->r.fetch_dyn_pattern(s->dyn_index, s->dyn_offset);
    return save_state(path, s->r, s->curr, s->prev, s->dyn_index, s->dyn_offset, s->b.param.prev_dt) ? 1 : 0;
}
int ppfh_has_b200_ext(Session *s) { return s->be.b200_get_stats != nullptr; }
void ppfh_b200_stats(Session *s, PpfB200Stats *out) {
    memset(out, 0, sizeof(*out));
    if (s->be.b200_get_stats) s->be.b200_get_stats(out);
}
void ppfh_b200_reset_stats(Session *s) { if (s->be.b200_reset_stats) s->be.b200_reset_stats(); }
unsigned ppfh_b200_pairs(Session *s, unsigned kind, unsigned *out, unsigned cap) {
    return s->be.b200_fetch_pairs ? s->be.b200_fetch_pairs(kind, out, cap) : 0;
}
// bytes a step moves across the ABI: H2D pins (update_constraint), D2H positions (fetch)
void ppfh_io_bytes(Session *s, unsigned long long *h2d, unsigned long long *d2h) {
    *h2d = (unsigned long long)s->b.fix.size() * sizeof(PpfFixPair) + (unsigned long long)s->b.pull.size() * sizeof(PpfPullPair);
    *d2h = 2ull * s->b.n_vert * 12ull;
}
void *ppfh_backend_symbol(Session *s, const char *name);
// pins of the last update_constraint: vertex indices + target positions; returns the count
unsigned ppfh_fix_targets(Session *s, unsigned *index, float *pos) {
    const unsigned n = (unsigned)s->b.fix.size();
    if (index && pos)
        for (unsigned i = 0; i < n; ++i) { index[i] = s->b.fix[i].index; memcpy(pos + 3 * i, s->b.fix[i].position.v, 12); }
    return n;
}
// mesh topology in global order: faces (3 per), edges (2 per), per-vertex (ghat, offset)
void ppfh_topology(Session *s, unsigned *face, unsigned *edge, float *vparam2) {
    if (face) memcpy(face, s->b.face.data(), s->b.face.size() * sizeof(PpfVec3u));
    if (edge) memcpy(edge, s->b.edge.data(), s->b.edge.size() * sizeof(PpfVec2u));
    if (vparam2)
        for (unsigned i = 0; i < s->b.n_vert; ++i) {
            const PpfVertexParam &vp = s->b.vparam[s->b.vprop[i].param_index];
            vparam2[2 * i] = vp.ghat; vparam2[2 * i + 1] = vp.offset;
        }
}
// pins evaluated at (time, prev_time) like the per-step make_constraint: returns the number of fix pins and fills, per pin,
// vertex index, target position, the displacement since prev_time (FixPair.step_delta) and the kinematic flag; initial
// positions of the scene's vertices come with ppfh_initial_positions
unsigned ppfh_constraint_at(Session *s, double time, double prev_time, unsigned *index, float *pos, float *delta, unsigned char *kinematic) {
    make_constraint(s->b, time, prev_time);
    const unsigned n = (unsigned)s->b.fix.size();
    for (unsigned i = 0; i < n && index; ++i) {
        index[i] = s->b.fix[i].index;
        memcpy(pos + 3 * i, s->b.fix[i].position.v, 12);
        memcpy(delta + 3 * i, s->b.fix[i].step_delta.v, 12);
        kinematic[i] = s->b.fix[i].kinematic ? 1 : 0;
    }
    return n;
}
void ppfh_initial_positions(Session *s, float *out) { memcpy(out, s->b.vert.data(), s->b.vert.size() * sizeof(Vec3f)); }
// slot-replay tables of the shell elements (builder.rs:1368-1496): 9 per face, 16 per hinge (remapped order h2, h1, h0, h3),
// plus the hinge quadruples, the fixed index table and, per off-diagonal block, its (row, slot) in the transpose table
void ppfh_slot_tables(Session *s, unsigned *face_slots, unsigned *hinge_slots, unsigned *hinge4, unsigned *fixed_offset, unsigned *fixed_col,
                      unsigned *transpose_offset, unsigned *transpose_rowslot) {
    const Built &b = s->b;
    if (face_slots) memcpy(face_slots, b.face_slots.data(), b.face_slots.size() * sizeof(unsigned));
    if (hinge_slots) memcpy(hinge_slots, b.hinge_slots.data(), b.hinge_slots.size() * sizeof(unsigned));
    if (hinge4) memcpy(hinge4, b.hinge.data(), b.hinge.size() * sizeof(PpfVec4u));
    if (fixed_offset) memcpy(fixed_offset, b.f_fixed.offset.data(), b.f_fixed.offset.size() * sizeof(unsigned));
    if (fixed_col) memcpy(fixed_col, b.f_fixed.data.data(), b.f_fixed.data.size() * sizeof(unsigned));
    if (transpose_offset) memcpy(transpose_offset, b.f_transpose.offset.data(), b.f_transpose.offset.size() * sizeof(unsigned));
    if (transpose_rowslot) memcpy(transpose_rowslot, b.f_transpose.data.data(), b.f_transpose.data.size() * sizeof(PpfVec2u));
}
// the static collision mesh of the built scene (Constraint.mesh): counts[3] = vertices, faces, edges; optional copies of the
// positions (3 per vertex), faces (3), edges (2) and per-vertex / per-edge / per-face (ghat, offset, friction)
void ppfh_static_mesh(Session *s, unsigned *counts, float *vert, unsigned *face, unsigned *edge, float *vpar, float *epar, float *fpar) {
    const Built &b = s->b;
    counts[0] = (unsigned)b.cm_vert.size(); counts[1] = (unsigned)b.cm_face.size(); counts[2] = (unsigned)b.cm_edge.size();
    if (vert) memcpy(vert, b.cm_vert.data(), b.cm_vert.size() * sizeof(PpfVec3f));
    if (face) memcpy(face, b.cm_face.data(), b.cm_face.size() * sizeof(PpfVec3u));
    if (edge) memcpy(edge, b.cm_edge.data(), b.cm_edge.size() * sizeof(PpfVec2u));
    if (vpar) for (size_t i = 0; i < b.cm_vert.size(); ++i) { const PpfVertexParam &q = b.cm_vparam[b.cm_vprop[i].param_index]; vpar[3 * i] = q.ghat; vpar[3 * i + 1] = q.offset; vpar[3 * i + 2] = q.friction; }
    if (epar) for (size_t i = 0; i < b.cm_edge.size(); ++i) { const PpfEdgeParam &q = b.cm_eparam[b.cm_eprop[i].param_index]; epar[3 * i] = q.ghat; epar[3 * i + 1] = q.offset; epar[3 * i + 2] = q.friction; }
    if (fpar) for (size_t i = 0; i < b.cm_face.size(); ++i) { const PpfFaceParam &q = b.cm_fparam[b.cm_fprop[i].param_index]; fpar[3 * i] = q.ghat; fpar[3 * i + 1] = q.offset; fpar[3 * i + 2] = q.friction; }
}
// rod data of the built scene: returns rod_count; per rod edge (rest length, initial length, mass, stiffness, bend),
// per vertex (mass, rest bend angle), the 9 rod-bend slots per vertex and the fixed index table (offset[n_vert+1], columns)
unsigned ppfh_rods(Session *s, float *rod5, float *vert2, unsigned *bend_slots9, unsigned *fixed_offset, unsigned *fixed_col) {
    const Built &b = s->b;
    if (rod5)
        for (unsigned i = 0; i < b.n_rod; ++i) {
            const PpfEdgeProp &p = b.eprop[i];
            const PpfEdgeParam &q = b.eparam[p.param_index];
            rod5[5 * i] = p.length; rod5[5 * i + 1] = p.initial_length; rod5[5 * i + 2] = p.mass; rod5[5 * i + 3] = q.stiffness; rod5[5 * i + 4] = q.bend;
        }
    if (vert2)
        for (unsigned i = 0; i < b.n_vert; ++i) { vert2[2 * i] = b.vprop[i].mass; vert2[2 * i + 1] = b.vprop[i].rest_bend_angle; }
    if (bend_slots9) memcpy(bend_slots9, b.rod_bend_slots.data(), b.rod_bend_slots.size() * sizeof(unsigned));
    if (fixed_offset) memcpy(fixed_offset, b.f_fixed.offset.data(), b.f_fixed.offset.size() * sizeof(unsigned));
    if (fixed_col) memcpy(fixed_col, b.f_fixed.data.data(), b.f_fixed.data.size() * sizeof(unsigned));
    return b.n_rod;
}
}

#include <dlfcn.h>
extern "C" void *ppfh_backend_symbol(Session *s, const char *name) { return s->be.handle ? dlsym(s->be.handle, name) : nullptr; }

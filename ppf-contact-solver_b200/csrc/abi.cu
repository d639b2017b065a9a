// abi.cu -- the exported C ABI (include/ppf_abi.h) of the B200 backend.
//
// Each extern "C" function replaces the same-named entry point of the reference backend
// (crates/ppf-cts-solver/src/cpp/main/main.cu:1658-2350, cited per function in ppf_abi.h).
// There is no CPU fallback: without a CUDA device initialize() stamps a fatal code and exits,
// exactly like the reference (main.cu:1879-1893).
#include "backend.hpp"
#include "math_rod.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstring>

namespace ppf {
State g;
unsigned long long g_launches = 0;
unsigned long long g_dev_alloc = 0, g_dev_free = 0;
cudaStream_t g_alloc_stream = nullptr;
}
using namespace ppf;

extern "C" {
__attribute__((visibility("default"))) unsigned char g_ppf_fatal_code = 0;
__attribute__((visibility("default"))) char g_ppf_fatal_detail[512] = {0};
// weak fallback; the host's #[no_mangle] print_rust wins when present (stub.cpp:8)
__attribute__((weak, visibility("default"))) void print_rust(const char *msg) { (void)msg; }
}

namespace ppf {

void log_info(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    print_rust(buf);
}

[[noreturn]] void fatal(unsigned char code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ppf_fatal_detail, sizeof(g_ppf_fatal_detail), fmt, ap);
    va_end(ap);
    g_ppf_fatal_code = code;
    fprintf(stderr, "PPF FATAL (%u): %s\n", (unsigned)code, g_ppf_fatal_detail);
    fflush(stderr);
    exit(1); // never abort(): the host's atexit hook reads the code (main/fatal.hpp:17-24)
}

// Scope gate: a dataset that holds something this backend does not implement is REFUSED -- initialize() returns false with
// g_ppf_fatal_code = 4 and the reason in g_ppf_fatal_detail, the way the reference reports an initial intersection
// (main.cu:247-263) -- instead of taking the process down: an in-process host (ctypes / PyO3) keeps running and can tell why.
struct ScopeRefusal {};
[[noreturn]] static void refuse(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ppf_fatal_detail, sizeof(g_ppf_fatal_detail), fmt, ap);
    va_end(ap);
    g_ppf_fatal_code = 4;
    fprintf(stderr, "PPF REFUSED (4): %s\n", g_ppf_fatal_detail);
    fflush(stderr);
    throw ScopeRefusal{};
}

// ---------------------------------------------------------------------------------------------
// device-side maintenance kernels
// ---------------------------------------------------------------------------------------------
__global__ void k_clear_pin_index(unsigned *fix_index, unsigned *pull_index, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { fix_index[i] = 0; pull_index[i] = 0; }
}
__global__ void k_set_fix_index(const PpfFixPair *fix, unsigned n, unsigned *fix_index) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fix_index[fix[i].index] = i + 1;
}
__global__ void k_set_pull_index(const PpfPullPair *pull, unsigned n, unsigned *pull_index) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pull_index[pull[i].index] = i + 1;
}
__global__ void k_fixed_flags(const unsigned *fix_index, const uint4 *face, const uint2 *edge, const uint4 *hinge,
                              unsigned nv, unsigned nf, unsigned ne, unsigned nh, unsigned char *vfixed,
                              unsigned char *ffixed, unsigned char *efixed, unsigned char *hfixed) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) vfixed[i] = fix_index[i] > 0;
    if (i < nf) { uint4 f = face[i]; ffixed[i] = fix_index[f.x] > 0 && fix_index[f.y] > 0 && fix_index[f.z] > 0; }
    if (i < ne) { uint2 e = edge[i]; efixed[i] = fix_index[e.x] > 0 && fix_index[e.y] > 0; }
    if (i < nh) { uint4 h = hinge[i]; hfixed[i] = fix_index[h.x] > 0 && fix_index[h.y] > 0 && fix_index[h.z] > 0 && fix_index[h.w] > 0; }
}

// reference: update_constraint rebuilds fix_index / pull_index and every element's `fixed`
// flag on the HOST and re-uploads all prop arrays (main.cu:2011-2076); here only the pin
// arrays cross PCIe and the flags are recomputed on the device.
void apply_constraint(State &s, const PpfConstraint *c) {
    cudaStream_t q = s.stream;
    s.host.constraint = *c;
    s.nfix = c->fix.size;
    s.npull = c->pull.size;
    s.fix.upload(c->fix.data, c->fix.size, q);
    s.pull.upload(c->pull.data, c->pull.size, q);
    s.nstitch = c->stitch.size;
    s.stitches.upload(c->stitch.data, c->stitch.size, q);
    s.ntgroup = c->torque_groups.size; s.ntvert = c->torque_vertices.size;
    s.tgroups.upload(c->torque_groups.data, s.ntgroup, q);
    s.tverts.upload(c->torque_vertices.data, s.ntvert, q);
    if (s.ntgroup) s.tresult.reserve(size_t(7) * s.ntgroup);
    for (unsigned k = 0; k < s.ntgroup; ++k) {
        const PpfTorqueGroup &tg = c->torque_groups.data[k];
        if (tg.vertex_count == 0 || tg.vertex_start + tg.vertex_count > s.ntvert || tg.hint_vertex >= s.nv || tg.axis_component > 2)
            fatal(4, "update_constraint: torque group %u is malformed", k);
    }
    for (unsigned k = 0; k < s.ntvert; ++k)
        if (c->torque_vertices.data[k].index >= s.nv || c->torque_vertices.data[k].group_id >= s.ntgroup) fatal(4, "update_constraint: torque vertex %u is malformed", k);
    s.nsphere = c->sphere.size;
    s.nfloor = c->floor.size;
    s.spheres.upload(c->sphere.data, c->sphere.size, q);
    s.floors.upload(c->floor.data, c->floor.size, q);
    k_clear_pin_index<<<div_up(s.nv, 256), 256, 0, q>>>(s.fix_index.p, s.pull_index.p, s.nv);
    if (s.nfix) k_set_fix_index<<<div_up(s.nfix, 256), 256, 0, q>>>(s.fix.p, s.nfix, s.fix_index.p);
    if (s.npull) k_set_pull_index<<<div_up(s.npull, 256), 256, 0, q>>>(s.pull.p, s.npull, s.pull_index.p);
    recompute_fixed_flags(s);
    LAUNCH_COUNT(3);
    CUDA_OK(cudaStreamSynchronize(q)); // the caller may free/reuse its arrays after we return
}

void recompute_fixed_flags(State &s) {
    unsigned n = std::max(std::max(s.nv, s.nf), std::max(s.ne, s.nh));
    k_fixed_flags<<<div_up(n, 256), 256, 0, s.stream>>>(s.fix_index.p, s.face.p, s.edge.p, s.hinge.p, s.nv, s.nf, s.ne,
                                                       s.nh, s.vfixed.p, s.ffixed.p, s.efixed.p, s.hfixed.p);
    LAUNCH_COUNT(1);
}

static void flatten_vecvec(const PpfVecVecU32 &vv, unsigned rows, std::vector<unsigned> &off, std::vector<unsigned> &idx) {
    off.assign(rows + 1, 0);
    idx.clear();
    if (vv.size == 0) return;
    for (unsigned i = 0; i <= rows && i <= vv.size; ++i) off[i] = vv.offset[i];
    for (unsigned i = vv.size + 1; i <= rows; ++i) off[i] = off[vv.size];
    idx.assign(vv.data, vv.data + off[rows]);
}

void upload_dataset(State &s) {
    const PpfDataSet &d = s.host;
    const PpfParamSet &prm = *s.param;
    cudaStream_t q = s.stream;
    const unsigned nv = s.nv = d.vertex.curr.size;
    const unsigned nf = s.nf = d.mesh.mesh.face.size;
    const unsigned ne = s.ne = d.mesh.mesh.edge.size;
    const unsigned nh = s.nh = d.mesh.mesh.hinge.size;
    s.nsv = d.surface_vert_count;

    // ---- scope gate: refuse loudly what this backend does not implement (DESIGN.md "out of scope") ----
    if (d.rod_count > ne) fatal(4, "rod_count %u exceeds the edge count %u", d.rod_count, ne);
    if (d.rod_count && (d.edge_hess_slots.size < 4u * d.rod_count || d.rod_bend_hess_slots.size != 9u * d.surface_vert_count))
        refuse("B200 backend: slot-replay tables are required (PPF_SLOT_REPLAY=0 is not supported)");
    if (d.mesh.mesh.tet.size) refuse("B200 backend: tetrahedral solids are not implemented");
    if (d.prop.pdrd_body.size) refuse("B200 backend: PDRD rigid bodies are not implemented");
    if (d.constraint.stitch.size && d.stitch_hess_slots.size != 36u * d.constraint.stitch.size)
        refuse("B200 backend: slot-replay tables are required (stitch_hess_slots)");
    if (prm.disable_pin_dof_removal) refuse("B200 backend: barrier-held fix pins (disable_pin_dof_removal) are not implemented");
    if (d.translation_lock.size) refuse("B200 backend: translation/rotation locks are not implemented");
    // per-object contact statistics (S/statistics.rs configure_dataset): optional; every array is empty without the session input
    s.n_stat_obj = d.statistics_contact_count.size;
    if (s.n_stat_obj) {
        if (d.statistics_object_index.size != nv) fatal(4, "statistics_object_index has %u entries for %u vertices", d.statistics_object_index.size, nv);
        s.stat_obj.upload(d.statistics_object_index.data, nv, q);
        if (d.statistics_static_object_index.size) s.stat_sobj.upload(d.statistics_static_object_index.data, d.statistics_static_object_index.size, q);
        s.stat_count.reserve(s.n_stat_obj);
        s.stat_count.zero(s.n_stat_obj, q);
    }
    for (unsigned i = 0; i < d.grain_inv_inertia.size; ++i)
        if (d.grain_inv_inertia.data[i] > 0.0f) refuse("B200 backend: SAND grains are not implemented");
    if (nf != d.shell_face_count || s.nsv != nv) refuse("B200 backend: solid (tet) surfaces and interior vertices are not implemented (shell faces and rods only)");
    for (unsigned i = 0; i < d.param_arrays.face.size; ++i) {
        const PpfFaceParam &fp = d.param_arrays.face.data[i];
        if (fp.model != PPF_MODEL_BARAFFWITKIN && fp.model != PPF_MODEL_ARAP && fp.model != PPF_MODEL_STVK && fp.model != PPF_MODEL_SNHK)
            refuse("B200 backend: shell model %d is not implemented (Baraff-Witkin, ARAP, StVK and SNHk are)", (int)fp.model);
    }
    if (d.face_hess_slots.size != 9u * nf || d.hinge_hess_slots.size != 16u * nh)
        refuse("B200 backend: slot-replay tables are required (PPF_SLOT_REPLAY=0 is not supported)");

    // ---- positions ----
    std::vector<float4> tmp4(nv);
    for (unsigned i = 0; i < nv; ++i) tmp4[i] = make_float4(d.vertex.curr.data[i].v[0], d.vertex.curr.data[i].v[1], d.vertex.curr.data[i].v[2], 0.f);
    s.curr.upload(tmp4.data(), nv, q);
    CUDA_OK(cudaStreamSynchronize(q));
    for (unsigned i = 0; i < nv; ++i) tmp4[i] = make_float4(d.vertex.prev.data[i].v[0], d.vertex.prev.data[i].v[1], d.vertex.prev.data[i].v[2], 0.f);
    s.prev.upload(tmp4.data(), nv, q);
    CUDA_OK(cudaStreamSynchronize(q));
    s.eval_x.reserve(nv); s.target.reserve(nv); s.x_old.reserve(nv); s.velocity.reserve(nv);

    // ---- topology ----
    {
        std::vector<uint4> f4(nf), h4(nh);
        std::vector<uint2> e2(ne);
        for (unsigned i = 0; i < nf; ++i) f4[i] = make_uint4(d.mesh.mesh.face.data[i].v[0], d.mesh.mesh.face.data[i].v[1], d.mesh.mesh.face.data[i].v[2], 0);
        for (unsigned i = 0; i < nh; ++i) h4[i] = make_uint4(d.mesh.mesh.hinge.data[i].v[0], d.mesh.mesh.hinge.data[i].v[1], d.mesh.mesh.hinge.data[i].v[2], d.mesh.mesh.hinge.data[i].v[3]);
        for (unsigned i = 0; i < ne; ++i) e2[i] = make_uint2(d.mesh.mesh.edge.data[i].v[0], d.mesh.mesh.edge.data[i].v[1]);
        s.face.upload(f4.data(), nf, q); s.hinge.upload(h4.data(), nh, q); s.edge.upload(e2.data(), ne, q);
        CUDA_OK(cudaStreamSynchronize(q));
        std::vector<unsigned> off, idx;
        flatten_vecvec(d.mesh.neighbor.vertex.face, nv, off, idx);
        s.vf_off.upload(off.data(), off.size(), q); s.vf_idx.upload(idx.data(), idx.size(), q);
        CUDA_OK(cudaStreamSynchronize(q));
        flatten_vecvec(d.mesh.neighbor.vertex.edge, nv, off, idx);
        s.ve_off.upload(off.data(), off.size(), q); s.ve_idx.upload(idx.data(), idx.size(), q);
        CUDA_OK(cudaStreamSynchronize(q));
        flatten_vecvec(d.mesh.neighbor.edge.face, ne, off, idx);
        s.ef_off.upload(off.data(), off.size(), q); s.ef_idx.upload(idx.data(), idx.size(), q);
        CUDA_OK(cudaStreamSynchronize(q));
    }
    // ---- resolved per-primitive parameters ----
    {
        std::vector<float4> v(nv), e(ne), f(nf), fm(nf), ir(nf), hm(nh);
        std::vector<float2> fs(nf);
        s.has_strainlimit = false;
        s.has_iso_faces = false;
        std::vector<float> va(nv), hs(nh);
        std::vector<unsigned char> vc(nv), fsk(nf);
        for (unsigned i = 0; i < nv; ++i) {
            const PpfVertexProp &p = d.prop.vertex.data[i];
            PpfVertexParam vp{0, 0, 0};
            if (p.param_index < d.param_arrays.vertex.size) vp = d.param_arrays.vertex.data[p.param_index];
            v[i] = make_float4(vp.ghat, vp.offset, vp.friction, p.mass);
            va[i] = p.area;
            vc[i] = p.collider;
        }
        for (unsigned i = 0; i < ne; ++i) {
            const PpfEdgeProp &p = d.prop.edge.data[i];
            const PpfEdgeParam &ep = d.param_arrays.edge.data[p.param_index];
            e[i] = make_float4(ep.ghat, ep.offset, ep.friction, p.mass);
        }
        for (unsigned i = 0; i < nf; ++i) {
            const PpfFaceProp &p = d.prop.face.data[i];
            const PpfFaceParam &fp = d.param_arrays.face.data[p.param_index];
            f[i] = make_float4(fp.ghat, fp.offset, fp.friction, p.mass);
            fm[i] = make_float4(fp.mu, fp.lambda, p.mass, (float)(int)fp.model);
            if (fp.model != PPF_MODEL_BARAFFWITKIN) s.has_iso_faces = true;
            // strainlimiting.cu:63-64: the limit is relaxed by the smallest shrink factor
            fs[i] = make_float2(fp.strainlimit, fp.strainlimit > 0.0f ? (fp.strainlimit + 1.0f) / fminf(fp.shrink_x, fp.shrink_y) - 1.0f : 0.0f);
            if (fp.strainlimit > 0.0f && !(p.rest_excluded || p.collider)) s.has_strainlimit = true;
            const PpfMat2x2f &m = d.inv_rest2x2.data[i];
            ir[i] = make_float4(m.m[0], m.m[1], m.m[2], m.m[3]);
            fsk[i] = p.rest_excluded || p.collider;
        }
        for (unsigned i = 0; i < nh; ++i) {
            const PpfHingeProp &p = d.prop.hinge.data[i];
            const PpfHingeParam &hp = d.param_arrays.hinge.data[p.param_index];
            // energy.cu:463-517
            float areal = 0.0f;
            int cnt = 0;
            for (int k = 0; k < 4; ++k) {
                const PpfVertexProp &vp = d.prop.vertex.data[d.mesh.mesh.hinge.data[i].v[k]];
                if (vp.area > 0.0f) { areal += vp.mass / vp.area; ++cnt; }
            }
            if (cnt > 0) areal /= cnt;
            float bend_dir = hp.bend;
            if (!(p.uv_edge_sin2 < 0.0f)) bend_dir = hp.bend + hp.bend_warp * p.uv_edge_sin2 + hp.bend_weft * (1.0f - p.uv_edge_sin2);
            const float BEND_SCALE = 1.28e-5f;
            float k = (p.area > 1e-12f) ? BEND_SCALE * bend_dir * (p.length * p.length / p.area) * areal : 0.0f;
            if (p.collider || (d.mesh.type.hinge.size > i && (d.mesh.type.hinge.data[i] & 1))) k = 0.0f;
            hs[i] = k;
            hm[i] = make_float4(p.length, p.area, p.rest_angle, bend_dir);
        }
        {
            std::vector<float> fd(nf), hd(nh);
            s.has_fdamp = s.has_hdamp = false;
            for (unsigned i = 0; i < nf; ++i) { fd[i] = d.param_arrays.face.data[d.prop.face.data[i].param_index].deform_damping; s.has_fdamp = s.has_fdamp || fd[i] > 0.0f; }
            for (unsigned i = 0; i < nh; ++i) { hd[i] = d.param_arrays.hinge.data[d.prop.hinge.data[i].param_index].bend_damping; s.has_hdamp = s.has_hdamp || hd[i] > 0.0f; }
            s.fdamp.upload(fd.data(), nf, q); s.hdamp.upload(hd.data(), nh, q);
            std::vector<float2> fpl(nf), hpl(nh);
            s.has_fplast = s.has_hplast = false;
            for (unsigned i = 0; i < nf; ++i) {
                const PpfFaceParam &fp = d.param_arrays.face.data[d.prop.face.data[i].param_index];
                fpl[i] = make_float2(i < d.shell_face_count ? fp.plasticity : 0.0f, fp.plasticity_threshold);
                s.has_fplast = s.has_fplast || fpl[i].x > 0.0f;
            }
            for (unsigned i = 0; i < nh; ++i) {
                const PpfHingeParam &hp = d.param_arrays.hinge.data[d.prop.hinge.data[i].param_index];
                hpl[i] = make_float2(hp.plasticity, hp.plasticity_threshold);
                s.has_hplast = s.has_hplast || hpl[i].x > 0.0f;
            }
            s.fplast.upload(fpl.data(), nf, q); s.hplast.upload(hpl.data(), nh, q);
            std::vector<float> fsh(nf);
            for (unsigned i = 0; i < nf; ++i) { const PpfFaceParam &fp = d.param_arrays.face.data[d.prop.face.data[i].param_index]; fsh[i] = d.prop.face.data[i].collider ? 0.0f : fminf(fp.shrink_x, fp.shrink_y); }
            s.fshrink.upload(fsh.data(), nf, q);
            CUDA_OK(cudaStreamSynchronize(q));
        }
        s.vcp.upload(v.data(), nv, q); s.ecp.upload(e.data(), ne, q); s.fcp.upload(f.data(), nf, q);
        // the static collision mesh (Constraint.mesh; main.cu:1725-1770 copies it once, lbvh.cu:1099-1115 builds its trees once)
        const PpfCollisionMesh &cm = d.constraint.mesh;
        s.cm_nv = cm.vertex.size; s.cm_nf = cm.face.size; s.cm_ne = cm.edge.size;
        if (s.cm_nv) {
            std::vector<float4> cx(s.cm_nv), cv(s.cm_nv), ce(std::max(s.cm_ne, 1u)), cf(std::max(s.cm_nf, 1u));
            std::vector<uint4> f4(std::max(s.cm_nf, 1u));
            std::vector<uint2> e2(std::max(s.cm_ne, 1u));
            for (unsigned i = 0; i < s.cm_nv; ++i) {
                cx[i] = make_float4(cm.vertex.data[i].v[0], cm.vertex.data[i].v[1], cm.vertex.data[i].v[2], 0.f);
                const PpfVertexParam &pa = cm.param_arrays.vertex.data[cm.prop.vertex.data[i].param_index];
                cv[i] = make_float4(pa.ghat, pa.offset, pa.friction, 0.f);
            }
            for (unsigned i = 0; i < s.cm_ne; ++i) {
                e2[i] = make_uint2(cm.edge.data[i].v[0], cm.edge.data[i].v[1]);
                const PpfEdgeParam &pa = cm.param_arrays.edge.data[cm.prop.edge.data[i].param_index];
                ce[i] = make_float4(pa.ghat, pa.offset, pa.friction, 0.f);
            }
            for (unsigned i = 0; i < s.cm_nf; ++i) {
                f4[i] = make_uint4(cm.face.data[i].v[0], cm.face.data[i].v[1], cm.face.data[i].v[2], 0);
                const PpfFaceParam &pa = cm.param_arrays.face.data[cm.prop.face.data[i].param_index];
                cf[i] = make_float4(pa.ghat, pa.offset, pa.friction, 0.f);
            }
            s.cm_x.upload(cx.data(), s.cm_nv, q); s.cm_vcp.upload(cv.data(), s.cm_nv, q);
            s.cm_ecp.upload(ce.data(), ce.size(), q); s.cm_fcp.upload(cf.data(), cf.size(), q);
            s.cm_face.upload(f4.data(), f4.size(), q); s.cm_edge.upload(e2.data(), e2.size(), q);
            CUDA_OK(cudaStreamSynchronize(q));
        }
        {
            std::vector<float> pr(std::max(nf, 1u), 0.0f);
            s.has_pressure = false;
            for (unsigned i = 0; i < nf; ++i) {
                pr[i] = d.param_arrays.face.data[d.prop.face.data[i].param_index].pressure;
                s.has_pressure = s.has_pressure || pr[i] > 0.0f;
            }
            s.fpressure.upload(pr.data(), pr.size(), q);
            CUDA_OK(cudaStreamSynchronize(q));
        }
        s.fmat.upload(fm.data(), nf, q); s.inv_rest.upload(ir.data(), nf, q); s.hmat.upload(hm.data(), nh, q);
        s.fstrain.upload(fs.data(), nf, q);
        s.varea.upload(va.data(), nv, q); s.hinge_stiff.upload(hs.data(), nh, q);
        s.vcollider.upload(vc.data(), nv, q); s.fskip.upload(fsk.data(), nf, q);
        CUDA_OK(cudaStreamSynchronize(q));
    }
    s.has_offsets = false;
    for (unsigned i = 0; i < d.param_arrays.face.size; ++i) s.has_offsets = s.has_offsets || d.param_arrays.face.data[i].offset > 0.0f;
    for (unsigned i = 0; i < d.param_arrays.edge.size; ++i) s.has_offsets = s.has_offsets || d.param_arrays.edge.data[i].offset > 0.0f;
    for (unsigned i = 0; i < d.param_arrays.vertex.size; ++i) s.has_offsets = s.has_offsets || d.param_arrays.vertex.data[i].offset > 0.0f;
    // ---- rods: Hookean edges [0, rod_count) and the bending stencils of interior rod vertices ----
    s.nrod = d.rod_count;
    s.nrodbend = 0;
    if (s.nrod) {
        std::vector<float2> rp(s.nrod);
        for (unsigned i = 0; i < s.nrod; ++i) { // energy.cu:272-302
            const PpfEdgeProp &p = d.prop.edge.data[i];
            const PpfEdgeParam &ep = d.param_arrays.edge.data[p.param_index];
            rp[i] = make_float2(p.length, ep.stiffness > 0.0f ? ep.stiffness * p.mass : 0.0f);
        }
        std::vector<float2> rs(s.nrod);
        s.has_rod_strainlimit = false;
        for (unsigned i = 0; i < s.nrod; ++i) { // strainlimiting.cu:135-139
            const PpfEdgeProp &p = d.prop.edge.data[i];
            const float limit = d.param_arrays.edge.data[p.param_index].strainlimit;
            rs[i] = make_float2(limit > 0.0f && p.initial_length > 0.0f ? limit : 0.0f, p.initial_length);
            if (rs[i].x > 0.0f) s.has_rod_strainlimit = true;
        }
        s.rod_sl.upload(rs.data(), s.nrod, q);
        s.rod_par.upload(rp.data(), s.nrod, q);
        s.rod_slots.upload(d.edge_hess_slots.data, size_t(4) * s.nrod, q);
        CUDA_OK(cudaStreamSynchronize(q));
        // energy.cu:556-641: a vertex with exactly two incident edges and no incident face bends about its neighbours
        const PpfVecVecU32 &vedge = d.mesh.neighbor.vertex.edge, &vface = d.mesh.neighbor.vertex.face;
        std::vector<uint4> el;
        std::vector<float2> par;
        std::vector<unsigned> sl;
        std::vector<float> rbd;
        std::vector<float2> rbp;
        for (unsigned i = 0; i < s.nsv; ++i) {
            if (i >= vedge.size || vedge.offset[i + 1] - vedge.offset[i] != 2) continue;
            if (i < vface.size && vface.offset[i + 1] != vface.offset[i]) continue;
            const unsigned e0 = vedge.data[vedge.offset[i]], e1 = vedge.data[vedge.offset[i] + 1];
            const PpfEdgeProp &p0 = d.prop.edge.data[e0], &p1 = d.prop.edge.data[e1];
            const PpfEdgeParam &q0 = d.param_arrays.edge.data[p0.param_index], &q1 = d.param_arrays.edge.data[p1.param_index];
            const float mass = d.prop.vertex.data[i].mass;
            const float k = rod_bend_coefficient(q0.bend, q1.bend, p0.length, p1.length, mass);
            if (!(mass > 0.0f && k > 0.0f)) continue;
            const PpfVec2u &c0 = d.mesh.mesh.edge.data[e0], &c1 = d.mesh.mesh.edge.data[e1];
            const unsigned j = c0.v[0] == i ? c0.v[1] : c0.v[0], kk = c1.v[0] == i ? c1.v[1] : c1.v[0];
            el.push_back(make_uint4(j, i, kk, 0u));
            par.push_back(make_float2(k, d.prop.vertex.data[i].rest_bend_angle));
            rbd.push_back(0.5f * (q0.bend_damping + q1.bend_damping)); // energy.cu:570-571
            rbp.push_back(make_float2(0.5f * (q0.plasticity + q1.plasticity), 0.5f * (q0.plasticity_threshold + q1.plasticity_threshold))); // plasticity.cu:218-224
            for (int b = 0; b < 9; ++b) sl.push_back(d.rod_bend_hess_slots.data[size_t(9) * i + b]);
        }
        s.nrodbend = (unsigned)el.size();
        s.rb_elem.upload(el.data(), el.size(), q); s.rb_par.upload(par.data(), par.size(), q); s.rb_slots.upload(sl.data(), sl.size(), q);
        s.rb_damp.upload(rbd.data(), rbd.size(), q);
        s.rb_plast.upload(rbp.data(), rbp.size(), q);
        s.has_rb_plast = false;
        for (const float2 &v : rbp) s.has_rb_plast = s.has_rb_plast || v.x > 0.0f;
        s.has_rb_damp = false;
        for (float v : rbd) s.has_rb_damp = s.has_rb_damp || v > 0.0f;
        std::vector<float> rdm(s.nrod);
        s.has_rod_damp = false;
        for (unsigned i = 0; i < s.nrod; ++i) {
            rdm[i] = d.param_arrays.edge.data[d.prop.edge.data[i].param_index].deform_damping;
            s.has_rod_damp = s.has_rod_damp || rdm[i] > 0.0f;
        }
        s.rod_damp.upload(rdm.data(), s.nrod, q);
        CUDA_OK(cudaStreamSynchronize(q));
    }
    s.fix_index.reserve(nv); s.pull_index.reserve(nv);
    s.vfixed.reserve(nv); s.ffixed.reserve(nf); s.efixed.reserve(ne); s.hfixed.reserve(nh);
    s.fix.reserve(std::max(1u, d.constraint.fix.allocated)); s.pull.reserve(std::max(1u, d.constraint.pull.allocated));

    // ---- fixed matrix ----
    {
        const PpfVecVecU32 &t = d.fixed_index_table;
        s.fixed_nnz = t.offset[t.size];
        s.fx_off.upload(t.offset, t.size + 1, q);
        s.fx_col.upload(t.data, s.fixed_nnz, q);
        s.fx_val.reserve(size_t(9) * s.fixed_nnz); s.fx_el.reserve(size_t(9) * s.fixed_nnz);
        s.face_slots.upload(d.face_hess_slots.data, d.face_hess_slots.size, q);
        s.nstitch_slots = d.stitch_hess_slots.size;
        s.stitch_slots.upload(d.stitch_hess_slots.data, d.stitch_hess_slots.size, q);
        s.hinge_slots.upload(d.hinge_hess_slots.data, d.hinge_hess_slots.size, q);
        std::vector<unsigned> ds(nv);
        for (unsigned i = 0; i < nv; ++i) {
            if (t.offset[i + 1] == t.offset[i] || t.data[t.offset[i]] != i) fatal(4, "fixed_index_table row %u lacks its diagonal", i);
            ds[i] = t.offset[i];
        }
        s.fx_diag_slot.upload(ds.data(), nv, q);
        CUDA_OK(cudaStreamSynchronize(q));
    }
    s.diag.reserve(size_t(9) * nv); s.force.reserve(size_t(3) * nv); s.dx.reserve(size_t(3) * nv);
    s.counters.reserve(160); s.scalars.reserve(64);
    s.counters.zero(160, q); s.scalars.zero(64, q);
    if (s.cm_nv) { // trees of the static collision mesh, built once (they use the bounds words of `counters`)
        bvh_build_static(s, s.cm_bvh_face, 0, s.cm_nf, s.cm_x.p, s.cm_face.p, s.cm_edge.p, s.cm_fcp.p);
        bvh_build_static(s, s.cm_bvh_edge, 1, s.cm_ne, s.cm_x.p, s.cm_face.p, s.cm_edge.p, s.cm_ecp.p);
        bvh_build_static(s, s.cm_bvh_vert, 2, s.cm_nv, s.cm_x.p, s.cm_face.p, s.cm_edge.p, s.cm_vcp.p);
        CUDA_OK(cudaStreamSynchronize(q));
    }
    CUDA_OK(cudaMallocHost(&s.h_pinned, 256 * sizeof(float)));
    s.isect_records.reserve(PPF_MAX_INTERSECTION_RECORDS);
    s.h_isect.resize(PPF_MAX_INTERSECTION_RECORDS);
    s.dyn_cap = prm.csrmat_max_nnz;
    s.carried_offset.assign(nv + 1, 0);
    s.carried_index.clear();

    apply_constraint(s, &d.constraint);
}

} // namespace ppf

// ---------------------------------------------------------------------------------------------
// exported entry points
// ---------------------------------------------------------------------------------------------
#define PPF_EXPORT extern "C" __attribute__((visibility("default")))

PPF_EXPORT unsigned char ppf_fatal_code(void) { return g_ppf_fatal_code; }
PPF_EXPORT const char *ppf_fatal_detail(void) { return g_ppf_fatal_detail; }
PPF_EXPORT void set_log_path(const char *data_dir) { simplelog_set_path(data_dir); }

PPF_EXPORT bool initialize(PpfDataSet *dataset, PpfParamSet *param) {
    int num_device = 0;
    cudaError_t e = cudaGetDeviceCount(&num_device);
    if (e != cudaSuccess || num_device == 0)
        fatal(PPF_FATAL_CUDA_DRIVER, "no CUDA device is visible to this process (%s); this backend ships sm_100a code only and has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaDeviceProp props;
    int dev = 0;
    // One process per GPU: the launcher picks the device with PPF_B200_DEVICE (all GPUs stay visible so that
    // NCCL / peer mappings between the ranks can use NVLink); unset = the reference's behaviour (device 0).
    if (const char *dv = getenv("PPF_B200_DEVICE")) {
        const int want = atoi(dv);
        if (want < 0 || want >= num_device) fatal(PPF_FATAL_CUDA_DRIVER, "PPF_B200_DEVICE=%d but %d devices are visible", want, num_device);
        CUDA_OK(cudaSetDevice(want));
    }
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaGetDeviceProperties(&props, dev));
    log_info("cuda: device %d is %s, compute capability %d.%d, %d SMs", dev, props.name, props.major, props.minor, props.multiProcessorCount);
    if (props.major != 10) fatal(PPF_FATAL_CUDA_DRIVER, "this backend is built for sm_100a (B200) only; found compute capability %d.%d", props.major, props.minor);
    State &s = g;
    if (!s.stream) CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!g_alloc_stream && !getenv("PPF_B200_PLAIN_MALLOC")) { // stream-ordered growth of the buffers (backend.hpp DBuf); keep freed blocks in the pool
        int dev_id = 0, pools = 0;
        cudaGetDevice(&dev_id);
        cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, dev_id);
        cudaMemPool_t pool;
        if (pools && cudaDeviceGetDefaultMemPool(&pool, dev_id) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            g_alloc_stream = s.stream;
        }
        cudaGetLastError();
    }
    s.host = *dataset;
    s.param = param;
    // a second initialize() in the same process starts a new session: nothing of the previous scene may leak into its first step
    s.cand_cover_eval = s.bvh_cover_eval = s.bvh_relative = false;
    for (int k = 0; k < 4; ++k) { s.cand[k].count = 0; s.pairs[k].count = 0; s.last_num_contact[k] = 0; }
    s.dyn_valid = false;
    s.dyn_nnz = 0;
    s.carried_index.clear();
    s.carried_offset.clear();
    s.has_cw = s.cw_uploaded = false;
    s.h_isect_count = 0;
    ScopeLog logging("initialize");
    try {
        upload_dataset(s);
    } catch (const ScopeRefusal &) {
        logging.message("### refused: %s", g_ppf_fatal_detail);
        s.initialized = false;
        return false;
    }
    s.initialized = true;
    bool result = true;
    if (!param->disable_contact) {
        logging.push("lbvh build");
        bvh_build(s, s.bvh_face, 0, s.curr.p);
        bvh_build(s, s.bvh_edge, 1, s.curr.p);
        bvh_build(s, s.bvh_vert, 2, s.curr.p);
        logging.pop();
        logging.push("check intersection");
        bvh_refit(s, s.bvh_face, 0, s.prev.p, s.prev.p, 1.0f);
        bvh_refit(s, s.bvh_edge, 1, s.prev.p, s.prev.p, 1.0f);
        bvh_refit(s, s.bvh_vert, 2, s.prev.p, s.prev.p, 1.0f);
        bool ok_prev = check_intersection(s, s.prev.p, *param);
        bvh_refit(s, s.bvh_face, 0, s.curr.p, s.curr.p, 1.0f);
        bvh_refit(s, s.bvh_edge, 1, s.curr.p, s.curr.p, 1.0f);
        bvh_refit(s, s.bvh_vert, 2, s.curr.p, s.curr.p, 1.0f);
        bool ok_curr = ok_prev && check_intersection(s, s.curr.p, *param);
        if (!ok_prev || !ok_curr) {
            logging.message("### intersection detected");
            snprintf(g_ppf_fatal_detail, sizeof(g_ppf_fatal_detail),
                     "the scene is already self-intersecting at t=0; the solver advances from a separated configuration and cannot untangle one");
            g_ppf_fatal_code = PPF_FATAL_INIT_INTERSECTION;
            result = false;
        }
        logging.pop();
    }
    return result;
}

PPF_EXPORT void advance(PpfStepResult *result) {
    if (!g.initialized) fatal(4, "advance() called before initialize()");
    step_advance(g, result);
}

__global__ void k_unpack4(const float4 *src, float *dst, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float4 v = src[i]; dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z; }
}

PPF_EXPORT void fetch(void) {
    State &s = g;
    cudaStream_t q = s.stream;
    const unsigned nv = s.nv;
    // reuse dx as staging (3 floats per vertex)
    k_unpack4<<<div_up(nv, 256), 256, 0, q>>>(s.curr.p, s.dx.p, nv);
    CUDA_OK(cudaMemcpyAsync(s.host.vertex.curr.data, s.dx.p, size_t(12) * nv, cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    k_unpack4<<<div_up(nv, 256), 256, 0, q>>>(s.prev.p, s.dx.p, nv);
    CUDA_OK(cudaMemcpyAsync(s.host.vertex.prev.data, s.dx.p, size_t(12) * nv, cudaMemcpyDeviceToHost, q));
    if (s.n_stat_obj) // main.cu:1933-1936: the per-object contact counts of the last contact assembly
        CUDA_OK(cudaMemcpyAsync(s.host.statistics_contact_count.data, s.stat_count.p, sizeof(unsigned) * s.n_stat_obj, cudaMemcpyDeviceToHost, q));
    CUDA_OK(cudaStreamSynchronize(q));
    LAUNCH_COUNT(2);
}
// plastic creep changes inv_rest and the rest angles on the device: copy them back into the caller's host arrays so that
// a saved state carries them (main.cu:1939-1968)
PPF_EXPORT void fetch_inv_rest(void) {
    State &s = g;
    if (!s.nf || !s.has_fplast) return;
    std::vector<float4> h(s.nf);
    CUDA_OK(cudaMemcpy(h.data(), s.inv_rest.p, sizeof(float4) * s.nf, cudaMemcpyDeviceToHost));
    for (unsigned i = 0; i < s.nf && i < s.host.inv_rest2x2.size; ++i) {
        PpfMat2x2f &m = s.host.inv_rest2x2.data[i];
        m.m[0] = h[i].x; m.m[1] = h[i].y; m.m[2] = h[i].z; m.m[3] = h[i].w;
    }
}
PPF_EXPORT void fetch_rest_angles(void) {
    State &s = g;
    if (s.nh && s.has_hplast) {
        std::vector<float4> h(s.nh);
        CUDA_OK(cudaMemcpy(h.data(), s.hmat.p, sizeof(float4) * s.nh, cudaMemcpyDeviceToHost));
        for (unsigned i = 0; i < s.nh; ++i) s.host.prop.hinge.data[i].rest_angle = h[i].z;
    }
    if (s.nrodbend && s.has_rb_plast) {
        std::vector<uint4> el(s.nrodbend);
        std::vector<float2> par(s.nrodbend);
        CUDA_OK(cudaMemcpy(el.data(), s.rb_elem.p, sizeof(uint4) * s.nrodbend, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(par.data(), s.rb_par.p, sizeof(float2) * s.nrodbend, cudaMemcpyDeviceToHost));
        for (unsigned t = 0; t < s.nrodbend; ++t) s.host.prop.vertex.data[el[t].y].rest_bend_angle = par[t].y;
    }
}

PPF_EXPORT void fetch_dyn_counts(unsigned *n_value, unsigned *n_offset) {
    State &s = g;
    if (s.dyn_valid) build_fetch_pattern(s, s.carried_index, s.carried_offset), s.dyn_valid = false;
    *n_offset = s.nv + 1;
    *n_value = s.carried_offset.empty() ? 0 : s.carried_offset[s.nv];
}
PPF_EXPORT void fetch_dyn(unsigned *index, PpfMat3x3f *value, unsigned *offset) {
    (void)value; // unused by the reference as well (csrmat.cu:136-144)
    State &s = g;
    if (s.dyn_valid) build_fetch_pattern(s, s.carried_index, s.carried_offset), s.dyn_valid = false;
    memcpy(offset, s.carried_offset.data(), sizeof(unsigned) * (s.nv + 1));
    if (!s.carried_index.empty()) memcpy(index, s.carried_index.data(), sizeof(unsigned) * s.carried_index.size());
}
PPF_EXPORT void update_dyn(unsigned *index, unsigned *offset) {
    // The carried pattern only seeds the reference's slab reservation (csrmat.cu:146-153); this
    // backend rebuilds the dynamic pattern from the detected pairs every assembly, so the restored
    // pattern is kept solely so that fetch_dyn round-trips it until the next advance().
    State &s = g;
    s.carried_offset.assign(offset, offset + s.nv + 1);
    s.carried_index.assign(index, index + offset[s.nv]);
    s.dyn_valid = false;
}

PPF_EXPORT void update_constraint(const PpfConstraint *constraint) {
    // the static collision mesh is taken at initialize, where its trees are built once (the reference never rebuilds them either)
    if (constraint->mesh.vertex.size != g.cm_nv || constraint->mesh.face.size != g.cm_nf)
        fatal(4, "update_constraint: the static collision mesh changed size (%u vertices, %u at initialize)", constraint->mesh.vertex.size, g.cm_nv);
    if (36u * constraint->stitch.size > g.nstitch_slots) fatal(4, "update_constraint: %u stitches but slots for %u only", constraint->stitch.size, g.nstitch_slots / 36u);
    apply_constraint(g, constraint);
}
PPF_EXPORT void update_rest_shape(const PpfRestShapeUpdate *update) {
    State &s = g;
    if (update->inv_rest3x3.size || update->exclude_tet.size) refuse("B200 backend: tetrahedral rest shapes are not implemented");
    if (update->inv_rest2x2.size) {
        std::vector<float4> ir(update->inv_rest2x2.size);
        for (unsigned i = 0; i < update->inv_rest2x2.size; ++i) {
            const PpfMat2x2f &m = update->inv_rest2x2.data[i];
            ir[i] = make_float4(m.m[0], m.m[1], m.m[2], m.m[3]);
        }
        s.inv_rest.upload(ir.data(), ir.size(), s.stream);
        CUDA_OK(cudaStreamSynchronize(s.stream));
    }
    if (update->exclude_face.size) {
        std::vector<unsigned char> fsk(s.nf);
        for (unsigned i = 0; i < s.nf; ++i) {
            const bool ex = i < update->exclude_face.size && update->exclude_face.data[i] != 0;
            s.host.prop.face.data[i].rest_excluded = ex;
            fsk[i] = ex || s.host.prop.face.data[i].collider;
        }
        s.fskip.upload(fsk.data(), s.nf, s.stream);
        CUDA_OK(cudaStreamSynchronize(s.stream));
    }
}

__global__ void k_override_velocity(const unsigned *idx, unsigned n, const float4 *curr, float4 *prev, float dx, float dy, float dz) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { unsigned v = idx[i]; float4 c = curr[v]; prev[v] = make_float4(c.x - dx, c.y - dy, c.z - dz, 0.f); }
}
PPF_EXPORT void override_velocity(const unsigned *indices, unsigned count, float vx, float vy, float vz, float dt) {
    if (count == 0 || dt <= 0.0f) return;
    State &s = g;
    DBuf<unsigned> d;
    d.upload(indices, count, s.stream);
    k_override_velocity<<<div_up(count, 256), 256, 0, s.stream>>>(d.p, count, s.curr.p, s.prev.p, vx * dt, vy * dt, vz * dt);
    CUDA_OK(cudaStreamSynchronize(s.stream));
    d.free();
}
PPF_EXPORT void gather_current_positions(const unsigned *indices, unsigned count, float *out) {
    if (count == 0) return;
    State &s = g;
    std::vector<float4> h(s.nv);
    CUDA_OK(cudaMemcpy(h.data(), s.curr.p, sizeof(float4) * s.nv, cudaMemcpyDeviceToHost));
    for (unsigned i = 0; i < count; ++i) { out[3 * i] = h[indices[i]].x; out[3 * i + 1] = h[indices[i]].y; out[3 * i + 2] = h[indices[i]].z; }
}
__global__ void k_override_angular(const unsigned *idx, unsigned n, const float4 *curr, float4 *prev, float wx, float wy, float wz,
                                   float cx, float cy, float cz, float dt) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned v = idx[i];
    float4 c = curr[v], p = prev[v];
    float rx = c.x - cx, ry = c.y - cy, rz = c.z - cz;
    float vwx = wy * rz - wz * ry, vwy = wz * rx - wx * rz, vwz = wx * ry - wy * rx;
    prev[v] = make_float4(p.x - (vwx * dt), p.y - (vwy * dt), p.z - (vwz * dt), 0.f);
}
PPF_EXPORT void override_angular_velocity(const unsigned *indices, unsigned count, float wx, float wy, float wz, float cx,
                                          float cy, float cz, float dt) {
    if (count == 0 || dt <= 0.0f) return;
    State &s = g;
    DBuf<unsigned> d;
    d.upload(indices, count, s.stream);
    k_override_angular<<<div_up(count, 256), 256, 0, s.stream>>>(d.p, count, s.curr.p, s.prev.p, wx, wy, wz, cx, cy, cz, dt);
    CUDA_OK(cudaStreamSynchronize(s.stream));
    d.free();
}
// collision windows (main.cu:2230-2330): vertex -> object map, up to 8 [start, end) windows per object.  The tables are kept on
// the host until the first refresh (the reference's host calls this right after initialize; either order works here).
__global__ void k_cw_refresh(unsigned n, const unsigned *__restrict__ dmap, const float *__restrict__ windows,
                             const unsigned *__restrict__ counts, float time, unsigned char *__restrict__ active) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned dm = dmap[i], cnt = counts[dm];
    bool a = cnt == 0;
    for (unsigned w = 0; w < cnt && !a; ++w) a = time >= windows[dm * 16 + 2 * w] && time < windows[dm * 16 + 2 * w + 1];
    active[i] = a ? 1 : 0;
}
PPF_EXPORT void init_collision_windows(const unsigned *vert_dmap, unsigned vert_count, const float *windows,
                                       const unsigned *window_counts, unsigned n_groups) {
    State &s = g;
    s.h_cw_dmap.assign(vert_dmap, vert_dmap + vert_count);
    s.h_cw_windows.assign(windows, windows + size_t(n_groups) * 16); // MAX_COLLISION_WINDOWS = 8 pairs per object
    s.h_cw_counts.assign(window_counts, window_counts + n_groups);
    for (unsigned i = 0; i < vert_count; ++i)
        if (vert_dmap[i] >= n_groups) fatal(4, "init_collision_windows: vertex %u maps to object %u of %u", i, vert_dmap[i], n_groups);
    s.has_cw = false;
    for (unsigned i = 0; i < n_groups; ++i) s.has_cw |= window_counts[i] != 0;
    s.cw_uploaded = false;
    // until the next refresh_collision_active every vertex takes part in contact (main.cu:2230-2264 memsets the flags to active)
    if (s.has_cw && vert_count) {
        s.cw_active.reserve(vert_count);
        CUDA_OK(cudaMemsetAsync(s.cw_active.p, 1, vert_count, s.stream));
    }
    s.cand_cover_eval = false;
}
PPF_EXPORT void refresh_collision_active(float time) {
    State &s = g;
    if (!s.has_cw) return;
    if (!s.nv || s.h_cw_dmap.size() != s.nv) fatal(4, "refresh_collision_active: %zu window vertices for a scene of %u", s.h_cw_dmap.size(), s.nv);
    if (!s.cw_uploaded) {
        s.cw_dmap.upload(s.h_cw_dmap.data(), s.h_cw_dmap.size(), s.stream);
        s.cw_windows.upload(s.h_cw_windows.data(), s.h_cw_windows.size(), s.stream);
        s.cw_counts.upload(s.h_cw_counts.data(), s.h_cw_counts.size(), s.stream);
        s.cw_active.reserve(s.nv);
        s.cw_uploaded = true;
    }
    k_cw_refresh<<<div_up(s.nv, 256), 256, 0, s.stream>>>(s.nv, s.cw_dmap.p, s.cw_windows.p, s.cw_counts.p, time, s.cw_active.p);
    LAUNCH_COUNT(1);
}

PPF_EXPORT unsigned fetch_intersection_records(PpfIntersectionRecord *out, unsigned max_count) {
    State &s = g;
    unsigned count = std::min(s.h_isect_count, max_count);
    if (count) memcpy(out, s.h_isect.data(), count * sizeof(PpfIntersectionRecord));
    return count;
}

// ---- extensions ----
PPF_EXPORT void ppf_b200_get_stats(PpfB200Stats *out) {
    const State &s = g;
    memset(out, 0, sizeof(*out));
    out->ms_advance = s.tm.ms_advance; out->ms_lbvh_build = s.tm.ms_lbvh_build; out->ms_assembly = s.tm.ms_assembly;
    out->ms_contact = s.tm.ms_contact; out->ms_linsolve = s.tm.ms_linsolve; out->ms_spmv = s.tm.ms_spmv;
    out->ms_aabb_update = s.tm.ms_aabb_update; out->ms_line_search = s.tm.ms_line_search;
    out->n_advance = s.tm.n_advance; out->n_newton = s.tm.n_newton; out->n_cg_iter = s.tm.n_cg_iter; out->n_spmv = s.tm.n_spmv;
    out->n_kernel_launches = g_launches;
    out->spmv_n_row = s.spmv_n_row; out->spmv_n_off = s.spmv_n_off;
    for (int k = 0; k < 4; ++k) out->last_num_contact[k] = s.last_num_contact[k];
}
PPF_EXPORT void ppf_b200_reset_stats(void) { g.tm = Timers{}; g_launches = 0; }

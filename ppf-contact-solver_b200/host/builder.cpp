// builder.cpp -- topology, props and the repr(C) DataSet / ParamSet.
// Restates crates/ppf-cts-solver/src/mesh.rs:66-293, scene.rs:1538-1930 (make_props),
// main.rs:204-258 (non-rod edge props) and builder.rs:339-1681 (build, make_param) for
// shell and rod scenes (no tets / PDRD / SAND / stitches, which none of the BASELINE
// configs contain).  Edge and hinge numbering: the reference iterates a std HashMap
// (random per process, mesh.rs:196-229,257-293); this host uses sorted-key order so a
// scene is reproducible.  Both backends are always fed the same DataSet.
#include "ppf_host.hpp"
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>

namespace ppfh {

template <class T> void OwnedVecVec<T>::from(const std::vector<std::vector<T>> &rows) {
    offset.assign(rows.size() + 1, 0);
    size_t n = 0;
    for (size_t i = 0; i < rows.size(); ++i) {
        offset[i] = (unsigned)n;
        n += rows[i].size();
    }
    offset[rows.size()] = (unsigned)n;
    data.clear();
    data.reserve(n);
    for (auto &r : rows) data.insert(data.end(), r.begin(), r.end());
}
template struct OwnedVecVec<unsigned>;
template struct OwnedVecVec<PpfVec2u>;

template <class V, class T> static void set_vec(V &dst, std::vector<T> &src) {
    dst.data = src.empty() ? nullptr : reinterpret_cast<decltype(dst.data)>(src.data());
    dst.size = dst.allocated = (unsigned)src.size();
}
template <class VV, class T> static void set_vecvec(VV &dst, OwnedVecVec<T> &src) {
    dst.data = src.data.empty() ? nullptr : reinterpret_cast<decltype(dst.data)>(src.data.data());
    dst.offset = src.offset.data();
    dst.size = (unsigned)src.offset.size() - 1;
    dst.nnz = dst.nnz_allocated = (unsigned)src.data.size();
    dst.offset_allocated = (unsigned)src.offset.size();
}

static inline Vec3f sub(const Vec3f &a, const Vec3f &b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
static inline Vec3f cross(const Vec3f &a, const Vec3f &b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
static inline float dot(const Vec3f &a, const Vec3f &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline float norm(const Vec3f &a) { return std::sqrt(dot(a, a)); }

template <class P> static unsigned dedup(std::vector<P> &store, const P &p) { // builder.rs:1749-1759 (first occurrence keeps its index)
    if (store.size() <= 16) { // the common case (a handful of distinct parameter sets): linear scan, no allocation
        for (unsigned i = 0; i < store.size(); ++i)
            if (std::memcmp(&store[i], &p, sizeof(P)) == 0) return i;
    } else if (std::memcmp(&store.back(), &p, sizeof(P)) == 0)
        return (unsigned)store.size() - 1;
    static std::map<const void *, std::map<std::string, unsigned>> index_of; // per store: bytes -> index (large stores)
    auto &m = index_of[&store];
    if (m.size() != store.size()) { // first use on this store, or the store was filled through the linear path / reset
        m.clear();
        for (unsigned i = 0; i < store.size(); ++i) m.emplace(std::string(reinterpret_cast<const char *>(&store[i]), sizeof(P)), i);
    }
    const std::string key(reinterpret_cast<const char *>(&p), sizeof(P));
    auto it = m.find(key);
    if (it != m.end()) return it->second;
    store.push_back(p);
    m.emplace(key, (unsigned)store.size() - 1);
    return (unsigned)store.size() - 1;
}

// scene.rs:1945-2013 + builder.rs:1823-1935 (make_collision_mesh): the static objects merged into one contact-only mesh; face
// parameters per object (deduplicated), edge / vertex parameters = the area-weighted average of the incident faces
static void build_static_mesh(Built &b) {
    const SceneSpec &spec = b.spec;
    const float ws = float(spec.args.world_scaling);
    b.cm_vert.clear(); b.cm_face.clear(); b.cm_edge.clear(); b.cm_vprop.clear(); b.cm_eprop.clear(); b.cm_fprop.clear();
    b.cm_vparam.clear(); b.cm_eparam.clear(); b.cm_fparam.clear();
    b.cmesh = PpfCollisionMesh{};
    if (spec.statics.empty()) return;
    for (const StaticObject &o : spec.statics) {
        const unsigned base = (unsigned)b.cm_vert.size();
        for (const Vec3d &p : o.vert) {
            PpfVec3f q;
            for (int k = 0; k < 3; ++k) q.v[k] = (float(p[k]) + float(o.displacement[k])) * ws;
            b.cm_vert.push_back(q);
        }
        PpfFaceParam pa{};
        pa.model = PPF_MODEL_BARAFFWITKIN; // Model::default()
        pa.friction = float(o.friction); pa.ghat = float(o.contact_gap) * ws; pa.offset = float(o.contact_offset) * ws;
        pa.shrink_x = pa.shrink_y = 1.0f;
        const unsigned pi = dedup(b.cm_fparam, pa);
        for (const auto &t : o.tri) {
            PpfVec3u f; for (int k = 0; k < 3; ++k) f.v[k] = base + t[k];
            const Vec3f x0{b.cm_vert[f.v[0]].v[0], b.cm_vert[f.v[0]].v[1], b.cm_vert[f.v[0]].v[2]}, x1{b.cm_vert[f.v[1]].v[0], b.cm_vert[f.v[1]].v[1], b.cm_vert[f.v[1]].v[2]},
                x2{b.cm_vert[f.v[2]].v[0], b.cm_vert[f.v[2]].v[1], b.cm_vert[f.v[2]].v[2]};
            PpfFaceProp fp{};
            fp.area = 0.5f * norm(cross(sub(x1, x0), sub(x2, x0)));
            if (!(fp.area > 0.0f)) throw std::runtime_error("degenerate static face");
            fp.param_index = pi;
            b.cm_face.push_back(f); b.cm_fprop.push_back(fp);
        }
    }
    const unsigned nv = (unsigned)b.cm_vert.size(), nf = (unsigned)b.cm_face.size();
    // unique undirected edges in order of first appearance over the face rows (mesh.rs:66-123)
    std::map<std::pair<unsigned, unsigned>, unsigned> edge_of;
    std::vector<std::vector<unsigned>> edge_faces, vert_faces(nv);
    for (unsigned i = 0; i < nf; ++i)
        for (int k = 0; k < 3; ++k) {
            vert_faces[b.cm_face[i].v[k]].push_back(i);
            unsigned a = b.cm_face[i].v[k], c = b.cm_face[i].v[(k + 1) % 3];
            if (a > c) std::swap(a, c);
            auto it = edge_of.find({a, c});
            if (it == edge_of.end()) { it = edge_of.emplace(std::make_pair(a, c), (unsigned)b.cm_edge.size()).first; b.cm_edge.push_back(PpfVec2u{{a, c}}); edge_faces.push_back({}); }
            edge_faces[it->second].push_back(i);
        }
    auto weighted = [&](const std::vector<unsigned> &faces, float &ghat, float &offset, float &friction) { // builder.rs:1770-1795
        float gs = 0.0f, os = 0.0f, fs = 0.0f, as = 0.0f;
        for (unsigned j : faces) {
            const PpfFaceParam &pa = b.cm_fparam[b.cm_fprop[j].param_index];
            const float area = b.cm_fprop[j].area;
            gs += area * pa.ghat; os += area * pa.offset; fs += area * pa.friction; as += area;
        }
        ghat = gs / as; offset = os / as; friction = fs / as;
    };
    for (unsigned i = 0; i < b.cm_edge.size(); ++i) {
        PpfEdgeParam pa{};
        weighted(edge_faces[i], pa.ghat, pa.offset, pa.friction);
        PpfEdgeProp ep{};
        ep.param_index = dedup(b.cm_eparam, pa);
        b.cm_eprop.push_back(ep);
    }
    for (unsigned i = 0; i < nv; ++i) {
        if (vert_faces[i].empty()) throw std::runtime_error("static vertex without a face");
        PpfVertexParam pa{};
        weighted(vert_faces[i], pa.ghat, pa.offset, pa.friction);
        PpfVertexProp vp{};
        vp.rest_bend_angle = 3.14159265358979323846f;
        vp.param_index = dedup(b.cm_vparam, pa);
        b.cm_vprop.push_back(vp);
    }
    PpfCollisionMesh &m = b.cmesh;
    set_vec(m.vertex, b.cm_vert); set_vec(m.face, b.cm_face); set_vec(m.edge, b.cm_edge);
    set_vec(m.prop.vertex, b.cm_vprop); set_vec(m.prop.edge, b.cm_eprop); set_vec(m.prop.face, b.cm_fprop);
    set_vec(m.param_arrays.vertex, b.cm_vparam); set_vec(m.param_arrays.edge, b.cm_eparam); set_vec(m.param_arrays.face, b.cm_fparam);
}

void build_scene(const SceneSpec &spec, Built &b) {
    b = Built{};
    b.spec = spec;
    const SimArgs &a = spec.args;

    // ---- index map: global index = order of first appearance in the face rows, object by
    // object; loose vertices follow (index_map.rs:68-153) ----
    std::vector<std::vector<int>> l2g(spec.objects.size());
    for (size_t o = 0; o < spec.objects.size(); ++o) l2g[o].assign(spec.objects[o].vert.size(), -1);
    auto assign = [&](size_t o, unsigned lv) {
        if (l2g[o][lv] < 0) {
            l2g[o][lv] = (int)b.vert_origin.size();
            b.vert_origin.push_back({(unsigned)o, lv});
        }
    };
    for (size_t o = 0; o < spec.objects.size(); ++o) // rods first (index_map.rs:81-94)
        for (auto &e : spec.objects[o].rod)
            for (int k = 0; k < 2; ++k) assign(o, e[k]);
    for (size_t o = 0; o < spec.objects.size(); ++o)
        for (auto &t : spec.objects[o].tri)
            for (int k = 0; k < 3; ++k) assign(o, t[k]);
    for (size_t o = 0; o < spec.objects.size(); ++o)
        for (unsigned lv = 0; lv < spec.objects[o].vert.size(); ++lv) assign(o, lv);
    const unsigned nv = b.n_vert = (unsigned)b.vert_origin.size();
    b.vert.resize(nv);
    b.vert_dmap.resize(nv);
    const float ws = float(a.world_scaling);
    for (unsigned g = 0; g < nv; ++g) {
        auto [o, lv] = b.vert_origin[g];
        const Vec3d &p = spec.objects[o].vert[lv];
        const Vec3d &d = spec.objects[o].displacement;
        b.vert[g] = Vec3f{float(p[0] + d[0]) * ws, float(p[1] + d[1]) * ws, float(p[2] + d[2]) * ws};
        b.vert_dmap[g] = o;
    }
    std::vector<unsigned> face_object;
    for (size_t o = 0; o < spec.objects.size(); ++o)
        for (auto &t : spec.objects[o].tri) {
            b.face.push_back(PpfVec3u{{(unsigned)l2g[o][t[0]], (unsigned)l2g[o][t[1]], (unsigned)l2g[o][t[2]]}});
            face_object.push_back((unsigned)o);
        }
    const unsigned nf = (unsigned)b.face.size();

    // ---- edges and hinges (mesh.rs:188-293), sorted-key order ----
    // One record per (face, side); sorting by the undirected key groups the faces around an edge in
    // ascending face order, which reproduces "first face seen owns the record" of the reference.
    struct Side { unsigned long long key; unsigned face, e0, e1, e2; };
    std::vector<Side> sides;
    sides.reserve(size_t(3) * nf);
    for (unsigned i = 0; i < nf; ++i)
        for (int j = 0; j < 3; ++j) {
            const unsigned e0 = b.face[i].v[j], e1 = b.face[i].v[(j + 1) % 3], e2 = b.face[i].v[(j + 2) % 3];
            const unsigned lo = std::min(e0, e1), hi = std::max(e0, e1);
            sides.push_back(Side{((unsigned long long)lo << 32) | hi, i, e0, e1, e2});
        }
    std::sort(sides.begin(), sides.end(), [](const Side &x, const Side &y) { return x.key != y.key ? x.key < y.key : x.face < y.face; });
    // rod edges come first and keep their own vertex order (mesh.rs:216-220)
    std::vector<unsigned> rod_object;
    for (size_t o = 0; o < spec.objects.size(); ++o)
        for (auto &e : spec.objects[o].rod) {
            b.edge.push_back(PpfVec2u{{(unsigned)l2g[o][e[0]], (unsigned)l2g[o][e[1]]}});
            b.edge_face.push_back({});
            rod_object.push_back((unsigned)o);
        }
    const unsigned nrod = b.n_rod = (unsigned)rod_object.size();
    for (size_t k = 0; k < sides.size();) {
        size_t m = k;
        while (m < sides.size() && sides[m].key == sides[k].key) ++m;
        const Side &s0 = sides[k];
        b.edge.push_back(PpfVec2u{{s0.e0, s0.e1}});
        if (m - k >= 2) b.edge_face.push_back({s0.face, sides[k + 1].face});
        else b.edge_face.push_back({s0.face});
        if (m - k == 2) { // interior edge with exactly two faces -> hinge (more than two: excluded, mesh.rs:268-279)
            b.hinge.push_back(PpfVec4u{{s0.e0, s0.e1, s0.e2, sides[k + 1].e2}});
            b.hinge_face.push_back({s0.face, sides[k + 1].face});
        }
        k = m;
    }
    const unsigned ne = (unsigned)b.edge.size(), nh = (unsigned)b.hinge.size();
    b.v_face.assign(nv, {}); b.v_hinge.assign(nv, {}); b.v_edge.assign(nv, {}); b.v_rod.assign(nv, {});
    for (unsigned i = 0; i < nf; ++i) for (int k = 0; k < 3; ++k) b.v_face[b.face[i].v[k]].push_back(i);
    for (unsigned i = 0; i < nh; ++i) for (int k = 0; k < 4; ++k) b.v_hinge[b.hinge[i].v[k]].push_back(i);
    for (unsigned i = 0; i < ne; ++i) for (int k = 0; k < 2; ++k) b.v_edge[b.edge[i].v[k]].push_back(i);
    for (unsigned i = 0; i < nrod; ++i) for (int k = 0; k < 2; ++k) b.v_rod[b.edge[i].v[k]].push_back(i);

    // ---- face props / params (scene.rs:1656-1830) ----
    std::vector<float> face_area(nf);
    for (unsigned i = 0; i < nf; ++i) { // triutils.rs:9-17
        const Vec3f &x0 = b.vert[b.face[i].v[0]], &x1 = b.vert[b.face[i].v[1]], &x2 = b.vert[b.face[i].v[2]];
        face_area[i] = 0.5f * norm(cross(sub(x1, x0), sub(x2, x0)));
        if (!(face_area[i] > 0.0f)) throw std::runtime_error("degenerate face");
    }
    b.fprop.resize(nf);
    // the parameter set is per object: resolve the string-keyed map once per object, not once per face
    std::vector<PpfFaceParam> obj_fparam(spec.objects.size());
    std::vector<float> obj_density(spec.objects.size(), 0.0f);
    for (size_t o = 0; o < spec.objects.size(); ++o) {
        if (spec.objects[o].tri.empty()) continue;
        const auto &pm = spec.objects[o].param;
        auto get = [&](const char *k) { return float(pm.at(k)); };
        PpfFaceParam p;
        std::memset(&p, 0, sizeof(p));
        p.model = (PpfModel)(int)pm.at("model");
        const float young = get("young-mod"), poiss = get("poiss-rat");
        p.mu = young / (2.0f * (1.0f + poiss)); // builder.rs:1739-1743
        p.lambda = young * poiss / ((1.0f + poiss) * (1.0f - 2.0f * poiss));
        p.friction = get("friction"); p.ghat = get("contact-gap"); p.offset = get("contact-offset");
        p.bend = get("bend"); p.strainlimit = get("strain-limit");
        p.shrink_x = get("shrink-x"); p.shrink_y = get("shrink-y"); p.pressure = get("pressure");
        p.deform_damping = get("deformation-damping"); p.bend_damping = get("bending-damping");
        p.plasticity = get("plasticity"); p.plasticity_threshold = get("plasticity-threshold");
        p.bend_plasticity = get("bend-plasticity"); p.bend_plasticity_threshold = get("bend-plasticity-threshold");
        obj_fparam[o] = p;
        obj_density[o] = get("density");
    }
    for (unsigned i = 0; i < nf; ++i) {
        PpfFaceProp fp;
        std::memset(&fp, 0, sizeof(fp));
        fp.area = face_area[i];
        fp.mass = obj_density[face_object[i]] * face_area[i];
        fp.param_index = dedup(b.fparam, obj_fparam[face_object[i]]);
        b.fprop[i] = fp;
    }
    // ---- non-rod edge props (main.rs:225-258, builder.rs:1800-1821) ----
    b.eprop.resize(ne);
    std::vector<PpfEdgeParam> obj_eparam(spec.objects.size());
    std::vector<float> obj_rod_density(spec.objects.size(), 0.0f), obj_length_factor(spec.objects.size(), 1.0f);
    for (size_t o = 0; o < spec.objects.size(); ++o) { // rod parameters, once per object (scene.rs:1542-1648)
        if (spec.objects[o].rod.empty()) continue;
        const auto &pm = spec.objects[o].rod_param;
        auto get = [&](const char *k) { return float(pm.at(k)); };
        PpfEdgeParam p;
        std::memset(&p, 0, sizeof(p));
        p.stiffness = get("young-mod"); p.bend = get("bend"); p.ghat = get("contact-gap"); p.offset = get("contact-offset");
        p.friction = get("friction"); p.strainlimit = get("strain-limit");
        p.plasticity = get("bend-plasticity"); p.plasticity_threshold = get("bend-plasticity-threshold");
        p.bend_rest_from_geometry = get("bend-rest-from-geometry") != 0.0f;
        p.deform_damping = get("deformation-damping"); p.bend_damping = get("bending-damping");
        if (!(get("density") > 0.0f) || !(p.stiffness > 0.0f) || !(p.ghat > 0.0f)) throw std::runtime_error("rod density / young-mod / contact-gap must be positive");
        obj_eparam[o] = p;
        obj_rod_density[o] = get("density");
        obj_length_factor[o] = get("length-factor");
    }
    for (unsigned i = 0; i < nrod; ++i) { // rod edges
        const PpfEdgeParam &p = obj_eparam[rod_object[i]];
        PpfEdgeProp ep;
        std::memset(&ep, 0, sizeof(ep));
        float length = norm(sub(b.vert[b.edge[i].v[1]], b.vert[b.edge[i].v[0]]));
        if (!(length > 0.0f)) throw std::runtime_error("degenerate rod edge");
        ep.initial_length = length;
        ep.mass = obj_rod_density[rod_object[i]] * length;
        length *= obj_length_factor[rod_object[i]];
        ep.length = length;
        ep.param_index = dedup(b.eparam, p);
        b.eprop[i] = ep;
    }
    for (unsigned i = nrod; i < ne; ++i) {
        float gs = 0, os = 0, fs = 0, as = 0;
        for (unsigned j : b.edge_face[i]) {
            const PpfFaceParam &fp = b.fparam[b.fprop[j].param_index];
            const float ar = face_area[j];
            gs += ar * fp.ghat; os += ar * fp.offset; fs += ar * fp.friction; as += ar;
        }
        PpfEdgeParam p;
        std::memset(&p, 0, sizeof(p));
        p.ghat = gs / as; p.offset = os / as; p.friction = fs / as;
        PpfEdgeProp ep;
        std::memset(&ep, 0, sizeof(ep));
        ep.length = ep.initial_length = norm(sub(b.vert[b.edge[i].v[1]], b.vert[b.edge[i].v[0]]));
        ep.param_index = dedup(b.eparam, p);
        b.eprop[i] = ep;
    }

    // ---- pins -> global indices; initial constraint at (t, t) (main.rs make_constraint(0,0)) ----
    for (size_t o = 0; o < spec.objects.size(); ++o)
        for (auto &pin : spec.objects[o].pins) {
            Built::GPin g;
            g.object = (unsigned)o; g.ops = pin.ops; g.pull_w = pin.pull_w; g.unpin_time = pin.unpin_time;
            for (unsigned lv : pin.index) {
                g.index.push_back((unsigned)l2g[o][lv]);
                const Vec3d &p = spec.objects[o].vert[lv];
                g.rest.push_back(Vec3f{float(p[0]) * ws, float(p[1]) * ws, float(p[2]) * ws});
            }
            b.pins.push_back(std::move(g));
        }
    for (auto &st : spec.stitches) { // seams in global indices (scene.rs:2200-2212)
        PpfStitch q;
        std::memset(&q, 0, sizeof(q));
        for (int k = 0; k < 3; ++k) {
            q.index.v[k] = (unsigned)l2g[st.src][st.src_index[k]]; q.weight.v[k] = float(st.src_weight[k]);
            q.index.v[3 + k] = (unsigned)l2g[st.dst][st.dst_index[k]]; q.weight.v[3 + k] = float(st.dst_weight[k]);
        }
        q.stiffness = float(st.stiffness);
        b.stitch.push_back(q);
    }
    build_static_mesh(b);
    make_constraint(b, 0.0, 0.0);

    // ---- builder.rs:388-432 fixed flags ----
    std::vector<char> is_fixed(nv, 0);
    for (auto &f : b.fix) is_fixed[f.index] = 1;
    for (unsigned i = 0; i < nf; ++i)
        b.fprop[i].fixed = is_fixed[b.face[i].v[0]] && is_fixed[b.face[i].v[1]] && is_fixed[b.face[i].v[2]];
    for (unsigned i = 0; i < ne; ++i) b.eprop[i].fixed = is_fixed[b.edge[i].v[0]] && is_fixed[b.edge[i].v[1]];

    // ---- vertex props (builder.rs:468-644) and params (781-863) ----
    b.vprop.assign(nv, PpfVertexProp{});
    for (auto &vp : b.vprop) { std::memset(&vp, 0, sizeof(vp)); vp.rest_bend_angle = float(M_PI); }
    for (unsigned i = 0; i < b.fix.size(); ++i) b.vprop[b.fix[i].index].fix_index = i + 1;
    for (unsigned i = 0; i < b.pull.size(); ++i) b.vprop[b.pull[i].index].pull_index = i + 1;
    for (unsigned i = 0; i < nf; ++i)
        for (int k = 0; k < 3; ++k) b.vprop[b.face[i].v[k]].area += b.fprop[i].area / 3.0f;
    for (unsigned i = 0; i < nf; ++i)
        for (int k = 0; k < 3; ++k) b.vprop[b.face[i].v[k]].mass += b.fprop[i].mass / 3.0f;
    for (unsigned i = 0; i < nrod; ++i) // builder.rs:625-643
        for (int k = 0; k < 2; ++k) b.vprop[b.edge[i].v[k]].mass += b.eprop[i].mass / 2.0f;
    for (unsigned j = 0; j < nv; ++j) { // builder.rs:521-577: rest bend angle of interior rod vertices from the initial pose
        if (b.v_rod[j].size() != 2 || !b.v_face[j].empty()) continue;
        bool from_geometry = false;
        for (unsigned ei : b.v_rod[j]) from_geometry = from_geometry || b.eparam[b.eprop[ei].param_index].bend_rest_from_geometry;
        if (!from_geometry) continue;
        const PpfVec2u &c0 = b.edge[b.v_rod[j][0]], &c1 = b.edge[b.v_rod[j][1]];
        const unsigned o0 = c0.v[0] == j ? c0.v[1] : c0.v[0], o1 = c1.v[0] == j ? c1.v[1] : c1.v[0];
        const Vec3f e0 = sub(b.vert[o0], b.vert[j]), e1 = sub(b.vert[o1], b.vert[j]);
        const float n0 = norm(e0), n1 = norm(e1);
        if (n0 <= 0.0f || n1 <= 0.0f) continue;
        b.vprop[j].rest_bend_angle = std::acos(std::min(1.0f, std::max(-1.0f, dot(e0, e1) / (n0 * n1))));
    }
    for (unsigned j = 0; j < nv; ++j) {
        float gs = 0, os = 0, fs = 0, wsum = 0;
        for (unsigned fi : b.v_face[j]) {
            const PpfFaceParam &fp = b.fparam[b.fprop[fi].param_index];
            const float w = b.fprop[fi].area;
            gs += w * fp.ghat; os += w * fp.offset; fs += w * fp.friction; wsum += w;
        }
        for (unsigned ei : b.v_edge[j]) { // rod edges weigh in by their rest length (builder.rs:810-826)
            if (ei >= nrod) continue;
            const PpfEdgeParam &ep = b.eparam[b.eprop[ei].param_index];
            const float w = b.eprop[ei].length;
            gs += w * ep.ghat; os += w * ep.offset; fs += w * ep.friction; wsum += w;
        }
        if (wsum > 0.0f) {
            PpfVertexParam p{gs / wsum, os / wsum, fs / wsum};
            b.vprop[j].param_index = dedup(b.vparam, p);
        }
    }
    // ---- hinge props (builder.rs:865-1020) ----
    b.hprop.resize(nh);
    for (unsigned i = 0; i < nh; ++i) {
        const PpfVec4u &h = b.hinge[i];
        float os = 0, gs = 0, bs = 0, as = 0, ds = 0, ps = 0, pts = 0;
        bool all_fixed = true;
        for (unsigned j : b.hinge_face[i]) {
            const PpfFaceParam &fp = b.fparam[b.fprop[j].param_index];
            const float ar = b.fprop[j].area;
            all_fixed = all_fixed && b.fprop[j].fixed;
            os += ar * fp.offset; gs += ar * fp.ghat; bs += ar * fp.bend; as += ar; ds += ar * fp.bend_damping;
            ps += ar * fp.bend_plasticity; pts += ar * fp.bend_plasticity_threshold;
        }
        PpfHingeParam p;
        std::memset(&p, 0, sizeof(p));
        p.bend = all_fixed ? 0.0f : bs / as;
        p.bend_damping = all_fixed ? 0.0f : ds / as; // builder.rs:881-920
        p.plasticity = ps / as; p.plasticity_threshold = pts / as;
        p.ghat = gs / as; p.offset = os / as;
        PpfHingeProp hp;
        std::memset(&hp, 0, sizeof(hp));
        hp.length = norm(sub(b.vert[h.v[0]], b.vert[h.v[1]]));
        hp.area = as; hp.rest_angle = 0.0f; hp.uv_edge_sin2 = -1.0f; hp.fixed = all_fixed;
        hp.param_index = dedup(b.hparam, p);
        b.hprop[i] = hp;
    }
    // ---- inv_rest2x2 (builder.rs:157-229, no UV) ----
    b.inv_rest2x2.resize(nf);
    for (unsigned i = 0; i < nf; ++i) {
        const Vec3f &x0 = b.vert[b.face[i].v[0]], &x1 = b.vert[b.face[i].v[1]], &x2 = b.vert[b.face[i].v[2]];
        const Vec3f d0 = sub(x1, x0), d1 = sub(x2, x0);
        const Vec3f n = cross(d0, d1);
        Vec3f e2 = cross(n, d0);
        const float e2n = norm(e2), d0n = norm(d0);
        e2 = Vec3f{e2[0] / e2n, e2[1] / e2n, e2[2] / e2n};
        const Vec3f e1 = Vec3f{d0[0] / d0n, d0[1] / d0n, d0[2] / d0n};
        // d_mat = [e1; e2] * [d0 d1]  (2x2, column-major m[col*2+row])
        const float m00 = dot(e1, d0), m10 = dot(e2, d0), m01 = dot(e1, d1), m11 = dot(e2, d1);
        const float det = m00 * m11 - m01 * m10;
        if (det == 0.0f) throw std::runtime_error("degenerate shell face");
        PpfMat2x2f inv;
        inv.m[0] = m11 / det;  // (0,0)
        inv.m[1] = -m10 / det; // (1,0)
        inv.m[2] = -m01 / det; // (0,1)
        inv.m[3] = m00 / det;  // (1,1)
        b.inv_rest2x2[i] = inv;
    }
    // ---- fixed_index_table / transpose_table / slot tables (builder.rs:1244-1496) ----
    b.fixed_index_table.assign(nv, {});
    {
        std::vector<std::vector<unsigned>> &t = b.fixed_index_table;
        auto ins = [&](unsigned i, unsigned j) { if (i <= j) t[i].push_back(j); };
        for (unsigned i = 0; i < nv; ++i) ins(i, i);
        for (auto &e : b.edge) for (int p = 0; p < 2; ++p) for (int q = 0; q < 2; ++q) ins(e.v[p], e.v[q]);
        for (auto &f : b.face) for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) ins(f.v[p], f.v[q]);
        for (auto &h : b.hinge) for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q) ins(h.v[p], h.v[q]);
        for (auto &q : b.stitch) for (int p = 0; p < 6; ++p) for (int r = 0; r < 6; ++r) ins(q.index.v[p], q.index.v[r]); // builder.rs:1315-1321
        for (unsigned i = 0; i < nv; ++i) { // rod-bending stencil (j, i, k): the pair (j, k) is no edge (builder.rs:1289-1313)
            if (b.v_edge[i].size() != 2 || !b.v_face[i].empty()) continue;
            const PpfVec2u &c0 = b.edge[b.v_edge[i][0]], &c1 = b.edge[b.v_edge[i][1]];
            const unsigned st[3] = {c0.v[0] == i ? c0.v[1] : c0.v[0], i, c1.v[0] == i ? c1.v[1] : c1.v[0]};
            for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) ins(st[p], st[q]);
        }
        for (auto &row : t) { std::sort(row.begin(), row.end()); row.erase(std::unique(row.begin(), row.end()), row.end()); }
    }
    b.transpose_table.assign(nv, {});
    std::vector<unsigned> row_offset(nv + 1, 0);
    {
        unsigned sum = 0;
        for (unsigned i = 0; i < nv; ++i) {
            row_offset[i] = sum;
            for (unsigned k = 0; k < b.fixed_index_table[i].size(); ++k) {
                const unsigned j = b.fixed_index_table[i][k];
                if (i != j) b.transpose_table[j].push_back(PpfVec2u{{i, sum + k}});
            }
            sum += (unsigned)b.fixed_index_table[i].size();
        }
        row_offset[nv] = sum;
    }
    auto slot_of = [&](unsigned i, unsigned j) -> unsigned {
        if (i > j) return 0xFFFFFFFFu;
        const auto &row = b.fixed_index_table[i];
        auto it = std::lower_bound(row.begin(), row.end(), j);
        if (it == row.end() || *it != j) throw std::runtime_error("slot precompute: block missing");
        return row_offset[i] + unsigned(it - row.begin());
    };
    for (auto &f : b.face) for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) b.face_slots.push_back(slot_of(f.v[p], f.v[q]));
    for (auto &e : b.edge) for (int p = 0; p < 2; ++p) for (int q = 0; q < 2; ++q) b.edge_slots.push_back(slot_of(e.v[p], e.v[q]));
    for (auto &h : b.hinge) {
        const unsigned r[4] = {h.v[2], h.v[1], h.v[0], h.v[3]};
        for (int p = 0; p < 4; ++p) for (int q = 0; q < 4; ++q) b.hinge_slots.push_back(slot_of(r[p], r[q]));
    }
    for (auto &q : b.stitch) for (int p = 0; p < 6; ++p) for (int r = 0; r < 6; ++r) b.stitch_slots.push_back(slot_of(q.index.v[p], q.index.v[r])); // builder.rs:1468-1478
    b.rod_bend_slots.assign(size_t(9) * nv, 0xFFFFFFFFu);
    for (unsigned i = 0; i < nv; ++i) { // builder.rs:1446-1467
        if (b.v_edge[i].size() != 2 || !b.v_face[i].empty()) continue;
        const PpfVec2u &c0 = b.edge[b.v_edge[i][0]], &c1 = b.edge[b.v_edge[i][1]];
        const unsigned st[3] = {c0.v[0] == i ? c0.v[1] : c0.v[0], i, c1.v[0] == i ? c1.v[1] : c1.v[0]};
        for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) b.rod_bend_slots[size_t(9) * i + p * 3 + q] = slot_of(st[p], st[q]);
    }
    b.type_face.assign(nf, 0); b.type_vertex.assign(nv, 0); b.type_hinge.assign(nh, 0);

    // ---- vertex state (builder.rs:1571-1587) ----
    const float dt = std::min(float(a.dt), 0.9999f / float(a.fps)); // builder.rs:273-275
    b.curr.resize(nv); b.prev.resize(nv);
    for (unsigned g = 0; g < nv; ++g) {
        const Vec3d &vel = spec.objects[b.vert_origin[g].first].velocity;
        for (int k = 0; k < 3; ++k) {
            b.curr[g].v[k] = b.vert[g][k];
            b.prev[g].v[k] = b.vert[g][k] - dt * float(vel[k]);
        }
    }
    // ---- per-vertex zero arrays (builder.rs:748-762, 1079) ----
    b.zero3.assign(nv, PpfVec3f{{0, 0, 0}});
    b.zero1.assign(nv, 0.0f);
    b.zero9.assign(nv, PpfMat3x3f{{0, 0, 0, 0, 0, 0, 0, 0, 0}});
    b.lock_index.assign(nv, 0xFFFFFFFFu);

    // ---- assemble the DataSet ----
    b.f_v_face.from(b.v_face); b.f_v_hinge.from(b.v_hinge); b.f_v_edge.from(b.v_edge); b.f_v_rod.from(b.v_rod);
    b.f_hinge_face.from(b.hinge_face); b.f_edge_face.from(b.edge_face);
    b.f_fixed.from(b.fixed_index_table); b.f_transpose.from(b.transpose_table);
    PpfDataSet &d = b.dataset;
    std::memset(&d, 0, sizeof(d));
    set_vec(d.vertex.prev, b.prev); set_vec(d.vertex.curr, b.curr);
    set_vec(d.mesh.mesh.face, b.face); set_vec(d.mesh.mesh.hinge, b.hinge); set_vec(d.mesh.mesh.edge, b.edge);
    set_vecvec(d.mesh.neighbor.vertex.face, b.f_v_face); set_vecvec(d.mesh.neighbor.vertex.hinge, b.f_v_hinge);
    set_vecvec(d.mesh.neighbor.vertex.edge, b.f_v_edge); set_vecvec(d.mesh.neighbor.vertex.rod, b.f_v_rod);
    set_vecvec(d.mesh.neighbor.hinge.face, b.f_hinge_face); set_vecvec(d.mesh.neighbor.edge.face, b.f_edge_face);
    set_vec(d.mesh.type.face, b.type_face); set_vec(d.mesh.type.vertex, b.type_vertex); set_vec(d.mesh.type.hinge, b.type_hinge);
    set_vec(d.prop.vertex, b.vprop); set_vec(d.prop.edge, b.eprop); set_vec(d.prop.face, b.fprop); set_vec(d.prop.hinge, b.hprop);
    set_vec(d.param_arrays.vertex, b.vparam); set_vec(d.param_arrays.edge, b.eparam);
    set_vec(d.param_arrays.face, b.fparam); set_vec(d.param_arrays.hinge, b.hparam);
    set_vec(d.inv_rest2x2, b.inv_rest2x2);
    d.constraint = b.constraint;
    set_vecvec(d.fixed_index_table, b.f_fixed); set_vecvec(d.transpose_table, b.f_transpose);
    d.rod_count = nrod; d.shell_face_count = nf; d.surface_vert_count = nv;
    set_vec(d.grain_omega, b.zero3); set_vec(d.grain_inv_inertia, b.zero1); set_vec(d.grain_torque, b.zero3);
    set_vec(d.grain_ang_stiff, b.zero1); set_vec(d.grain_contact_normal, b.zero3);
    set_vec(d.grain_inv_inertia_center, b.zero1); set_vec(d.grain_omega_prev, b.zero3);
    set_vec(d.grain_A, b.zero9); set_vec(d.grain_B, b.zero9); set_vec(d.grain_grot, b.zero3);
    set_vec(d.face_hess_slots, b.face_slots); set_vec(d.edge_hess_slots, b.edge_slots);
    set_vec(d.hinge_hess_slots, b.hinge_slots); set_vec(d.rod_bend_hess_slots, b.rod_bend_slots);
    set_vec(d.stitch_hess_slots, b.stitch_slots);
    set_vec(d.translation_lock_index, b.lock_index);

    // ---- make_param (builder.rs:1629-1681) ----
    PpfParamSet &p = b.param;
    std::memset(&p, 0, sizeof(p));
    p.time = 0.0;
    p.air_friction = float(a.air_friction); p.air_density = float(a.air_density); p.constraint_tol = float(a.constraint_tol);
    p.prev_dt = dt; p.dt = dt; p.playback = float(a.playback);
    p.min_newton_steps = a.min_newton_steps; p.target_toi = float(a.target_toi);
    p.stitch_length_factor = float(a.stitch_length_factor);
    p.cg_max_iter = a.cg_max_iter; p.cg_tol = float(a.cg_tol); p.line_search_max_t = float(a.line_search_max_t);
    p.ccd_eps = float(a.ccd_eps); p.max_dx = float(a.max_dx); p.eiganalysis_eps = float(a.eiganalysis_eps);
    p.friction_eps = float(a.friction_eps); p.isotropic_air_friction = float(a.isotropic_air_friction);
    for (int k = 0; k < 3; ++k) { p.gravity.v[k] = float(a.gravity[k]); p.wind.v[k] = float(a.wind[k]); }
    p.barrier = a.barrier == "quad" ? PPF_BARRIER_QUAD : a.barrier == "log" ? PPF_BARRIER_LOG : PPF_BARRIER_CUBIC;
    p.friction_mode = a.friction_mode == "max" ? PPF_FRICTION_MAX : a.friction_mode == "mean" ? PPF_FRICTION_MEAN : PPF_FRICTION_MIN;
    p.csrmat_max_nnz = a.csrmat_max_nnz;
    p.fix_xz = float(a.fix_xz * a.world_scaling);
    p.disable_contact = a.disable_contact; p.inactive_momentum = a.inactive_momentum;
    p.precond = a.precond == "schwarz" ? PPF_PRECOND_SCHWARZ : PPF_PRECOND_BLOCKJACOBI;
    p.schwarz_levels = a.schwarz_levels; p.max_newton_steps = a.max_newton_steps;
    p.disable_pin_dof_removal = false;
    update_param(b, 0.0);
}

} // namespace ppfh

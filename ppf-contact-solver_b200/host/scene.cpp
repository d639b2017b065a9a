// scene.cpp -- mesh generators, named scenes, pin operations.
// See ppf_host.hpp for the reference files each function restates.
#include "ppf_host.hpp"
#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace ppfh {

// crates/ppf-cts-core/src/datamodel/mesh.rs:606-615 (dx, ny, dy) + :242-295 (verts, faces)
void gen_cylinder(double r, double min_x, double max_x, unsigned n, std::vector<Vec3d> &V,
                  std::vector<std::array<unsigned, 3>> &F) {
    const double dx = (max_x - min_x) / double(n);
    const double ny_f = 2.0 * M_PI * r / dx;
    unsigned ny = (unsigned)ny_f;
    if (ny < 3) ny = 3;
    const double dy = 2.0 * M_PI / double(ny);
    V.assign(size_t(n + 1) * ny, Vec3d{0, 0, 0});
    for (unsigned j = 0; j < ny; ++j) {
        const double theta = double(j) * dy, s = std::sin(theta), c = std::cos(theta);
        for (unsigned i = 0; i <= n; ++i) V[size_t(n + 1) * j + i] = Vec3d{min_x + double(i) * dx, s * r, c * r};
    }
    F.assign(size_t(2) * n * ny, {0u, 0u, 0u});
    for (unsigned j = 0; j < ny; ++j)
        for (unsigned i = 0; i < n; ++i) {
            const size_t idx = size_t(j) * n + i;
            const unsigned v0 = (n + 1) * j + i, v1 = v0 + 1;
            const unsigned v2 = (n + 1) * ((j + 1) % ny) + (i + 1), v3 = (n + 1) * ((j + 1) % ny) + i;
            if ((i % 2) == (j % 2)) {
                F[2 * idx] = {v1, v2, v0};
                F[2 * idx + 1] = {v3, v0, v2};
            } else {
                F[2 * idx] = {v0, v1, v3};
                F[2 * idx + 1] = {v2, v3, v1};
            }
        }
}

// mesh.rs:181-206 (faces) + :621-667 (verts); square(res, size) = rectangle(res, size, size)
void gen_square(unsigned res, double size, Vec3d ex, Vec3d ey, std::vector<Vec3d> &V,
                std::vector<std::array<unsigned, 3>> &F) {
    const unsigned res_x = res;
    unsigned res_y = (unsigned)(double(res_x) * 1.0);
    if (res_y < 2) res_y = 2;
    const double size_x = size, size_y = size * double(res_y) / double(res_x);
    const double step = std::min(size_x / (double(res_x) - 1.0), size_y / (double(res_y) - 1.0));
    V.assign(size_t(res_x) * res_y, Vec3d{0, 0, 0});
    for (unsigned i = 0; i < res_x; ++i) {
        const double x = -size_x / 2.0 + step * i;
        for (unsigned j = 0; j < res_y; ++j) {
            const double y = -size_y / 2.0 + step * j;
            V[size_t(i) * res_y + j] = Vec3d{ex[0] * x + ey[0] * y, ex[1] * x + ey[1] * y, ex[2] * x + ey[2] * y};
        }
    }
    F.clear();
    for (unsigned j = 0; j + 1 < res_y; ++j)
        for (unsigned i = 0; i + 1 < res_x; ++i) {
            const unsigned v0 = i * res_y + j, v1 = v0 + 1, v2 = v0 + res_y, v3 = v2 + 1;
            if ((i % 2) == (j % 2)) {
                F.push_back({v0, v1, v3});
                F.push_back({v0, v3, v2});
            } else {
                F.push_back({v0, v1, v2});
                F.push_back({v1, v3, v2});
            }
        }
}

// crates/ppf-cts-core/src/datamodel/mesh.rs:359-438: icosahedron refined `subdiv_count` times, edge midpoints pushed to the sphere
void gen_icosphere(double r, unsigned subdiv_count, std::vector<Vec3d> &V, std::vector<std::array<unsigned, 3>> &F) {
    const double phi = (1.0 + std::sqrt(5.0)) / 2.0, nrm = std::sqrt(1.0 + phi * phi);
    const double raw[12][3] = {{-1, phi, 0}, {1, phi, 0}, {-1, -phi, 0}, {1, -phi, 0}, {0, -1, phi}, {0, 1, phi},
                               {0, -1, -phi}, {0, 1, -phi}, {phi, 0, -1}, {phi, 0, 1}, {-phi, 0, -1}, {-phi, 0, 1}};
    static const unsigned base[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
                                         {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    V.clear();
    F.clear();
    for (auto &v : raw) V.push_back(Vec3d{v[0] / nrm, v[1] / nrm, v[2] / nrm});
    for (auto &f : base) F.push_back({f[0], f[1], f[2]});
    for (unsigned it = 0; it < subdiv_count; ++it) {
        std::map<std::pair<unsigned, unsigned>, unsigned> mid;
        std::vector<std::array<unsigned, 3>> nf;
        nf.reserve(F.size() * 4);
        auto get_mid = [&](unsigned i, unsigned j) {
            const auto key = i < j ? std::make_pair(i, j) : std::make_pair(j, i);
            auto f = mid.find(key);
            if (f != mid.end()) return f->second;
            const Vec3d &p0 = V[i], &p1 = V[j];
            const Vec3d m{0.5 * (p0[0] + p1[0]), 0.5 * (p0[1] + p1[1]), 0.5 * (p0[2] + p1[2])};
            const double len = std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
            const unsigned idx = (unsigned)V.size();
            V.push_back(Vec3d{m[0] / len, m[1] / len, m[2] / len});
            mid.emplace(key, idx);
            return idx;
        };
        for (auto &t : F) {
            const unsigned a = t[0], b = t[1], c = t[2];
            const unsigned m0 = get_mid(a, b), m1 = get_mid(b, c), m2 = get_mid(c, a);
            nf.push_back({a, m0, m2});
            nf.push_back({m0, b, m1});
            nf.push_back({m2, m1, c});
            nf.push_back({m0, m1, m2});
        }
        F.swap(nf);
    }
    for (auto &v : V) v = Vec3d{v[0] * r, v[1] * r, v[2] * r};
}

// crates/ppf-cts-core/src/kernels/scene_build/transform.rs:97-120
std::vector<unsigned> grab(const std::vector<Vec3d> &V, Vec3d dir, double eps) {
    std::vector<unsigned> out;
    double max_dot = -1e300;
    for (auto &v : V) max_dot = std::max(max_dot, v[0] * dir[0] + v[1] * dir[1] + v[2] * dir[2]);
    for (unsigned i = 0; i < V.size(); ++i)
        if (V[i][0] * dir[0] + V[i][1] * dir[1] + V[i][2] * dir[2] > max_dot - eps) out.push_back(i);
    return out;
}

static std::map<std::string, double> tri_defaults() { // params.rs:295-420 (ObjectKind::Tri)
    return {{"model", 2.0 /* baraff-witkin */}, {"density", 1.0}, {"young-mod", 1000.0}, {"poiss-rat", 0.35},
            {"bend", 10.0}, {"shrink-x", 1.0}, {"shrink-y", 1.0}, {"contact-gap", 1e-3}, {"contact-offset", 0.0},
            {"strain-limit", 0.0}, {"friction", 0.0}, {"pressure", 0.0}, {"deformation-damping", 0.0}, {"bending-damping", 0.0},
            {"plasticity", 0.0}, {"plasticity-threshold", 0.0}, {"bend-plasticity", 0.0}, {"bend-plasticity-threshold", 0.0}};
}

static std::map<std::string, double> rod_defaults() { // params.rs:295-420 (ObjectKind::Rod)
    return {{"density", 1.0}, {"young-mod", 1e4}, {"bend", 0.0}, {"contact-gap", 1e-3}, {"contact-offset", 1e-3},
            {"strain-limit", 0.0}, {"friction", 0.0}, {"length-factor", 1.0}, {"bend-plasticity", 0.0},
            {"bend-plasticity-threshold", 0.0}, {"bend-rest-from-geometry", 0.0}, {"deformation-damping", 0.0},
            {"bending-damping", 0.0}};
}

// make_woven_cylinder(n, offset, scale) of examples/woven.ipynb and examples/large-woven.ipynb (same code in both):
// v_steps axial strands of n vertices on a cylinder of radius 0.5, and h_steps - 2 closed rings of 3 v_steps
// vertices whose radius weaves over and under the strands by 1.2 * offset.
void gen_woven_cylinder(unsigned n, double offset, double scale, std::vector<std::vector<Vec3d>> &strands, std::vector<bool> &closed) {
    const double dx = 1.0 / double(n - 1), width = 1.25;
    scale = 2.0 * 1.48 * scale;
    const unsigned v_steps = (unsigned)(25.0 * scale);
    const double sep = 0.5;
    strands.clear();
    closed.clear();
    for (unsigned i = 0; i < v_steps; ++i) {
        const double theta = 2.0 * M_PI * double(i) / double(v_steps);
        std::vector<Vec3d> xyz(n);
        for (unsigned k = 0; k < n; ++k) xyz[k] = Vec3d{width * (2.0 * dx * double(k) - 1.0), sep * std::sin(theta), sep * std::cos(theta)};
        strands.push_back(std::move(xyz));
        closed.push_back(false);
    }
    const unsigned h_steps = (unsigned)(30.0 * scale);
    const unsigned ring_steps = v_steps * 3;
    if (ring_steps % 2 != 0) throw std::runtime_error("woven cylinder: ring_steps must be even");
    const double amp = 1.2 * offset, dx_h = 1.0 / double(h_steps - 1);
    const unsigned half_v = v_steps / 2;
    for (unsigned i = 1; i + 1 < h_steps; ++i) {
        const double sgn = (i % 2 == 0) ? 1.0 : -1.0;
        std::vector<Vec3d> xyz(ring_steps);
        for (unsigned j = 0; j < ring_steps; ++j) {
            const double theta = 2.0 * M_PI * double(j) / double(ring_steps);
            const double r = sep + sgn * amp * std::cos(double(half_v) * theta);
            xyz[j] = Vec3d{width * (2.0 * dx_h * double(i) - 1.0), r * std::sin(theta), r * std::cos(theta)};
        }
        strands.push_back(std::move(xyz));
        closed.push_back(true);
    }
}

// The woven family (examples/woven.ipynb, examples/large-woven.ipynb): one rod object per strand; the open
// (axial) strands have both ends pinned with spin [-> scale] -> move_by; the notebooks' random() term of the
// angular velocity is dropped.
static SceneSpec woven_scene(const std::string &name, unsigned n, double offset, double scale, double gap, double contact_offset,
                             double friction, double move_delta, double t_end, double pin_scale, double iso_air, unsigned max_nnz, int frames,
                             double strain_limit = 0.0, double ang_vel = 360.0) {
    SceneSpec s;
    s.name = name;
    std::vector<std::vector<Vec3d>> strands;
    std::vector<bool> closed;
    gen_woven_cylinder(n, offset, scale, strands, closed);
    for (size_t k = 0; k < strands.size(); ++k) {
        ShellObject o;
        o.vert = strands[k];
        const unsigned m = (unsigned)o.vert.size();
        for (unsigned i = 0; i + 1 < m; ++i) o.rod.push_back({i, i + 1});
        if (closed[k]) o.rod.push_back({m - 1, 0u});
        double seg = 0.0; // mean drawn spacing (np.diff: the closing segment of a ring is not counted) times length-factor
        for (unsigned i = 0; i + 1 < m; ++i) {
            const Vec3d &a = o.vert[i], &c = o.vert[i + 1];
            seg += std::sqrt((c[0] - a[0]) * (c[0] - a[0]) + (c[1] - a[1]) * (c[1] - a[1]) + (c[2] - a[2]) * (c[2] - a[2]));
        }
        seg = seg / double(m - 1) * 0.8;
        o.rod_param = rod_defaults();
        o.rod_param["bend"] = 1e-3 * (seg / 1e-2) * (seg / 1e-2);
        o.rod_param["young-mod"] = 1e5;
        o.rod_param["contact-gap"] = gap;
        o.rod_param["contact-offset"] = contact_offset;
        o.rod_param["friction"] = friction;
        o.rod_param["length-factor"] = 0.8;
        o.rod_param["strain-limit"] = strain_limit;
        if (!closed[k])
            for (int side = 0; side < 2; ++side) {
                const double sg = side == 0 ? -1.0 : 1.0;
                Pin p;
                p.index = grab(o.vert, Vec3d{sg, 0, 0});
                PinOp spin; spin.kind = PinOp::Spin; spin.axis = Vec3d{-sg, 0, 0}; spin.angular_velocity = ang_vel;
                PinOp mv;   mv.kind = PinOp::MoveBy; mv.delta = Vec3d{-sg * move_delta, 0, 0}; mv.t_start = 0; mv.t_end = t_end;
                if (pin_scale != 1.0) {
                    PinOp scl; scl.kind = PinOp::Scale; scl.factor = pin_scale; scl.t_start = 0; scl.t_end = t_end;
                    p.ops = {spin, scl, mv};
                } else p.ops = {spin, mv};
                o.pins.push_back(p);
            }
        s.objects.push_back(std::move(o));
    }
    s.args.dt = 1.0 / 120.0;
    s.args.frames = frames;
    s.args.gravity = Vec3d{0, 0, 0};
    s.args.isotropic_air_friction = iso_air;
    s.args.friction_mode = "max";
    s.args.csrmat_max_nnz = max_nnz;
    return s;
}

// The twist family (examples/twist.ipynb, examples/large-twist.ipynb): one cylinder, both end
// rings pinned with spin -> scale -> move_by applied in declaration order.
static SceneSpec twist_scene(const std::string &name, double r, double half_len, unsigned n, double bend, double gap,
                             double young, double poiss, double density, double ang_vel, double move_delta,
                             double t_end, double scale_rate, unsigned max_nnz, const char *friction_mode, int frames) {
    SceneSpec s;
    s.name = name;
    ShellObject o;
    gen_cylinder(r, -half_len, half_len, n, o.vert, o.tri);
    o.param = tri_defaults();
    o.param["bend"] = bend;
    o.param["contact-gap"] = gap;
    o.param["young-mod"] = young;
    o.param["poiss-rat"] = poiss;
    o.param["density"] = density;
    const double scale = 1.0 - t_end * scale_rate;
    for (int side = 0; side < 2; ++side) {
        const double sg = side == 0 ? -1.0 : 1.0;
        Pin p;
        p.index = grab(o.vert, Vec3d{sg, 0, 0});
        PinOp spin;  spin.kind = PinOp::Spin;  spin.axis = Vec3d{-sg, 0, 0}; spin.angular_velocity = ang_vel;
        PinOp scl;   scl.kind = PinOp::Scale;  scl.factor = scale; scl.t_start = 0; scl.t_end = t_end;
        PinOp mv;    mv.kind = PinOp::MoveBy;  mv.delta = Vec3d{-sg * move_delta, 0, 0}; mv.t_start = 0; mv.t_end = t_end;
        p.ops = {spin, scl, mv};
        o.pins.push_back(p);
    }
    s.objects.push_back(o);
    s.args.dt = 1.0 / 60.0;
    s.args.frames = frames;
    s.args.gravity = Vec3d{0, 0, 0};
    s.args.csrmat_max_nnz = max_nnz;
    s.args.friction_mode = friction_mode;
    return s;
}

SceneSpec make_scene(const std::string &name, unsigned n) {
    // "<scene>@quad" / "<scene>@log": the same scene with the quadratic / logarithmic barrier (param.rs Barrier, ParamSet.barrier)
    for (const char *suffix : {"@quad", "@log", "@cubic"}) {
        const std::string sfx = suffix;
        if (name.size() > sfx.size() && name.compare(name.size() - sfx.size(), sfx.size(), sfx) == 0) {
            SceneSpec s = make_scene(name.substr(0, name.size() - sfx.size()), n);
            s.name = name;
            s.args.barrier = sfx.substr(1);
            return s;
        }
    }
    if (name == "twist") // examples/twist.ipynb
        return twist_scene(name, 0.5, 1.0, n ? n : 360, 500.0, 4e-3, 1e4, 0.35, 1.0, 360.0, 0.2, 10.0, 0.01, 35000000u, "min", 430);
    if (name == "large-twist") // examples/large-twist.ipynb (headline)
        return twist_scene(name, 0.9, 1.75, n ? n : 1410, 500.0 / 3.5, 1.75e-3, 1e4, 0.25, 3.5, 180.0, 0.5, 16.66, 0.015,
                           160000000u, "max", 2000);
    if (name == "fast-twist") { // small, contact-rich within ~1 s: test scene (same family, faster spin, larger gap)
        unsigned nn = n ? n : 48;
        SceneSpec s = twist_scene(name, 0.5, 1.0, nn, 500.0, 2.0 / nn * 0.35, 1e4, 0.35, 1.0, 720.0, 0.2, 10.0, 0.01, 4000000u, "min", 120);
        return s;
    }
    if (name == "five-twist") { // examples/five-twist.ipynb (.jitter() left out: SURVEY.md 8d asks for a fixed scene)
        SceneSpec s;
        s.name = name;
        const double half_length = 2.6;
        const unsigned nn = n ? n : 180;
        for (int i = 0; i < 5; ++i) {
            const double t = 2.0 * i / 5.0 * M_PI;
            const double x = 0.0, y = std::cos(t), z = std::sin(t); // .at(0, x, y) of the notebook
            ShellObject o;
            gen_cylinder(0.55, -half_length, half_length, nn, o.vert, o.tri);
            o.param = tri_defaults();
            o.param["bend"] = 500.0 / 3.5;
            o.param["contact-gap"] = n ? 5.2 / nn * 0.14 : 4e-3; // the notebook's 4e-3 at n=180 (0.14 edge lengths), scaled when n is overridden
            o.param["young-mod"] = 1e4;
            o.param["poiss-rat"] = 0.25;
            o.param["density"] = 3.5;
            o.displacement = Vec3d{x, y, z};
            const double ang = 360.0, move_delta = 1.0, scale = 0.75, t_end = 25.0, t_wait = 5.0;
            for (int side = 0; side < 2; ++side) {
                const double sg = side == 0 ? -1.0 : 1.0; // left pin first
                Pin p;
                p.index = grab(o.vert, Vec3d{sg, 0, 0});
                PinOp spin; spin.kind = PinOp::Spin; spin.axis = Vec3d{-sg, 0, 0}; spin.angular_velocity = ang; spin.t_start = 0; spin.t_end = t_wait;
                PinOp mv; mv.kind = PinOp::MoveBy; mv.delta = Vec3d{-sg * move_delta, 0, 0}; mv.t_start = 0; mv.t_end = t_end;
                PinOp scl; scl.kind = PinOp::Scale; scl.factor = scale; scl.t_start = 0; scl.t_end = t_wait; scl.center = Vec3d{-x + sg * half_length, -y, -z};
                PinOp orbit; orbit.kind = PinOp::Spin; orbit.center = Vec3d{-x, -y, -z}; orbit.axis = Vec3d{-sg, 0, 0}; orbit.angular_velocity = 180.0;
                orbit.t_start = t_wait; orbit.t_end = 15.0;
                p.ops = {spin, mv, scl, orbit};
                o.pins.push_back(p);
            }
            s.objects.push_back(o);
        }
        s.args.dt = 1.0 / 60.0;
        s.args.frames = 480;
        s.args.gravity = Vec3d{0, 0, 0};
        s.args.friction_mode = "max";
        s.args.csrmat_max_nnz = 40000000u;
        return s;
    }
    if (name == "large-five-twist") { // examples/large-five-twist.ipynb (8.2 M vertices at its own resolution; README.md:616)
        SceneSpec s;
        s.name = name;
        const double half_length = 2.0 / 0.75, rad = 1.75;
        const unsigned nn = n ? n : 1181;
        for (int i = 0; i < 5; ++i) {
            const double t = 2.0 * i / 5.0 * M_PI;
            const double x = 0.0, y = rad * std::cos(t), z = rad * std::sin(t); // .at(0, x, y) of the notebook
            ShellObject o;
            gen_cylinder(1.0, -half_length, half_length, nn, o.vert, o.tri);
            o.param = tri_defaults();
            o.param["bend"] = 500.0 / 3.5;
            o.param["contact-gap"] = n ? 2.0 * half_length / nn * 0.3875 : 1.75e-3; // the notebook's 1.75e-3 at n=1181 (0.39 edge lengths), scaled with n
            o.param["young-mod"] = 1e4;
            o.param["poiss-rat"] = 0.25;
            o.param["density"] = 3.5;
            o.displacement = Vec3d{x, y, z};
            const double ang = 180.0, move_delta = 1.0, scale = 0.25, t_end = 34.0, t_wait = 500.0 / 60.0, s_wait = 16.0;
            for (int side = 0; side < 2; ++side) {
                const double sg = side == 0 ? -1.0 : 1.0; // left pin first
                Pin p;
                p.index = grab(o.vert, Vec3d{sg, 0, 0});
                PinOp spin; spin.kind = PinOp::Spin; spin.axis = Vec3d{-sg, 0, 0}; spin.angular_velocity = ang; spin.t_start = 0; spin.t_end = t_wait;
                PinOp scl; scl.kind = PinOp::Scale; scl.factor = scale; scl.t_start = 0; scl.t_end = s_wait; scl.center = Vec3d{-x + sg * half_length, -y, -z};
                PinOp mv; mv.kind = PinOp::MoveBy; mv.delta = Vec3d{-sg * move_delta, 0, 0}; mv.t_start = 0; mv.t_end = t_end;
                PinOp orbit; orbit.kind = PinOp::Spin; orbit.center = Vec3d{-x, -y, -z}; orbit.axis = Vec3d{-sg, 0, 0}; orbit.angular_velocity = 180.0;
                orbit.t_start = t_wait; orbit.t_end = 150.0;
                p.ops = {spin, scl, mv, orbit};
                o.pins.push_back(p);
            }
            s.objects.push_back(std::move(o));
        }
        s.args.dt = 1.0 / 60.0;
        s.args.frames = 2400;
        s.args.gravity = Vec3d{0, 0, 0};
        s.args.friction_mode = "max";
        s.args.csrmat_max_nnz = 90000000u;
        return s;
    }
    if (name == "curtain" || name == "drape-sheets") {
        // examples/curtain.ipynb: fifteen curtains pinned along their top edge, a fully pinned sphere mesh pushed through them;
        // examples/drape.ipynb: five sheets pinned at two corners settling onto a fully pinned sphere mesh (.jitter() left out)
        const bool curtain = name == "curtain";
        SceneSpec s;
        s.name = name;
        const unsigned res = n ? n : (curtain ? 64 : 128);
        const unsigned subdiv = (n == 0 || n >= 64) ? 4u : (n >= 32 ? 3u : 2u);
        for (int i = 0; i < (curtain ? 15 : 5); ++i) {
            ShellObject o;
            if (curtain) gen_square(res, 2.0, Vec3d{0, 0, 1}, Vec3d{0, 1, 0}, o.vert, o.tri);
            else gen_square(res, 2.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
            o.param = tri_defaults();
            o.param["strain-limit"] = 0.05;
            o.displacement = curtain ? Vec3d{i * 0.25, 0, 0} : Vec3d{0, 0.05 * i, 0};
            Pin p;
            if (curtain) p.index = grab(o.vert, Vec3d{0, 1, 0});
            else {
                p.index = grab(o.vert, Vec3d{1, 0, -1});
                auto c = grab(o.vert, Vec3d{-1, 0, -1});
                p.index.insert(p.index.end(), c.begin(), c.end());
            }
            o.pins.push_back(p);
            s.objects.push_back(std::move(o));
        }
        ShellObject sp;
        gen_icosphere(0.5, subdiv, sp.vert, sp.tri);
        sp.param = tri_defaults();
        sp.displacement = curtain ? Vec3d{-1, 0, 0} : Vec3d{0, -0.5 - 0.05, 0};
        Pin p;
        for (unsigned k = 0; k < sp.vert.size(); ++k) p.index.push_back(k);
        if (curtain) {
            PinOp mv; mv.kind = PinOp::MoveBy; mv.delta = Vec3d{8, 0, 0}; mv.t_start = 0.0; mv.t_end = 5.0;
            p.ops = {mv};
        }
        sp.pins.push_back(p);
        s.objects.push_back(std::move(sp));
        if (!curtain) { s.args.dt = 0.01; s.args.frames = 100; }
        return s;
    }
    if (name == "belt") { // examples/belt.ipynb: three pinned rollers move outward and spin, tensioning three strain-limited belts (friction 0.5)
        SceneSpec s;
        s.name = name;
        auto rotate_y90 = [](std::vector<Vec3d> &V) { for (auto &v : V) v = Vec3d{v[2], v[1], -v[0]}; }; // .rotate(90, "y")
        const unsigned nroll = 21; // the scene exists at the notebook's resolution only (the 2 mm spacing of the nested belts needs it)
        (void)n;
        for (int i = 0; i < 3; ++i) {
            const double t = 2.0 * i * M_PI / 3.0, x = 0.65 * std::sin(t), y = 0.65 * std::cos(t);
            ShellObject o;
            gen_cylinder(0.5, -0.5, 0.5, nroll, o.vert, o.tri);
            rotate_y90(o.vert);
            o.param = tri_defaults();
            o.param["friction"] = 0.5;
            o.displacement = Vec3d{x, y, 3e-3 * i}; // the notebook's 1e-2 * random() jitter, fixed
            Pin p;
            for (unsigned k = 0; k < o.vert.size(); ++k) p.index.push_back(k);
            PinOp spin; spin.kind = PinOp::Spin; spin.axis = Vec3d{0, 0, 1}; spin.angular_velocity = 60.0; spin.t_start = 1.5; spin.t_end = 3.0;
            PinOp mv; mv.kind = PinOp::MoveBy; mv.delta = Vec3d{0.511 * x, 0.511 * y, 0}; mv.t_start = 0.0; mv.t_end = 1.0;
            p.ops = {spin, mv};
            o.pins.push_back(p);
            s.objects.push_back(std::move(o));
        }
        for (int k = 0; k < 3; ++k) {
            const double r = 1.25 + 2e-3 * k, hw = 0.4 - 0.05 * k;
            const unsigned nb = (unsigned)(44 * hw);
            ShellObject o;
            gen_cylinder(r, -hw, hw, nb, o.vert, o.tri);
            rotate_y90(o.vert);
            o.param = tri_defaults();
            o.param["strain-limit"] = 0.06;
            s.objects.push_back(std::move(o));
        }
        s.args.dt = 0.01;
        s.args.frames = 200;
        s.args.gravity = Vec3d{0, 0, 0};
        s.args.min_newton_steps = 32;
        s.args.friction_mode = "max";
        return s;
    }
    if (name == "yarn") { // examples/yarn.ipynb: make_strands(4e-3, 0.5, res): interlaced yarns pulled apart at both ends, closed loops along the edges
        SceneSpec s;
        s.name = name;
        const double res = 1.0, width = 0.5; // the notebook's make_strands(4e-3, 0.5, 1.0); the weave does not scale with a resolution override
        (void)n;
        const double offset = 1.3 * 4e-3;
        const unsigned n_vertical = (unsigned)(res * 0.28 / offset), n_per_seg = (unsigned)(res * 20), n_segs = (unsigned)(width * n_vertical);
        const unsigned n_points = n_segs * n_per_seg;
        const double dx = 1.0 / double(n_points - 1);
        auto add_rod = [&](std::vector<Vec3d> V, bool closed) {
            ShellObject o;
            o.vert = std::move(V);
            const unsigned m = (unsigned)o.vert.size();
            for (unsigned i = 0; i + 1 < m; ++i) o.rod.push_back({i, i + 1});
            if (closed) o.rod.push_back({m - 1, 0u});
            o.rod_param = rod_defaults();
            o.rod_param["bend"] = 0.0;
            o.rod_param["young-mod"] = 1e5;
            o.rod_param["contact-gap"] = 1e-3;
            o.rod_param["contact-offset"] = 2.3e-3;
            o.rod_param["length-factor"] = 0.85;
            if (!closed)
                for (int side = 0; side < 2; ++side) {
                    const double sg = side == 0 ? -1.0 : 1.0;
                    Pin p;
                    p.index = grab(o.vert, Vec3d{sg, 0, 0});
                    PinOp mv; mv.kind = PinOp::MoveBy; mv.delta = Vec3d{sg * 5.0, 0, 0}; mv.t_start = 0; mv.t_end = 10.0;
                    p.ops = {mv};
                    o.pins.push_back(p);
                }
            s.objects.push_back(std::move(o));
        };
        for (unsigned k = 0; k < n_vertical; ++k) {
            const double y_base = n_vertical > 1 ? double(k) / double(n_vertical - 1) - 0.5 : 0.0;
            std::vector<Vec3d> xyz(n_points);
            for (unsigned j = 0; j < n_points; ++j) {
                const double x_mod = double(j % n_per_seg) / double(n_per_seg), t = 2.0 * M_PI * x_mod;
                xyz[j] = Vec3d{width * (2.0 * dx * j - 1.0) - width * (0.5 / n_segs) * std::sin(2.0 * t), y_base + (0.85 / n_vertical) * std::sin(t),
                               0.75 * offset * std::cos(2.0 * t)};
            }
            add_rod(std::move(xyz), false);
        }
        for (int pos = 0; pos < 2; ++pos)
            for (unsigned k = 0; k + 1 < n_segs; ++k) {
                const double dxl = 2.0 * width / n_segs;
                const double y_base = pos == 0 ? 0.5 + 0.25 * dxl : -0.5 - 0.25 * dxl, z_base = 0.15 * dxl;
                double x_center = dxl * (k + 0.77) - width;
                if (pos == 1) x_center += 0.5 * dxl;
                const double r = 0.78 * width / n_segs, theta = 0.25 * M_PI;
                std::vector<Vec3d> xyz(n_per_seg);
                for (unsigned j = 0; j < n_per_seg; ++j) {
                    const double t = 2.0 * M_PI * j / double(n_per_seg), zv = r * std::cos(t);
                    xyz[j] = Vec3d{x_center + r * std::sin(t), pos == 0 ? y_base + zv * std::sin(theta) : y_base - zv * std::sin(theta), z_base + zv * std::cos(theta)};
                }
                add_rod(std::move(xyz), true);
            }
        s.args.dt = 1e-2;
        s.args.frames = 120;
        s.args.gravity = Vec3d{0, 0, 0};
        s.args.friction_mode = "max";
        return s;
    }
    if (name == "woven") { // examples/woven.ipynb; an overridden n keeps the vertex spacing along the strands (scale ~ n)
        const unsigned nn = n ? n : 256;
        double scale = 2.0 * double(nn) / 256.0;
        if (n && (unsigned(25.0 * 2.0 * 1.48 * scale) & 1u)) scale = (double(unsigned(25.0 * 2.0 * 1.48 * scale) + 1) + 1e-9) / (25.0 * 2.0 * 1.48); // even strand count
        return woven_scene(name, nn, 4e-3, scale, 1.5e-3, 2e-3, 0.01, 0.15, 10.0, 1.0, 1e-3, 10000000u, 450);
    }
    if (name == "woven-plastic") { // test scene: the woven cylinder with bend plasticity on the rods
        const unsigned nn = n ? n : 32;
        double scale = 2.0 * double(nn) / 256.0;
        if (unsigned(25.0 * 2.0 * 1.48 * scale) & 1u) scale = (double(unsigned(25.0 * 2.0 * 1.48 * scale) + 1) + 1e-9) / (25.0 * 2.0 * 1.48);
        SceneSpec s = woven_scene(name, nn, 4e-3, scale, 1.5e-3, 2e-3, 0.01, 0.15, 10.0, 1.0, 1e-3, 10000000u, 240);
        for (auto &o : s.objects) { o.rod_param["bend-plasticity"] = 30.0; o.rod_param["bend-plasticity-threshold"] = 1e-3; }
        return s;
    }
    if (name == "woven-damped") { // test scene: the woven cylinder with Rayleigh damping on the rod stretch and bending
        const unsigned nn = n ? n : 32;
        double scale = 2.0 * double(nn) / 256.0;
        if (unsigned(25.0 * 2.0 * 1.48 * scale) & 1u) scale = (double(unsigned(25.0 * 2.0 * 1.48 * scale) + 1) + 1e-9) / (25.0 * 2.0 * 1.48);
        SceneSpec s = woven_scene(name, nn, 4e-3, scale, 1.5e-3, 2e-3, 0.01, 0.15, 10.0, 1.0, 1e-3, 10000000u, 240);
        for (auto &o : s.objects) { o.rod_param["deformation-damping"] = 0.01; o.rod_param["bending-damping"] = 0.05; }
        return s;
    }
    if (name == "woven-sl") { // test scene: the woven cylinder stretched by 1.6 % along its axis within 0.5 s (slow spin), rods with a 2 % strain limit
        const unsigned nn = n ? n : 32;
        double scale = 2.0 * double(nn) / 256.0;
        if (unsigned(25.0 * 2.0 * 1.48 * scale) & 1u) scale = (double(unsigned(25.0 * 2.0 * 1.48 * scale) + 1) + 1e-9) / (25.0 * 2.0 * 1.48);
        return woven_scene(name, nn, 4e-3, scale, 1.5e-3, 2e-3, 0.01, -0.015, 0.5, 1.0, 1e-3, 10000000u, 240, 0.02, 0.0);
    }
    if (name == "large-woven") { // examples/large-woven.ipynb
        const unsigned nn = n ? n : 1024;
        double scale = 10.0 * double(nn) / 1024.0;
        if (n && (unsigned(25.0 * 2.0 * 1.48 * scale) & 1u)) scale = (double(unsigned(25.0 * 2.0 * 1.48 * scale) + 1) + 1e-9) / (25.0 * 2.0 * 1.48);
        return woven_scene(name, nn, 0.0015, scale, 0.125e-3, 0.0015 / 2.0, 0.001, 0.3, 20.0, 1.0 - 20.0 * 0.01, 1e-5, 180000000u, 1000);
    }
    if (name == "hang") { // examples/hang.ipynb
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_square(n ? n : 128, 2.0, Vec3d{1, 0, 0}, Vec3d{0, 1, 0}, o.vert, o.tri);
        o.param = tri_defaults();
        o.param["strain-limit"] = 0.01;
        o.param["young-mod"] = 100.0;
        o.param["bend"] = 1.0;
        Pin p;
        auto a = grab(o.vert, Vec3d{-1, 1, 0}), b = grab(o.vert, Vec3d{1, 1, 0});
        p.index = a;
        p.index.insert(p.index.end(), b.begin(), b.end());
        o.pins.push_back(p);
        s.objects.push_back(o);
        s.args.frames = 200;
        const double jitter = 0.35;
        s.args.gravity_y_keys = {{0.0, -9.8}, {1.0, -9.8}, {1.0 + jitter, 9.8}, {2.0, -9.8}};
        return s;
    }
    if (name == "headless-cw-open" || name == "headless-cw-closed") { // `headless` with a collision window on the moving sphere:
        // open = [0, 100) (the same run as `headless`), closed = [0, 0.3): from t = 0.3 on the sphere passes through the curtains
        SceneSpec s = make_scene("headless", n);
        s.name = name;
        s.objects.back().collision_windows = {{0.0, name == "headless-cw-open" ? 100.0 : 0.3}};
        return s;
    }
    if (name == "headless") { // examples/headless.py (the CI smoke scene): five curtains pinned along their top edge, a fully pinned
        // sphere mesh pushed through them (.jitter() left out; UV-less rest frames, like `hang`)
        SceneSpec s;
        s.name = name;
        const unsigned res = n ? n : 64;
        const unsigned subdiv = (n == 0 || n >= 64) ? 4u : (n >= 32 ? 3u : 2u);
        for (int i = 0; i < 5; ++i) {
            ShellObject o;
            gen_square(res, 2.0, Vec3d{0, 0, 1}, Vec3d{0, 1, 0}, o.vert, o.tri);
            o.param = tri_defaults();
            o.param["strain-limit"] = 0.05;
            o.displacement = Vec3d{i * 0.25, 0, 0};
            Pin p;
            p.index = grab(o.vert, Vec3d{0, 1, 0});
            o.pins.push_back(p);
            s.objects.push_back(std::move(o));
        }
        ShellObject sp;
        gen_icosphere(0.5, subdiv, sp.vert, sp.tri);
        sp.param = tri_defaults();
        sp.displacement = Vec3d{-1, 0, 0};
        Pin p;
        for (unsigned k = 0; k < sp.vert.size(); ++k) p.index.push_back(k);
        PinOp mv; mv.kind = PinOp::MoveBy; mv.delta = Vec3d{8, 0, 0}; mv.t_start = 0.0; mv.t_end = 5.0;
        p.ops = {mv};
        sp.pins.push_back(p);
        s.objects.push_back(std::move(sp));
        s.args.dt = 0.01;
        s.args.min_newton_steps = 8;
        s.args.frames = 60;
        return s;
    }
    if (name == "drape") { // test scene: a sheet thrown onto an invisible sphere standing on an invisible floor (analytic colliders with friction)
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_square(n ? n : 32, 1.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
        o.param = tri_defaults();
        o.param["friction"] = 0.3;
        o.displacement = Vec3d{0.03, 0.45, 0.02};
        o.velocity = Vec3d{0, -3.0, 0};
        s.objects.push_back(o);
        SphereCollider sp; sp.center = Vec3d{0, 0, 0}; sp.radius = 0.4; sp.friction = 0.3;
        WallCollider fl; fl.position = Vec3d{0, -0.08, 0}; fl.normal = Vec3d{0, 1, 0}; fl.friction = 0.3;
        s.spheres.push_back(sp);
        s.walls.push_back(fl);
        s.args.frames = 120;
        s.args.dt = 1e-2;
        s.args.friction_mode = "min";
        return s;
    }
    if (name == "balloon") { // test scene: a closed icosphere shell in zero gravity inflated by FaceParam.pressure
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_icosphere(0.3, (n == 0 || n >= 24) ? 3u : 2u, o.vert, o.tri);
        o.param = tri_defaults();
        o.param["pressure"] = 100.0;
        o.displacement = Vec3d{0.4, 0.5, 0.3}; // away from the origin: the per-face terms are translation-variant, their sum is not
        s.objects.push_back(o);
        s.args.gravity = Vec3d{0, 0, 0};
        s.args.frames = 60;
        s.args.dt = 1e-2;
        return s;
    }
    if (name == "sheet-torque") { // test scene: a free sheet in zero gravity spun up by a torque pin on its middle strip
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_square(n ? n : 24, 1.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
        o.param = tri_defaults();
        Pin p;
        for (unsigned k = 0; k < o.vert.size(); ++k)
            if (std::fabs(o.vert[k][0]) < 0.3) p.index.push_back(k); // the strip |x| < 0.3: its longest spread is along z
        PinOp t; t.kind = PinOp::Torque; t.axis_component = 0; t.magnitude = 0.05; t.hint_vertex = 0; t.t_start = 0.0; t.t_end = 0.2;
        p.ops = {t};
        o.pins.push_back(p);
        s.objects.push_back(o);
        s.args.gravity = Vec3d{0, 0, 0};
        s.args.frames = 60;
        s.args.dt = 1e-2;
        return s;
    }
    if (name == "drape-mesh") { // test scene: `drape` with the colliders as STATIC TRIANGLE MESHES (Constraint.mesh): an icosphere
        // standing on a ground square, both with friction; the sheet is thrown onto them
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_square(n ? n : 32, 1.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
        o.param = tri_defaults();
        o.param["friction"] = 0.3;
        o.displacement = Vec3d{0.03, 0.45, 0.02};
        o.velocity = Vec3d{0, -3.0, 0};
        s.objects.push_back(o);
        StaticObject ball;
        gen_icosphere(0.4, (n == 0 || n >= 32) ? 3u : 2u, ball.vert, ball.tri);
        ball.friction = 0.3;
        s.statics.push_back(ball);
        StaticObject ground;
        gen_square(8, 3.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, ground.vert, ground.tri);
        ground.displacement = Vec3d{0, -0.42, 0};
        ground.friction = 0.3;
        s.statics.push_back(ground);
        s.args.frames = 120;
        s.args.dt = 1e-2;
        s.args.friction_mode = "min";
        return s;
    }
    if (name == "seam") { // test scene: two sheets side by side joined by a row of stitches; the first hangs from two corners, the second from the seam
        SceneSpec s;
        s.name = name;
        const unsigned res = n ? n : 16;
        for (int k = 0; k < 2; ++k) {
            ShellObject o;
            gen_square(res, 1.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
            o.param = tri_defaults();
            o.displacement = Vec3d{k == 0 ? -0.51 : 0.51, 0, 0};
            if (k == 0) {
                Pin p;
                auto a = grab(o.vert, Vec3d{-1, 0, -1}), c = grab(o.vert, Vec3d{-1, 0, 1});
                p.index = a;
                p.index.insert(p.index.end(), c.begin(), c.end());
                o.pins.push_back(p);
            }
            s.objects.push_back(std::move(o));
        }
        for (unsigned j = 0; j < res; ++j) { // gen_square: vertex (i, j) has index i * res + j, x grows with i
            StitchSpec st;
            st.src = 0; st.dst = 1;
            const unsigned a = (res - 1) * res + j, c = j;
            st.src_index = {a, a, a};
            if (j + 1 < res && (j % 2)) { st.dst_index = {c, c + 1, c + 1}; st.dst_weight = {0.5, 0.5, 0.0}; } // every other stitch lands on an edge midpoint
            else st.dst_index = {c, c, c};
            s.stitches.push_back(st);
        }
        s.args.frames = 60;
        s.args.dt = 1e-2;
        return s;
    }
    if (name == "lift") { // test scene: a sheet pinned at two corners, lifted by an invisible sphere that rises from below (kinematic collider)
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_square(n ? n : 32, 2.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
        o.param = tri_defaults();
        Pin p;
        auto a = grab(o.vert, Vec3d{-1, 0, -1}), c = grab(o.vert, Vec3d{1, 0, -1});
        p.index = a;
        p.index.insert(p.index.end(), c.begin(), c.end());
        o.pins.push_back(p);
        s.objects.push_back(o);
        SphereCollider sp;
        sp.center = Vec3d{0.1, -0.9, 0.2};
        sp.radius = 0.5;
        sp.keys.push_back({0.6, Vec3d{0.1, -0.2, 0.2}, 0.5});
        s.spheres.push_back(sp);
        s.args.frames = 90;
        s.args.dt = 1e-2;
        return s;
    }
    if (name == "sheet" || name == "sheet-arap" || name == "sheet-stvk" || name == "sheet-snhk" || name == "sheet-damped" || name == "sheet-arap-damped" ||
        name == "sheet-plastic" || name == "sheet-fixxz") { // free-hanging sheet without strain limit: elastic-only smoke scene (the suffix selects the shell model)
        SceneSpec s;
        s.name = name;
        ShellObject o;
        gen_square(n ? n : 32, 2.0, Vec3d{1, 0, 0}, Vec3d{0, 0, 1}, o.vert, o.tri);
        o.param = tri_defaults();
        if (name == "sheet-arap" || name == "sheet-arap-damped") o.param["model"] = 0.0;
        if (name == "sheet-plastic") { // membrane and bending rest shapes creep toward the deformed shape
            o.param["plasticity"] = 20.0; o.param["plasticity-threshold"] = 1e-3;
            o.param["bend-plasticity"] = 20.0; o.param["bend-plasticity-threshold"] = 1e-2;
        }
        if (name == "sheet-damped" || name == "sheet-arap-damped") { o.param["deformation-damping"] = 0.02; o.param["bending-damping"] = 0.05; } // Rayleigh damping (seconds)
        if (name == "sheet-stvk") o.param["model"] = 1.0;
        if (name == "sheet-snhk") o.param["model"] = 3.0;
        Pin p;
        auto a = grab(o.vert, Vec3d{-1, 0, -1}), b = grab(o.vert, Vec3d{1, 0, -1});
        p.index = a;
        p.index.insert(p.index.end(), b.begin(), b.end());
        o.pins.push_back(p);
        s.objects.push_back(o);
        s.args.frames = 60;
        s.args.dt = 1e-2;
        if (name == "sheet-fixxz") s.args.fix_xz = -0.4; // horizontal drag on everything higher than y = -0.4 (ParamSet.fix_xz)
        return s;
    }
    throw std::runtime_error("unknown scene: " + name);
}

// ---- pin operations: scene.rs:2747-2870 + pin_apply.rs:27-160 (f64 math, f32 in/out) ----
static double progress_at(double time, double t0, double t1, bool smooth) {
    const double den = t1 - t0;
    const double raw = den != 0.0 ? (time - t0) / den : 0.0;
    return smooth ? raw * raw * (3.0 - 2.0 * raw) : raw;
}
static bool apply_pin_op(const PinOp &op, Vec3f &pos, double time) {
    const double p[3] = {pos[0], pos[1], pos[2]};
    double r[3] = {0, 0, 0};
    switch (op.kind) {
    case PinOp::MoveBy: {
        if (time < op.t_start) return false;
        const double prog = time >= op.t_end ? 1.0 : progress_at(time, op.t_start, op.t_end, op.smooth);
        // delta is stored as f32 per vertex in the reference (scene.rs:2769-2774)
        for (int k = 0; k < 3; ++k) r[k] = p[k] + double(float(op.delta[k])) * prog;
        break;
    }
    case PinOp::Spin: {
        const double t_eff = std::min(time, op.t_end) - op.t_start;
        const double angle = double(float(op.angular_velocity)) * (M_PI / 180.0) * t_eff;
        if (angle <= 0.0) return false;
        const double c[3] = {double(float(op.center[0])), double(float(op.center[1])), double(float(op.center[2]))};
        const double ax0[3] = {double(float(op.axis[0])), double(float(op.axis[1])), double(float(op.axis[2]))};
        const double an = std::sqrt(ax0[0] * ax0[0] + ax0[1] * ax0[1] + ax0[2] * ax0[2]);
        const double inv = an > 0.0 ? 1.0 / an : 0.0;
        const double ax[3] = {ax0[0] * inv, ax0[1] * inv, ax0[2] * inv};
        const double ct = std::cos(angle), st = std::sin(angle), omc = 1.0 - ct;
        const double px = p[0] - c[0], py = p[1] - c[1], pz = p[2] - c[2];
        const double cx = ax[1] * pz - ax[2] * py, cy = ax[2] * px - ax[0] * pz, cz = ax[0] * py - ax[1] * px;
        const double dp = px * ax[0] + py * ax[1] + pz * ax[2];
        r[0] = px * ct + cx * st + ax[0] * dp * omc + c[0];
        r[1] = py * ct + cy * st + ax[1] * dp * omc + c[1];
        r[2] = pz * ct + cz * st + ax[2] * dp * omc + c[2];
        break;
    }
    case PinOp::Torque: return false; // a force, not a motion (pin.rs:131-134)
    case PinOp::Scale: {
        if (time < op.t_start) return false;
        double cur;
        const double target = double(float(op.factor));
        if (time >= op.t_end) cur = target;
        else cur = 1.0 + (target - 1.0) * progress_at(time, op.t_start, op.t_end, op.smooth);
        for (int k = 0; k < 3; ++k) {
            const double c = double(float(op.center[k]));
            r[k] = (p[k] - c) * cur + c;
        }
        break;
    }
    }
    pos = Vec3f{float(r[0]), float(r[1]), float(r[2])};
    return true;
}

// scene.rs:1944-2130 (fix / pull pins only; collision mesh, spheres, floors, torque are out of scope)
void make_constraint(Built &b, double time, double prev_time) {
    b.fix.clear();
    b.pull.clear();
    b.torque_groups.clear();
    b.torque_vertices.clear();
    for (auto &pin : b.pins) { // torque operations: one group per pin (scene.rs:2132-2198), full power inside the window
        if (pin.unpin_time >= 0.0 && time >= pin.unpin_time) continue;
        for (auto &op : pin.ops) {
            if (op.kind != PinOp::Torque || time < op.t_start || time > op.t_end) continue;
            PpfTorqueGroup g{};
            g.axis_component = op.axis_component;
            g.vertex_start = (unsigned)b.torque_vertices.size();
            g.vertex_count = (unsigned)pin.index.size();
            g.hint_vertex = pin.index.at(op.hint_vertex);
            for (unsigned ind : pin.index) b.torque_vertices.push_back(PpfTorqueVertex{float(op.magnitude), ind, (unsigned)b.torque_groups.size()});
            b.torque_groups.push_back(g);
            break; // one torque per pin group
        }
    }
    for (auto &pin : b.pins) {
        if (pin.unpin_time >= 0.0 && time >= pin.unpin_time) continue;
        bool torque_only = !pin.ops.empty(); // scene.rs:2061-2070: such vertices stay free
        for (auto &op : pin.ops) torque_only = torque_only && op.kind == PinOp::Torque;
        if (torque_only) continue;
        const Vec3d &d = b.spec.objects[pin.object].displacement;
        const Vec3f dx{float(d[0]), float(d[1]), float(d[2])};
        for (size_t i = 0; i < pin.index.size(); ++i) {
            bool kinematic = false;
            Vec3f position = pin.rest[i];
            for (auto &op : pin.ops) kinematic = apply_pin_op(op, position, time) || kinematic;
            Vec3f delta{0, 0, 0};
            if (kinematic) {
                Vec3f start = pin.rest[i];
                for (auto &op : pin.ops) apply_pin_op(op, start, prev_time);
                delta = Vec3f{position[0] - start[0], position[1] - start[1], position[2] - start[2]};
            }
            if (pin.pull_w > 0.0) {
                PpfPullPair pp{};
                for (int k = 0; k < 3; ++k) pp.position.v[k] = position[k] + dx[k];
                pp.index = pin.index[i];
                pp.weight = float(pin.pull_w);
                b.pull.push_back(pp);
            } else {
                PpfFixPair fp{};
                for (int k = 0; k < 3; ++k) {
                    fp.position.v[k] = position[k] + dx[k];
                    fp.step_delta.v[k] = delta[k];
                }
                fp.ghat = float(b.spec.args.constraint_ghat);
                fp.index = pin.index[i];
                fp.kinematic = kinematic;
                b.fix.push_back(fp);
            }
        }
    }
    // invisible walls and spheres (scene.rs:2213-2254); world scaling applies to positions, radii and thickness
    const float ws = float(b.spec.args.world_scaling);
    b.floor.clear();
    b.sphere.clear();
    // calc_coefficient (scene.rs:2014-2038): keyframe pair (j, k) and blend weight at `time`; timings[0] = 0
    auto blend = [&](const std::vector<double> &timings, bool smooth, size_t &j, size_t &k, float &w) {
        const size_t n = timings.size();
        if (time >= timings[n - 1]) { j = k = n - 1; w = 1.0f; return; }
        for (size_t i = 0; i + 1 < n; ++i)
            if (time >= timings[i] && time < timings[i + 1]) {
                double ww = (time - timings[i]) / (timings[i + 1] - timings[i]);
                if (smooth) ww = ww * ww * (3.0 - 2.0 * ww);
                j = i; k = i + 1; w = float(ww);
                return;
            }
        j = k = 0; w = 0.0f;
    };
    for (auto &w : b.spec.walls) {
        PpfFloor f{};
        std::vector<double> timings{0.0};
        std::vector<Vec3f> pos{Vec3f{float(w.position[0]) * ws, float(w.position[1]) * ws, float(w.position[2]) * ws}};
        for (auto &kf : w.keys) { timings.push_back(kf.first); pos.push_back(Vec3f{float(kf.second[0]) * ws, float(kf.second[1]) * ws, float(kf.second[2]) * ws}); }
        size_t j = 0, k = 0;
        float wt = 0.0f;
        const bool kinematic = timings.size() > 1;
        if (kinematic) blend(timings, w.smooth, j, k, wt);
        for (int c = 0; c < 3; ++c) { f.ground.v[c] = pos[j][c] * (1.0f - wt) + pos[k][c] * wt; f.up.v[c] = float(w.normal[c]); }
        f.ghat = float(w.contact_gap); f.friction = float(w.friction); f.thickness = float(w.thickness) * ws; f.kinematic = kinematic;
        b.floor.push_back(f);
    }
    for (auto &sp : b.spec.spheres) {
        PpfSphere q{};
        std::vector<double> timings{0.0};
        std::vector<Vec3f> pos{Vec3f{float(sp.center[0]) * ws, float(sp.center[1]) * ws, float(sp.center[2]) * ws}};
        std::vector<float> rad{float(sp.radius) * ws};
        for (auto &kf : sp.keys) {
            timings.push_back(kf.time);
            pos.push_back(Vec3f{float(kf.center[0]) * ws, float(kf.center[1]) * ws, float(kf.center[2]) * ws});
            rad.push_back(float(kf.radius) * ws);
        }
        size_t j = 0, k = 0;
        float wt = 0.0f;
        const bool kinematic = timings.size() > 1;
        if (kinematic) blend(timings, sp.smooth, j, k, wt);
        for (int c = 0; c < 3; ++c) q.center.v[c] = pos[j][c] * (1.0f - wt) + pos[k][c] * wt;
        q.radius = rad[j] * (1.0f - wt) + rad[k] * wt;
        q.ghat = float(sp.contact_gap); q.friction = float(sp.friction); q.thickness = float(sp.thickness) * ws;
        q.bowl = sp.hemisphere; q.reverse = sp.inverted; q.kinematic = kinematic;
        b.sphere.push_back(q);
    }
    PpfConstraint &c = b.constraint;
    c = PpfConstraint{};
    c.mesh = b.cmesh;
    c.stitch.data = b.stitch.data();
    c.stitch.size = c.stitch.allocated = (unsigned)b.stitch.size();
    c.floor.data = b.floor.data();
    c.floor.size = c.floor.allocated = (unsigned)b.floor.size();
    c.sphere.data = b.sphere.data();
    c.sphere.size = c.sphere.allocated = (unsigned)b.sphere.size();
    c.fix.data = b.fix.data();
    c.fix.size = c.fix.allocated = (unsigned)b.fix.size();
    c.pull.data = b.pull.data();
    c.pull.size = c.pull.allocated = (unsigned)b.pull.size();
    c.torque_groups.data = b.torque_groups.data();
    c.torque_groups.size = c.torque_groups.allocated = (unsigned)b.torque_groups.size();
    c.torque_vertices.data = b.torque_vertices.data();
    c.torque_vertices.size = c.torque_vertices.allocated = (unsigned)b.torque_vertices.size();
}

// scene.rs update_param: time-varying session params (only gravity keyframes are needed here)
void update_param(Built &b, double time) {
    const auto &keys = b.spec.args.gravity_y_keys;
    if (keys.empty()) return;
    double g = keys.back().second;
    if (time <= keys.front().first) g = keys.front().second;
    else
        for (size_t i = 0; i + 1 < keys.size(); ++i)
            if (time >= keys[i].first && time < keys[i + 1].first) {
                const double w = (time - keys[i].first) / (keys[i + 1].first - keys[i].first);
                g = keys[i].second * (1.0 - w) + keys[i + 1].second * w;
                break;
            }
    b.param.gravity.v[1] = float(g);
}

} // namespace ppfh

// backend.cpp -- dlopen binding of the 20-symbol C ABI and the per-step run loop.
// Restates crates/ppf-cts-solver/src/backend.rs:28-62 (extern block), 626-1025 (run) and
// 169-178 (State) for the calls the hot path needs.
#include "ppf_host.hpp"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <stdexcept>

namespace ppfh {

bool Backend::open(const std::string &lib_path, std::string &err) {
    path = lib_path;
    // RTLD_DEEPBIND keeps each backend's own globals (g_ppf_fatal_code, print_rust fallback ...)
    // bound to itself when two backends live in one process.
    handle = dlopen(lib_path.c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
    if (!handle) {
        err = dlerror();
        return false;
    }
#define PPFH_SYM(name)                                                                             \
    name = reinterpret_cast<decltype(name)>(dlsym(handle, #name));                                 \
    if (!name) {                                                                                   \
        err = std::string("missing symbol ") + #name + " in " + lib_path;                          \
        return false;                                                                              \
    }
    PPFH_SYM(initialize) PPFH_SYM(advance) PPFH_SYM(fetch) PPFH_SYM(fetch_inv_rest) PPFH_SYM(fetch_rest_angles)
    PPFH_SYM(fetch_dyn_counts) PPFH_SYM(fetch_dyn) PPFH_SYM(update_dyn) PPFH_SYM(update_constraint)
    PPFH_SYM(update_rest_shape) PPFH_SYM(override_velocity) PPFH_SYM(gather_current_positions)
    PPFH_SYM(override_angular_velocity) PPFH_SYM(init_collision_windows) PPFH_SYM(refresh_collision_active)
    PPFH_SYM(fetch_intersection_records) PPFH_SYM(set_log_path) PPFH_SYM(ppf_fatal_code) PPFH_SYM(ppf_fatal_detail)
#undef PPFH_SYM
    b200_get_stats = reinterpret_cast<decltype(b200_get_stats)>(dlsym(handle, "ppf_b200_get_stats"));
    b200_reset_stats = reinterpret_cast<decltype(b200_reset_stats)>(dlsym(handle, "ppf_b200_reset_stats"));
    b200_fetch_pairs = reinterpret_cast<decltype(b200_fetch_pairs)>(dlsym(handle, "ppf_b200_fetch_pairs"));
    return true;
}

bool Runner::init(const std::string &log_dir) {
    be->set_log_path(log_dir.c_str()); // main.rs:44
    time = b->param.time;
    curr_frame = int(std::floor(time * b->spec.args.fps));
    if (!be->initialize(&b->dataset, &b->param)) return false; // backend.rs:653
    // backend.rs:654-672 + scene.rs:2471-2542: one group per object, up to 8 windows each
    const unsigned n_groups = (unsigned)b->spec.objects.size();
    std::vector<float> windows(size_t(n_groups) * 16, 0.0f);
    std::vector<unsigned> counts(n_groups, 0u);
    bool any = false;
    for (unsigned o = 0; o < n_groups; ++o) {
        const auto &w = b->spec.objects[o].collision_windows;
        if (w.size() > 8) throw std::runtime_error("an object has more than 8 collision windows");
        counts[o] = (unsigned)w.size();
        for (size_t k = 0; k < w.size(); ++k) { windows[size_t(o) * 16 + 2 * k] = float(w[k].first); windows[size_t(o) * 16 + 2 * k + 1] = float(w[k].second); }
        any |= !w.empty();
    }
    if (any) be->init_collision_windows(b->vert_dmap.data(), (unsigned)b->vert_dmap.size(), windows.data(), counts.data(), n_groups);
    return true;
}

bool Runner::step() {
    // backend.rs:837-935
    update_param(*b, time);
    const double target_time = time + double(b->param.dt);
    make_constraint(*b, target_time, time);
    be->update_constraint(&b->constraint);
    be->refresh_collision_active(float(time));
    StepRecord rec;
    std::memset(&rec.result, 0, sizeof(rec.result));
    const auto t0 = std::chrono::steady_clock::now();
    be->advance(&rec.result);
    const auto t1 = std::chrono::steady_clock::now();
    rec.wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    const PpfStepResult &r = rec.result;
    const bool ok = r.ccd_success && r.pcg_success && r.intersection_free && r.newton_progress && r.pin_feasible &&
                    r.contact_separated;
    if (ok) {
        time = r.time;
        const int new_frame = int(std::floor(time * b->spec.args.fps));
        if (new_frame > curr_frame) curr_frame = new_frame;
    }
    rec.time = time;
    steps.push_back(rec);
    return ok;
}

void Runner::fetch_positions(std::vector<float> &curr, std::vector<float> &prev) {
    be->fetch(); // writes into b->curr / b->prev (the host arrays the DataSet points at)
    curr.assign(reinterpret_cast<float *>(b->curr.data()), reinterpret_cast<float *>(b->curr.data()) + 3 * b->n_vert);
    prev.assign(reinterpret_cast<float *>(b->prev.data()), reinterpret_cast<float *>(b->prev.data()) + 3 * b->n_vert);
}

void Runner::fetch_dyn_pattern(std::vector<unsigned> &index, std::vector<unsigned> &offset) {
    unsigned n_value = 0, n_offset = 0; // backend.rs:200-238
    be->fetch_dyn_counts(&n_value, &n_offset);
    index.assign(n_value, 0);
    offset.assign(n_offset, 0);
    be->fetch_dyn(index.data(), nullptr, offset.data());
}

static const unsigned STATE_MAGIC = 0x54535050u; // "PPST"
bool save_state(const std::string &path, const Runner &r, const std::vector<float> &curr, const std::vector<float> &prev,
                const std::vector<unsigned> &dyn_index, const std::vector<unsigned> &dyn_offset, float prev_dt) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    unsigned hdr[4] = {STATE_MAGIC, (unsigned)(curr.size() / 3), (unsigned)dyn_index.size(), (unsigned)dyn_offset.size()};
    fwrite(hdr, sizeof(unsigned), 4, f);
    fwrite(&r.time, sizeof(double), 1, f);
    fwrite(&prev_dt, sizeof(float), 1, f);
    fwrite(curr.data(), sizeof(float), curr.size(), f);
    fwrite(prev.data(), sizeof(float), prev.size(), f);
    fwrite(dyn_index.data(), sizeof(unsigned), dyn_index.size(), f);
    fwrite(dyn_offset.data(), sizeof(unsigned), dyn_offset.size(), f);
    fclose(f);
    return true;
}
bool load_state(const std::string &path, double &time, float &prev_dt, std::vector<float> &curr, std::vector<float> &prev,
                std::vector<unsigned> &dyn_index, std::vector<unsigned> &dyn_offset) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned hdr[4];
    bool ok = fread(hdr, sizeof(unsigned), 4, f) == 4 && hdr[0] == STATE_MAGIC;
    if (ok) {
        curr.resize(size_t(3) * hdr[1]); prev.resize(size_t(3) * hdr[1]);
        dyn_index.resize(hdr[2]); dyn_offset.resize(hdr[3]);
        ok = fread(&time, sizeof(double), 1, f) == 1 && fread(&prev_dt, sizeof(float), 1, f) == 1 &&
             fread(curr.data(), sizeof(float), curr.size(), f) == curr.size() &&
             fread(prev.data(), sizeof(float), prev.size(), f) == prev.size() &&
             fread(dyn_index.data(), sizeof(unsigned), dyn_index.size(), f) == dyn_index.size() &&
             fread(dyn_offset.data(), sizeof(unsigned), dyn_offset.size(), f) == dyn_offset.size();
    }
    fclose(f);
    return ok;
}

} // namespace ppfh

// main.cpp -- `ppf_host`: command-line stand-in for the reference's `ppf-contact-solver`
// binary (crates/ppf-cts-solver/src/main.rs:55-299) for synthetic scenes.  It builds the
// DataSet for a named scene, loads ONE backend shared library through the C ABI, runs K
// steps and writes positions / dynamic-CSR pattern / per-step records to --out.
#include "ppf_host.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>

// The reference backend imports this symbol from its host (main.rs:47).
extern "C" void print_rust(const char *msg) {
    if (getenv("PPF_HOST_QUIET") == nullptr) fprintf(stderr, "[backend] %s\n", msg);
}

using namespace ppfh;

static void write_bin(const std::string &path, const void *data, size_t bytes) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); return; }
    fwrite(data, 1, bytes, f);
    fclose(f);
}

int main(int argc, char **argv) {
    std::string lib, scene = "fast-twist", out = "gpurun_out/run", load;
    unsigned n = 0, steps = 10, save_every = 0, pairs_every = 0;
    bool quiet_steps = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--lib") lib = next();
        else if (a == "--scene") scene = next();
        else if (a == "--n") n = (unsigned)atoi(next());
        else if (a == "--steps") steps = (unsigned)atoi(next());
        else if (a == "--out") out = next();
        else if (a == "--load") load = next();
        else if (a == "--save-every") save_every = (unsigned)atoi(next());
        else if (a == "--pairs-every") pairs_every = (unsigned)atoi(next());
        else if (a == "--quiet") quiet_steps = true;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (lib.empty()) { fprintf(stderr, "usage: ppf_host --lib <libsimbackend_cuda.so> --scene <name> [--n N] [--steps K] [--out DIR] [--load STATE] [--save-every M]\n"); return 2; }
    mkdir(out.c_str(), 0755);
    const std::string data_dir = out + "/data";

    Built b;
    build_scene(make_scene(scene, n), b);
    fprintf(stderr, "[host] scene %s: V=%u F=%zu E=%zu H=%zu fix=%zu nnz_fixed=%zu dt=%g\n", scene.c_str(), b.n_vert,
            b.face.size(), b.edge.size(), b.hinge.size(), b.fix.size(), b.f_fixed.data.size(), b.param.dt);

    std::vector<unsigned> dyn_index, dyn_offset;
    bool loaded = false;
    if (!load.empty()) {
        double t; float prev_dt;
        std::vector<float> c, p;
        if (!load_state(load, t, prev_dt, c, p, dyn_index, dyn_offset) || c.size() != size_t(3) * b.n_vert) {
            fprintf(stderr, "cannot load state %s\n", load.c_str());
            return 2;
        }
        memcpy(b.curr.data(), c.data(), c.size() * sizeof(float));
        memcpy(b.prev.data(), p.data(), p.size() * sizeof(float));
        b.param.time = t;
        b.param.prev_dt = prev_dt;
        loaded = true;
    }

    Backend be;
    std::string err;
    if (!be.open(lib, err)) { fprintf(stderr, "backend: %s\n", err.c_str()); return 3; }
    Runner r;
    r.b = &b;
    r.be = &be;
    if (!r.init(data_dir)) {
        fprintf(stderr, "initialize failed (fatal code %u: %s)\n", (unsigned)be.ppf_fatal_code(), be.ppf_fatal_detail());
        return 4;
    }
    if (loaded && !b.spec.args.disable_contact && dyn_offset.size() == size_t(b.n_vert) + 1)
        be.update_dyn(dyn_index.data(), dyn_offset.data()); // backend.rs:687-694

    FILE *fs = fopen((out + "/steps.jsonl").c_str(), "w");
    std::vector<float> curr, prev;
    bool ok = true;
    double wall_total = 0;
    unsigned done = 0;
    for (unsigned s = 0; s < steps && ok; ++s) {
        ok = r.step();
        const StepRecord &rec = r.steps.back();
        wall_total += rec.wall_ms;
        ++done;
        if (fs)
            fprintf(fs, "{\"step\": %u, \"time\": %.9g, \"wall_ms\": %.3f, \"ok\": %s, \"ccd\": %d, \"pcg\": %d, \"isect_free\": %d}\n", s,
                    rec.time, rec.wall_ms, ok ? "true" : "false", rec.result.ccd_success, rec.result.pcg_success,
                    rec.result.intersection_free);
        if (!quiet_steps) fprintf(stderr, "[host] step %u t=%.6f wall=%.2f ms ok=%d\n", s, rec.time, rec.wall_ms, (int)ok);
        if (ok && save_every && ((s + 1) % save_every == 0)) {
            r.fetch_positions(curr, prev);
            r.fetch_dyn_pattern(dyn_index, dyn_offset);
            char nm[64];
            snprintf(nm, sizeof(nm), "/state_%05u.bin", s + 1);
            save_state(out + nm, r, curr, prev, dyn_index, dyn_offset, b.param.prev_dt);
        }
        if (ok && pairs_every && be.b200_fetch_pairs && ((s + 1) % pairs_every == 0)) {
            for (unsigned k = 0; k < 4; ++k) {
                unsigned cnt = be.b200_fetch_pairs(k, nullptr, 0);
                std::vector<unsigned> buf(size_t(2) * cnt);
                be.b200_fetch_pairs(k, buf.data(), cnt);
                char nm[64];
                snprintf(nm, sizeof(nm), "/pairs_%05u_%u.bin", s + 1, k);
                write_bin(out + nm, buf.data(), buf.size() * sizeof(unsigned));
            }
        }
    }
    if (fs) fclose(fs);
    r.fetch_positions(curr, prev);
    r.fetch_dyn_pattern(dyn_index, dyn_offset);
    write_bin(out + "/final_curr.bin", curr.data(), curr.size() * sizeof(float));
    write_bin(out + "/final_prev.bin", prev.data(), prev.size() * sizeof(float));
    write_bin(out + "/dyn_index.bin", dyn_index.data(), dyn_index.size() * sizeof(unsigned));
    write_bin(out + "/dyn_offset.bin", dyn_offset.data(), dyn_offset.size() * sizeof(unsigned));
    save_state(out + "/state_final.bin", r, curr, prev, dyn_index, dyn_offset, b.param.prev_dt);

    FILE *fj = fopen((out + "/summary.json").c_str(), "w");
    if (fj) {
        fprintf(fj, "{\"scene\": \"%s\", \"lib\": \"%s\", \"n_vert\": %u, \"n_face\": %zu, \"n_edge\": %zu, \"n_hinge\": %zu, "
                    "\"steps\": %u, \"ok\": %s, \"sim_time\": %.9g, \"frames\": %d, \"wall_ms_total\": %.3f, \"dyn_nnz\": %zu",
                scene.c_str(), lib.c_str(), b.n_vert, b.face.size(), b.edge.size(), b.hinge.size(), done,
                ok ? "true" : "false", r.time, r.curr_frame, wall_total, dyn_index.size());
        if (be.b200_get_stats) {
            PpfB200Stats st;
            be.b200_get_stats(&st);
            fprintf(fj, ", \"b200\": {\"ms_advance\": %.3f, \"ms_lbvh_build\": %.3f, \"ms_assembly\": %.3f, \"ms_contact\": %.3f, "
                        "\"ms_linsolve\": %.3f, \"ms_spmv\": %.3f, \"ms_aabb_update\": %.3f, \"ms_line_search\": %.3f, "
                        "\"n_advance\": %llu, \"n_newton\": %llu, \"n_cg_iter\": %llu, \"n_spmv\": %llu, \"n_kernel_launches\": %llu, "
                        "\"spmv_n_row\": %llu, \"spmv_n_off\": %llu}",
                    st.ms_advance, st.ms_lbvh_build, st.ms_assembly, st.ms_contact, st.ms_linsolve, st.ms_spmv,
                    st.ms_aabb_update, st.ms_line_search, st.n_advance, st.n_newton, st.n_cg_iter, st.n_spmv,
                    st.n_kernel_launches, st.spmv_n_row, st.spmv_n_off);
        }
        fprintf(fj, "}\n");
        fclose(fj);
    }
    fprintf(stderr, "[host] done: %u steps, ok=%d, sim_time=%.6f, wall=%.1f ms\n", done, (int)ok, r.time, wall_total);
    return ok ? 0 : 5;
}

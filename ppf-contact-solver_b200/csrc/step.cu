// step.cu -- one simulation step: BVH build, Newton loop {assemble, Dirichlet, PCG, CCD line search},
// final intersection check.
//
// Same control flow and log channels as the reference's main_helper::advance()
// (crates/ppf-cts-solver/src/cpp/main/main.cu:267-1649; recipe in SURVEY.md A.1): target
// x* = x + (x - x_prev) dt/prev_dt + dt^2 g, Newton iterations until toi_advanced >= target_toi and
// step >= min_newton_steps, then one error-reduction iteration with dt *= toi_advanced and the
// kinematic pins rewound; toi_advanced accumulated in double on the host.  The difference is in
// how the work is issued: everything runs on one stream, and the host synchronises only where it
// must take a decision (contact counts for buffer sizing, PCG status, max|dx|, toi, final check).
#include "backend.hpp"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <vector>

namespace ppf {

__global__ void k_velocity(unsigned n, const float4 *curr, const float4 *prev, const unsigned *fix_index, float prev_dt, float *speed2) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = curr[i], b = prev[i];
    const float ux = (a.x - b.x) / prev_dt, uy = (a.y - b.y) / prev_dt, uz = (a.z - b.z) / prev_dt;
    speed2[i] = fix_index[i] > 0 ? 0.0f : ux * ux + uy * uy + uz * uz;
}

__global__ void k_target(unsigned n, const float4 *curr, const float4 *prev, const unsigned *fix_index, const PpfFixPair *fix, float dt,
                         float prev_dt, float3 gravity, int inactive_momentum, float4 *target) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned fi = fix_index[i];
    if (fi > 0) {
        const PpfFixPair &f = fix[fi - 1];
        target[i] = make_float4(f.position.v[0], f.position.v[1], f.position.v[2], 0.f);
    } else {
        const float4 x1 = curr[i], x0 = prev[i];
        const float tr = dt / prev_dt, h2 = dt * dt;
        const float yx = (x1.x - x0.x) * tr + h2 * gravity.x, yy = (x1.y - x0.y) * tr + h2 * gravity.y, yz = (x1.z - x0.z) * tr + h2 * gravity.z;
        target[i] = inactive_momentum ? x1 : make_float4(x1.x + yx, x1.y + yy, x1.z + yz, 0.f);
    }
}

__global__ void k_rewind_fix(unsigned n, PpfFixPair *fix, float back) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fix[i].kinematic)
        for (int k = 0; k < 3; ++k) fix[i].position.v[k] = fix[i].position.v[k] - back * fix[i].step_delta.v[k];
}

__global__ void k_seed_dx(unsigned n, const unsigned *fix_index, const float4 *eval_x, const float4 *target, float *dx) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f, b = 0.f, c = 0.f;
    if (fix_index[i] > 0) {
        const float4 e = eval_x[i], t = target[i];
        a = e.x - t.x; b = e.y - t.y; c = e.z - t.z;
    }
    dx[3 * size_t(i)] = a; dx[3 * size_t(i) + 1] = b; dx[3 * size_t(i) + 2] = c;
}

__global__ void k_dx_norm(unsigned n, const float *dx, float *out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = dx[3 * size_t(i)], b = dx[3 * size_t(i) + 1], c = dx[3 * size_t(i) + 2];
    out[i] = sqrtf(a * a + b * b + c * c);
}

__global__ void k_apply_dx(unsigned n, float4 *eval_x, float4 *x_old, const float *dx, float toi_recale) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 e = eval_x[i];
    x_old[i] = e;
    eval_x[i] = make_float4(e.x - (toi_recale * dx[3 * size_t(i)]), e.y - (toi_recale * dx[3 * size_t(i) + 1]), e.z - (toi_recale * dx[3 * size_t(i) + 2]), 0.f);
}

// ParamSet.fix_xz: above the height fix_xz the horizontal part of the Newton update is pulled back toward the PREVIOUS step's
// position (data.vertex.prev, main.cu:361,1190-1208), fully from one unit above; prescribed (DOF-removed) vertices keep their keyframe
__global__ void k_fix_xz(unsigned n, float4 *eval_x, const float4 *prev, const unsigned *fix_index, float fix_xz) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || fix_index[i] != 0) return;
    float4 e = eval_x[i];
    if (e.y > fix_xz) {
        const float y = fminf(1.0f, e.y - fix_xz);
        const float4 z = prev[i];
        e.x -= y * (e.x - z.x);
        e.z -= y * (e.z - z.z);
        eval_x[i] = e;
    }
}

__global__ void k_lerp(unsigned n, float4 *eval_x, const float4 *x_old, float toi) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 e = eval_x[i], t = x_old[i];
    const float dx = toi * (e.x - t.x), dy = toi * (e.y - t.y), dz = toi * (e.z - t.z);
    eval_x[i] = make_float4(t.x + dx, t.y + dy, t.z + dz, 0.f);
}

// Device time of the phases, with CUDA events on the backend stream.  Events are only RECORDED while the step is being issued
// (no host wait per phase: round 1 synchronised on every stop()); the elapsed times are read once, when the step has ended.
struct PhaseClock {
    cudaStream_t q;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Span { size_t a, b; double *dst; };
    std::vector<Span> spans;
    explicit PhaseClock(cudaStream_t s) : q(s) {}
    size_t mark() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        cudaEventRecord(pool[used], q);
        return used++;
    }
    void collect() { // after the stream has been drained
        for (const Span &sp : spans) { float ms = 0.f; if (cudaEventElapsedTime(&ms, pool[sp.a], pool[sp.b]) == cudaSuccess) *sp.dst += ms; }
        spans.clear();
        used = 0;
    }
};
static PhaseClock *g_clock = nullptr;
struct EvTimer { // one phase: start() ... stop(&accumulator)
    size_t a = 0;
    explicit EvTimer(cudaStream_t) {}
    void start() { a = g_clock->mark(); }
    void stop(double *dst) { g_clock->spans.push_back({a, g_clock->mark(), dst}); }
};

// PPF_DUMP_LINSYS=<k> / PPF_DUMP_DIR: one-shot dump of the k-th assembled Newton system in the
// reference's own format (main/dump_linsys.hpp:93-190: magic "LSYS", nrow, n_off, COO upper-triangle
// 3x3 blocks of A_dyn and B_fixed, merged diagonal blocks A_ii + B_ii + C_i, rhs) so assembled
// systems of the two backends can be compared offline.
static void maybe_dump_linsys(State &s) {
    static int target = [] { const char *e = getenv("PPF_DUMP_LINSYS"); return e ? atoi(e) : -1; }();
    static int counter = 0;
    if (target < 0 || counter++ != target) return;
    CUDA_OK(cudaStreamSynchronize(s.stream));
    const unsigned nv = s.nv;
    std::vector<unsigned> fo(nv + 1), fc(s.fixed_nnz), dof(nv + 1, 0), dc(s.dyn_nnz);
    std::vector<float> fv(size_t(9) * s.fixed_nnz), dv(size_t(9) * s.dyn_nnz), diag(size_t(9) * nv), rhs(size_t(3) * nv);
    CUDA_OK(cudaMemcpy(fo.data(), s.fx_off.p, 4 * (nv + 1), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(fc.data(), s.fx_col.p, 4 * size_t(s.fixed_nnz), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(fv.data(), s.fx_val.p, 36 * size_t(s.fixed_nnz), cudaMemcpyDeviceToHost));
    if (s.dyn_nnz) {
        CUDA_OK(cudaMemcpy(dof.data(), s.dyn_off.p, 4 * (nv + 1), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(dc.data(), s.dyn_col.p, 4 * size_t(s.dyn_nnz), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(dv.data(), s.dyn_val.p, 36 * size_t(s.dyn_nnz), cudaMemcpyDeviceToHost));
    }
    CUDA_OK(cudaMemcpy(diag.data(), s.diag.p, 36 * size_t(nv), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(rhs.data(), s.force.p, 12 * size_t(nv), cudaMemcpyDeviceToHost));
    const char *dir = getenv("PPF_DUMP_DIR");
    char path[1024];
    snprintf(path, sizeof(path), "%s/linsys_%d.bin", dir ? dir : ".", target);
    FILE *f = fopen(path, "wb");
    if (!f) return;
    unsigned n_off = 0;
    for (unsigned i = 0; i < nv; ++i) {
        for (unsigned k = dof[i]; k < dof[i + 1]; ++k) if (dc[k] > i) ++n_off;
        for (unsigned k = fo[i]; k < fo[i + 1]; ++k) if (fc[k] > i) ++n_off;
    }
    const unsigned magic = 0x4C535953u;
    fwrite(&magic, 4, 1, f); fwrite(&nv, 4, 1, f); fwrite(&n_off, 4, 1, f);
    for (unsigned i = 0; i < nv; ++i) {
        for (unsigned k = dof[i]; k < dof[i + 1]; ++k)
            if (dc[k] > i) { fwrite(&i, 4, 1, f); fwrite(&dc[k], 4, 1, f); fwrite(&dv[size_t(9) * k], 4, 9, f); }
            else if (dc[k] == i) for (int e = 0; e < 9; ++e) diag[size_t(9) * i + e] += dv[size_t(9) * k + e];
        for (unsigned k = fo[i]; k < fo[i + 1]; ++k)
            if (fc[k] > i) { fwrite(&i, 4, 1, f); fwrite(&fc[k], 4, 1, f); fwrite(&fv[size_t(9) * k], 4, 9, f); }
            else if (fc[k] == i) for (int e = 0; e < 9; ++e) diag[size_t(9) * i + e] += fv[size_t(9) * k + e];
    }
    fwrite(diag.data(), 4, diag.size(), f);
    fwrite(rhs.data(), 4, rhs.size(), f);
    fclose(f);
    fprintf(stderr, "[dump_linsys] wrote %s: nrow=%u n_offdiag=%u (solve #%d)\n", path, nv, n_off, target);
}

void step_advance(State &s, PpfStepResult *out) {
    ScopeLog logging("advance");
    cudaStream_t q = s.stream;
    PpfStepResult result;
    result.time = s.param->time;
    result.ccd_success = result.pcg_success = result.intersection_free = true;
    result.newton_progress = result.pin_feasible = result.contact_separated = true;
    PpfParamSet prm = *s.param;
    const unsigned nv = s.nv;
    const unsigned gv = div_up(nv, 256);
    simplelog_set_time(prm.time);
    if (!g_clock) g_clock = new PhaseClock(q);
    EvTimer t_all(q), t_ph(q);
    t_all.start();
    // The sweep's candidate lists were filtered with the collision-window flags of the step that produced them: when windows
    // exist (refresh_collision_active runs before every advance) the first assembly of a step traverses afresh.
    if (s.has_cw) s.cand_cover_eval = false;
    auto finish = [&](bool ok) {
        if (!ok) s.cand_cover_eval = false; // a failed advance leaves eval_x to be reset: its sweep candidates cover nothing
        t_all.stop(&s.tm.ms_advance);
        CUDA_OK(cudaStreamSynchronize(q));
        g_clock->collect();
        s.tm.n_advance += 1;
        *out = result;
    };

    if (!prm.disable_contact) {
        logging.push("lbvh build");
        t_ph.start();
        bvh_build(s, s.bvh_face, 0, s.curr.p);
        bvh_build(s, s.bvh_edge, 1, s.curr.p);
        bvh_build(s, s.bvh_vert, 2, s.curr.p);
        t_ph.stop(&s.tm.ms_lbvh_build);
        logging.pop();
    }
    const float prev_dt = prm.prev_dt;
    {
        float *tmp = reinterpret_cast<float *>(s.velocity.p);
        k_velocity<<<gv, 256, 0, q>>>(nv, s.curr.p, s.prev.p, s.fix_index.p, prev_dt, tmp);
        LAUNCH_COUNT(1);
        logging.mark("max_u", sqrtf(reduce_max_f(s, tmp, nv, 0.0f)));
    }
    float dt = s.param->dt * s.param->playback;
    logging.mark("dt", dt);
    logging.mark("playback", s.param->playback);
    if (s.nf || s.nrod) logging.mark("max_sigma", max_stretch_ratio(s)); // main.cu:424-488
    const unsigned long long alloc0 = g_dev_alloc, free0 = g_dev_free;

    const float3 gravity = make_float3(prm.gravity.v[0], prm.gravity.v[1], prm.gravity.v[2]);
    auto compute_target = [&](float dtt) {
        k_target<<<gv, 256, 0, q>>>(nv, s.curr.p, s.prev.p, s.fix_index.p, s.fix.p, dtt, prev_dt, gravity, prm.inactive_momentum, s.target.p);
        LAUNCH_COUNT(1);
    };
    compute_target(dt);
    CUDA_OK(cudaMemcpyAsync(s.eval_x.p, s.curr.p, sizeof(float4) * nv, cudaMemcpyDeviceToDevice, q));

    double toi_advanced = 0.0;
    unsigned step = 1;
    bool final_step = false;
    float last_toi = 1.0f;
    const unsigned n_dof_removed = s.nfix; // every fix pin is an exact Dirichlet BC (no PDRD here)

    while (true) {
        if (!final_step && prm.max_newton_steps > 0 && step >= prm.max_newton_steps) {
            logging.message("### newton stalled: no acceptable step after %u iterations (last toi: %.2e, toi_advanced: %.2e)", step, last_toi, toi_advanced);
            result.newton_progress = false;
            finish(false);
            return;
        }
        if (final_step) logging.message("------ error reduction step ------");
        else logging.message("------ newton step %u ------", step);
        s.tm.n_newton += 1;

        CUDA_OK(cudaMemsetAsync(s.diag.p, 0, sizeof(float) * 9 * size_t(nv), q));
        CUDA_OK(cudaMemsetAsync(s.fx_val.p, 0, sizeof(float) * 9 * size_t(s.fixed_nnz), q));
        CUDA_OK(cudaMemsetAsync(s.force.p, 0, sizeof(float) * 3 * size_t(nv), q));
        s.dyn_nnz = 0;

        if (final_step) {
            if (toi_advanced < 1.0 && s.nfix) {
                k_rewind_fix<<<div_up(s.nfix, 256), 256, 0, q>>>(s.nfix, s.fix.p, 1.0f - static_cast<float>(toi_advanced));
                LAUNCH_COUNT(1);
            }
            dt *= toi_advanced;
            compute_target(dt);
        }
        logging.push("matrix assembly");
        t_ph.start();
        k_seed_dx<<<gv, 256, 0, q>>>(nv, s.fix_index.p, s.eval_x.p, s.target.p, s.dx.p);
        LAUNCH_COUNT(1);
        assemble_momentum(s, dt, prm);
        logging.push("asm elastic");
        assemble_elastic(s, dt, prm);
        assemble_damping(s, dt, prm);
        logging.pop();
        assemble_stitches(s, prm); // main.cu:793-796: after "asm elastic", before the snapshot
        logging.push("asm copy");
        CUDA_OK(cudaMemcpyAsync(s.fx_el.p, s.fx_val.p, sizeof(float) * 9 * size_t(s.fixed_nnz), cudaMemcpyDeviceToDevice, q));
        logging.pop();
        logging.push("asm strainlimit");
        assemble_strainlimit(s, prm);
        logging.pop();
        unsigned num_contact = 0, max_nnz_row = 0;
        float dyn_consumed = 0.0f;
        if (s.n_stat_obj) CUDA_OK(cudaMemsetAsync(s.stat_count.p, 0, sizeof(unsigned) * s.n_stat_obj, q)); // main.cu:825
        logging.push("asm contact");
        if (!prm.disable_contact) {
            EvTimer tc(q);
            tc.start();
            num_contact += assemble_contact(s, dt, prm, max_nnz_row, dyn_consumed);
            tc.stop(&s.tm.ms_contact);
        }
        logging.pop();
        logging.mark("dyn_consumed", dyn_consumed);
        logging.mark("max_nnz_row", max_nnz_row);
        num_contact += assemble_collision_mesh(s, prm); // static collision mesh (contact.cu:1834-1947)
        num_contact += assemble_colliders(s, prm); // embed_constraint_force_hessian (main.cu:873): spheres / walls, contact on or off
        logging.mark("num_contact", num_contact);
        t_ph.stop(&s.tm.ms_assembly);
        logging.pop();

        if (n_dof_removed > 0) {
            logging.push("asm dirichlet");
            apply_dirichlet(s);
            logging.pop();
        }

        maybe_dump_linsys(s);
        logging.push("linsolve");
        t_ph.start();
        const SolveResult sr = pcg_solve(s, prm, s.vfixed.p);
        t_ph.stop(&s.tm.ms_linsolve);
        logging.pop();
        logging.mark("iter", sr.iter);
        logging.mark("reresid", sr.resid);
        // a request for the Schwarz preconditioner that could not be served (several ranks) is reported the way the reference
        // reports its own fallback to block-Jacobi (main.cu:1124, solver.cu:2317)
        logging.mark("schwarz_fallback", prm.precond == PPF_PRECOND_SCHWARZ && !pcg_schwarz_served() ? 1 : 0);
        if (!sr.ok) {
            logging.message("### cg failed");
            result.pcg_success = false;
            finish(false);
            return;
        }

        float max_dx;
        {
            float *tmp = reinterpret_cast<float *>(s.velocity.p);
            k_dx_norm<<<gv, 256, 0, q>>>(nv, s.dx.p, tmp);
            LAUNCH_COUNT(1);
            max_dx = reduce_max_f(s, tmp, nv, 0.0f);
        }
        logging.mark("max_dx", max_dx);
        const float toi_recale = fminf(1.0f, prm.max_dx / max_dx);
        logging.mark("toi_recale", toi_recale);
        k_apply_dx<<<gv, 256, 0, q>>>(nv, s.eval_x.p, s.x_old.p, s.dx.p, toi_recale);
        LAUNCH_COUNT(1);
        if (prm.fix_xz) {
            k_fix_xz<<<gv, 256, 0, q>>>(nv, s.eval_x.p, s.prev.p, s.fix_index.p, prm.fix_xz);
            LAUNCH_COUNT(1);
        }

        float toi = 1.0f;
        if (!prm.disable_contact) {
            logging.push("aabb update");
            t_ph.start();
            static const bool swept_hull = getenv("PPF_B200_SWEPT_HULL") != nullptr; // A/B: the reference's swept-hull boxes
            auto refit = swept_hull ? bvh_refit : bvh_refit_relative;
            refit(s, s.bvh_face, 0, s.x_old.p, s.eval_x.p, prm.line_search_max_t);
            refit(s, s.bvh_edge, 1, s.x_old.p, s.eval_x.p, prm.line_search_max_t);
            refit(s, s.bvh_vert, 2, s.x_old.p, s.eval_x.p, prm.line_search_max_t);
            t_ph.stop(&s.tm.ms_aabb_update);
            logging.pop();
            logging.push("line search");
            t_ph.start();
            bool overlap = false;
            toi = fminf(toi, contact_line_search(s, s.x_old.p, s.eval_x.p, prm, overlap));
            if (s.cm_nv && !overlap) toi = fminf(toi, collision_mesh_line_search(s, s.x_old.p, s.eval_x.p, prm, overlap));
            t_ph.stop(&s.tm.ms_line_search);
            if (overlap) {
                logging.message("### contact starts overlapping: two surfaces begin the step already touching or overlapping "
                                "(a contact pair is inside the contact offset at the start of the step).");
                logging.pop();
                result.contact_separated = false;
                finish(false);
                return;
            }
            logging.pop();
        }
        if (s.nsphere || s.nfloor) { // vertex_constraint_line_search (contact.cu:2318-2470), inside contact::line_search in the reference
            unsigned bad_vert = 0xFFFFFFFFu;
            toi = fminf(toi, collider_line_search(s, s.x_old.p, s.eval_x.p, prm, bad_vert));
            if (bad_vert != 0xFFFFFFFFu) {
                logging.message("### infeasible pin: prescribed vertex %u is driven through an analytic collider (floor/sphere) it cannot yield to", bad_vert);
                logging.message("### re-author the pin's path so it stays outside the collider, or make it a soft pull pin so it can yield.");
                result.pin_feasible = false;
                finish(false);
                return;
            }
        }
        if (s.nf) { // strainlimiting::line_search (main.cu:1321-1335); the channel exists for every shell scene
            const float sl_toi = s.has_strainlimit ? strain_line_search(s, s.x_old.p, s.eval_x.p, prm) : 1.0f;
            toi = fminf(toi, sl_toi);
            logging.mark("SL_toi", sl_toi);
        }
        if (s.nrod) { // strainlimiting::rod_line_search (main.cu:1336-1343)
            const float sl_toi = s.has_rod_strainlimit ? rod_strain_line_search(s, s.x_old.p, s.eval_x.p, prm) : 1.0f;
            toi = fminf(toi, sl_toi);
            logging.mark("SL_rod_toi", sl_toi);
        }
        logging.mark("toi", toi);
        last_toi = toi;
        if (toi <= std::numeric_limits<float>::epsilon()) {
            logging.message("### ccd failed (toi: %.2e)", toi);
            result.ccd_success = false;
            finish(false);
            return;
        }
        if (!final_step) toi_advanced += std::max(0.0, 1.0 - toi_advanced) * static_cast<double>(toi_recale * toi);
        logging.message("* toi_advanced: %.2e", toi_advanced);
        k_lerp<<<gv, 256, 0, q>>>(nv, s.eval_x.p, s.x_old.p, toi);
        LAUNCH_COUNT(1);

        if (final_step) break;
        else if (toi_advanced >= s.param->target_toi && step >= s.param->min_newton_steps) final_step = true;
        else ++step; // target is kept intact here (the reference recomputes it only because it reuses the buffer)
    }

    if (!prm.disable_contact) {
        logging.push("aabb update");
        bvh_refit(s, s.bvh_face, 0, s.eval_x.p, s.eval_x.p, 1.0f);
        bvh_refit(s, s.bvh_edge, 1, s.eval_x.p, s.eval_x.p, 1.0f);
        bvh_refit(s, s.bvh_vert, 2, s.eval_x.p, s.eval_x.p, 1.0f);
        logging.pop();
        logging.push("check intersection");
        if (!check_intersection(s, s.eval_x.p, prm)) {
            logging.message("### intersection detected");
            result.intersection_free = false;
        }
        logging.pop();
    }
    logging.mark("toi_advanced", toi_advanced);
    logging.mark("newton_steps", step);
    logging.mark("final_dt", dt);
    logging.mark("device-alloc", static_cast<double>(g_dev_alloc - alloc0)); // main.cu:1636-1645
    logging.mark("device-free", static_cast<double>(g_dev_free - free0));
    logging.message("* device alloc/free this step: %llu / %llu", g_dev_alloc - alloc0, g_dev_free - free0);
    s.param->prev_dt = dt;
    s.param->time += static_cast<double>(s.param->prev_dt / s.param->playback);
    // prev <- curr, curr <- eval_x
    std::swap(s.prev.p, s.curr.p);
    std::swap(s.prev.cap, s.curr.cap);
    CUDA_OK(cudaMemcpyAsync(s.curr.p, s.eval_x.p, sizeof(float4) * nv, cudaMemcpyDeviceToDevice, q));
    apply_plasticity(s, *s.param); // main.cu:1613-1626: on the new positions, with the session's dt
    CUDA_OK(cudaStreamSynchronize(q));
    result.time = s.param->time;
    finish(true);
}

} // namespace ppf

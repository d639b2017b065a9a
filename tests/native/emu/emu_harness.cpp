// emu_harness.cpp -- TEST INFRASTRUCTURE: drives the backend's own upload + assembly code (csrc/abi.cu, csrc/elastic.cu,
// rewritten for the host by tests/test_kernel_emulation.py, see cuda_runtime.h next to this file) on a scene built by the C++
// host, and hands the assembled gradient / fixed-matrix blocks / diagonal blocks back for comparison with the oracle.
#include "backend.hpp"
#include "../../../ppf-contact-solver_b200/host/ppf_host.hpp"
#include <cstdarg>
#include <cstdio>
#include <vector>

PpfEmuDim blockIdx, blockDim, threadIdx;

// ---- pieces of the backend that the emulated files reference but that are not part of this check ----
// (csrc/step.cu and csrc/pcg.cu are not emulated: the PCG kernels stream the matrix with cp.async.bulk PTX)
namespace ppf {
void step_advance(State &, PpfStepResult *) {}
float reduce_max_f(State &, const float *d, size_t n, float init) { float m = init; for (size_t i = 0; i < n; ++i) m = d[i] > m ? d[i] : m; return m; }
float reduce_min_f(State &, const float *d, size_t n, float init) { float m = init; for (size_t i = 0; i < n; ++i) m = d[i] < m ? d[i] : m; return m; }
unsigned reduce_sum_u(State &, const unsigned *d, size_t n) { unsigned a = 0; for (size_t i = 0; i < n; ++i) a += d[i]; return a; }
struct ScopeLog::Impl {};
ScopeLog::ScopeLog(const char *) : impl(nullptr) {}
ScopeLog::~ScopeLog() {}
void ScopeLog::push(const char *) {}
void ScopeLog::pop() {}
void ScopeLog::mark(const char *, double) {}
void ScopeLog::message(const char *, ...) {}
void simplelog_set_path(const char *) {}
void simplelog_set_time(double) {}
} // namespace ppf
extern "C" void print_rust(const char *) {}

extern "C" __attribute__((visibility("default")))
int emu_assemble(const char *scene, unsigned n, const float *eval_x, float dt, float *force, float *fx_val, float *diag, unsigned *counts) {
    using namespace ppf;
    static ppfh::Built b; // keeps the host arrays alive
    ppfh::build_scene(ppfh::make_scene(scene, n), b);
    State &s = g;
    s.host = b.dataset;
    s.param = &b.param;
    s.stream = nullptr;
    upload_dataset(s);
    const unsigned nv = s.nv;
    counts[0] = nv; counts[1] = s.fixed_nnz; counts[2] = s.nrod; counts[3] = s.nrodbend; counts[4] = s.nstitch;
    if (!eval_x) return 0;
    for (unsigned i = 0; i < nv; ++i) s.eval_x.p[i] = make_float4(eval_x[3 * i], eval_x[3 * i + 1], eval_x[3 * i + 2], 0.f);
    std::memcpy(s.target.p, s.curr.p, sizeof(float4) * nv);
    std::memset(s.diag.p, 0, sizeof(float) * 9 * size_t(nv));
    std::memset(s.fx_val.p, 0, sizeof(float) * 9 * size_t(s.fixed_nnz));
    std::memset(s.force.p, 0, sizeof(float) * 3 * size_t(nv));
    const PpfParamSet prm = *s.param;
    assemble_momentum(s, dt, prm);
    assemble_elastic(s, dt, prm);
    assemble_damping(s, dt, prm);
    assemble_stitches(s, prm);
    std::memcpy(s.fx_el.p, s.fx_val.p, sizeof(float) * 9 * size_t(s.fixed_nnz));
    assemble_strainlimit(s, prm);
    std::memcpy(force, s.force.p, sizeof(float) * 3 * size_t(nv));
    std::memcpy(fx_val, s.fx_val.p, sizeof(float) * 9 * size_t(s.fixed_nnz));
    std::memcpy(diag, s.diag.p, sizeof(float) * 9 * size_t(nv));
    return 1;
}

// "max_sigma" diagnostic (k_face_sigma / k_rod_sigma) at caller-provided start-of-step positions
extern "C" __attribute__((visibility("default")))
float emu_max_sigma(const char *scene, unsigned n, const float *x) {
    using namespace ppf;
    static ppfh::Built b;
    ppfh::build_scene(ppfh::make_scene(scene, n), b);
    State &s = g;
    s.host = b.dataset;
    s.param = &b.param;
    s.stream = nullptr;
    upload_dataset(s);
    if (x) for (unsigned i = 0; i < s.nv; ++i) s.curr.p[i] = make_float4(x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.f);
    return max_stretch_ratio(s);
}

// plastic creep kernels (k_plastic_*) at caller-provided new positions, then the ABI's fetch_inv_rest / fetch_rest_angles:
// returns what a host would read back (inverse rest matrices, hinge rest angles, per-vertex rest bend angles)
extern "C" void fetch_inv_rest(void);
extern "C" void fetch_rest_angles(void);
extern "C" __attribute__((visibility("default")))
void emu_plasticity(const char *scene, unsigned n, const float *x, float *inv_rest, float *hinge_rest, float *vertex_rest) {
    using namespace ppf;
    static ppfh::Built b;
    ppfh::build_scene(ppfh::make_scene(scene, n), b);
    State &s = g;
    s.host = b.dataset;
    s.param = &b.param;
    s.stream = nullptr;
    upload_dataset(s);
    for (unsigned i = 0; i < s.nv; ++i) s.curr.p[i] = make_float4(x[3 * i], x[3 * i + 1], x[3 * i + 2], 0.f);
    apply_plasticity(s, *s.param);
    fetch_inv_rest();
    fetch_rest_angles();
    for (unsigned i = 0; i < s.nf; ++i) std::memcpy(inv_rest + 4 * size_t(i), b.inv_rest2x2[i].m, 16);
    for (unsigned i = 0; i < s.nh; ++i) hinge_rest[i] = b.hprop[i].rest_angle;
    for (unsigned i = 0; i < s.nv; ++i) vertex_rest[i] = b.vprop[i].rest_bend_angle;
}

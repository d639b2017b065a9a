// emu_harness.cpp -- TEST INFRASTRUCTURE: drives the backend's own upload + assembly code (csrc/abi.cu, csrc/elastic.cu,
// rewritten for the host by tests/test_kernel_emulation.py, see cuda_runtime.h next to this file) on a scene built by the C++
// host, and hands the assembled gradient / fixed-matrix blocks / diagonal blocks back for comparison with the oracle.
#include <algorithm>
#include <vector>
#include "backend.hpp"
#include "dist.hpp"
#include "../../../ppf-contact-solver_b200/host/ppf_host.hpp"
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

PpfEmuDim blockIdx, blockDim, threadIdx;

// ---- pieces of the backend that the emulated files reference but that are not part of this check ----
// (csrc/step.cu and csrc/pcg.cu are not emulated: the PCG kernels stream the matrix with cp.async.bulk PTX)
namespace ppf {
// Stand-ins for csrc/pcg.cu, whose kernels cannot run on the host: the Dirichlet elimination restated from k_dirichlet_rows /
// k_dirichlet_diag (pcg.cu:198-238) and a plain serial block-Jacobi PCG on the same three matrices (fixed upper-triangle
// block CSR + dynamic block CSR + per-vertex diagonal) with the same stopping rule (L1 residual ratio below cg_tol).  With them
// the backend's real step.cu / elastic.cu / contact.cu / bvh.cu run a whole advance() on the CPU.
static void lift_sub_host(float *force, unsigned row, const float *b, float px, float py, float pz, bool tr) {
    const float c0 = tr ? b[0] * px + b[1] * py + b[2] * pz : b[0] * px + b[3] * py + b[6] * pz;
    const float c1 = tr ? b[3] * px + b[4] * py + b[5] * pz : b[1] * px + b[4] * py + b[7] * pz;
    const float c2 = tr ? b[6] * px + b[7] * py + b[8] * pz : b[2] * px + b[5] * py + b[8] * pz;
    force[3 * size_t(row)] -= c0; force[3 * size_t(row) + 1] -= c1; force[3 * size_t(row) + 2] -= c2;
}
static void dirichlet_rows_host(State &s, const unsigned *off, const unsigned *col, float *val) {
    for (unsigned i = 0; i < s.nv; ++i) {
        const bool mi = s.vfixed.p[i] != 0;
        for (unsigned k = off[i]; k < off[i + 1]; ++k) {
            const unsigned j = col[k];
            const bool mj = s.vfixed.p[j] != 0;
            float *blk = val + size_t(9) * k;
            if (mi != mj) {
                const unsigned fixed = mj ? j : i;
                const float4 e = s.eval_x.p[fixed], t = s.target.p[fixed];
                lift_sub_host(s.force.p, mj ? i : j, blk, e.x - t.x, e.y - t.y, e.z - t.z, !mj);
            }
            if (mi || mj) for (int c = 0; c < 9; ++c) blk[c] = 0.0f;
        }
    }
}
void apply_dirichlet(State &s) {
    if (s.dyn_nnz) dirichlet_rows_host(s, s.dyn_off.p, s.dyn_col.p, s.dyn_val.p);
    dirichlet_rows_host(s, s.fx_off.p, s.fx_col.p, s.fx_val.p);
    for (unsigned i = 0; i < s.nv; ++i) {
        if (!s.vfixed.p[i]) continue;
        float *d = s.diag.p + size_t(9) * i;
        for (int c = 0; c < 9; ++c) d[c] = (c == 0 || c == 4 || c == 8) ? 1.0f : 0.0f;
        const float4 e = s.eval_x.p[i], t = s.target.p[i];
        s.force.p[3 * size_t(i)] = e.x - t.x; s.force.p[3 * size_t(i) + 1] = e.y - t.y; s.force.p[3 * size_t(i) + 2] = e.z - t.z;
    }
}
bool pcg_schwarz_served() { return false; } // the emulated build solves with block-Jacobi only
SolveResult pcg_solve(State &s, const PpfParamSet &prm, const unsigned char *) {
    const unsigned n = s.nv;
    std::vector<double> D(size_t(9) * n);
    for (size_t k = 0; k < D.size(); ++k) D[k] = s.diag.p[k];
    struct Off { unsigned i, j; const float *b; };
    std::vector<Off> offs;
    auto take = [&](const unsigned *off, const unsigned *col, const float *val) {
        for (unsigned i = 0; i < n; ++i)
            for (unsigned k = off[i]; k < off[i + 1]; ++k) {
                if (col[k] == i) for (int e = 0; e < 9; ++e) D[size_t(9) * i + e] += val[size_t(9) * k + e];
                else offs.push_back(Off{i, col[k], val + size_t(9) * k});
            }
    };
    take(s.fx_off.p, s.fx_col.p, s.fx_val.p);
    if (s.dyn_nnz) take(s.dyn_off.p, s.dyn_col.p, s.dyn_val.p);
    std::vector<double> Pinv(size_t(9) * n);
    for (unsigned i = 0; i < n; ++i) { // inverse of the symmetrised diagonal block (column-major)
        const double *m = &D[size_t(9) * i];
        const double a = m[0], b = 0.5 * (m[3] + m[1]), c = 0.5 * (m[6] + m[2]), d = m[4], e = 0.5 * (m[7] + m[5]), f = m[8];
        const double det = a * (d * f - e * e) - b * (b * f - e * c) + c * (b * e - d * c);
        double *o = &Pinv[size_t(9) * i];
        o[0] = (d * f - e * e) / det; o[1] = o[3] = (c * e - b * f) / det; o[2] = o[6] = (b * e - c * d) / det;
        o[4] = (a * f - c * c) / det; o[5] = o[7] = (b * c - a * e) / det; o[8] = (a * d - b * b) / det;
    }
    auto apply = [&](const std::vector<double> &v, std::vector<double> &out) {
        for (unsigned i = 0; i < n; ++i) {
            const double *m = &D[size_t(9) * i], *x = &v[3 * size_t(i)];
            for (int q = 0; q < 3; ++q) out[3 * size_t(i) + q] = m[q] * x[0] + m[3 + q] * x[1] + m[6 + q] * x[2];
        }
        for (const Off &o : offs) {
            const float *m = o.b;
            const double *xi = &v[3 * size_t(o.i)], *xj = &v[3 * size_t(o.j)];
            for (int q = 0; q < 3; ++q) {
                out[3 * size_t(o.i) + q] += m[q] * xj[0] + m[3 + q] * xj[1] + m[6 + q] * xj[2];           // B x_j
                out[3 * size_t(o.j) + q] += m[3 * q] * xi[0] + m[3 * q + 1] * xi[1] + m[3 * q + 2] * xi[2]; // B^T x_i
            }
        }
    };
    const size_t n3 = 3 * size_t(n);
    std::vector<double> x(n3), r(n3), z(n3), p(n3), Ap(n3);
    for (size_t k = 0; k < n3; ++k) { x[k] = s.dx.p[k]; r[k] = s.force.p[k]; }
    apply(x, Ap);
    double err0 = 0, rz0 = 0;
    auto precond = [&]() {
        for (unsigned i = 0; i < n; ++i) {
            const double *m = &Pinv[size_t(9) * i], *ri = &r[3 * size_t(i)];
            for (int q = 0; q < 3; ++q) z[3 * size_t(i) + q] = m[q] * ri[0] + m[3 + q] * ri[1] + m[6 + q] * ri[2];
        }
    };
    for (size_t k = 0; k < n3; ++k) r[k] -= Ap[k];
    precond();
    for (size_t k = 0; k < n3; ++k) { p[k] = z[k]; err0 += std::fabs(r[k]); rz0 += r[k] * z[k]; }
    SolveResult res{true, 1, 0.0f};
    if (err0 == 0.0) return res;
    for (unsigned iter = 1;; ++iter) {
        apply(p, Ap);
        double pAp = 0;
        for (size_t k = 0; k < n3; ++k) pAp += p[k] * Ap[k];
        if (!(pAp > 0.0)) { res.ok = false; res.iter = iter; return res; }
        const double alpha = rz0 / pAp;
        for (size_t k = 0; k < n3; ++k) { x[k] += alpha * p[k]; r[k] -= alpha * Ap[k]; }
        precond();
        double err = 0, rz1 = 0;
        for (size_t k = 0; k < n3; ++k) { err += std::fabs(r[k]); rz1 += r[k] * z[k]; }
        res.iter = iter; res.resid = float(err / err0);
        if (err / err0 < prm.cg_tol) break;
        if (iter >= prm.cg_max_iter) { res.ok = false; return res; }
        const double beta = rz1 / rz0;
        rz0 = rz1;
        for (size_t k = 0; k < n3; ++k) p[k] = z[k] + beta * p[k];
    }
    for (size_t k = 0; k < n3; ++k) s.dx.p[k] = float(x[k]);
    s.tm.n_cg_iter += res.iter;
    return res;
}
float reduce_max_f(State &, const float *d, size_t n, float init) { float m = init; for (size_t i = 0; i < n; ++i) m = d[i] > m ? d[i] : m; return m; }
float reduce_min_f(State &, const float *d, size_t n, float init) { float m = init; for (size_t i = 0; i < n; ++i) m = d[i] < m ? d[i] : m; return m; }
unsigned reduce_sum_u(State &, const unsigned *d, size_t n) { unsigned a = 0; for (size_t i = 0; i < n; ++i) a += d[i]; return a; }
// dist.cu (NCCL plumbing) is not part of the emulated build: one rank, the collectives are never reached
Dist g_dist;
void dist_allgather(const void *, void *, size_t, cudaStream_t) {}
void dist_allreduce_u32(unsigned *, size_t, bool, cudaStream_t) {}
// prims.cu's sort / unique (device kernels there): the same contracts on the host
void radix_sort_pairs_u32(State &, unsigned *keys, unsigned *keys_sorted, unsigned *vals, unsigned *vals_sorted, unsigned n, int bits) {
    std::vector<unsigned> order(n);
    for (unsigned i = 0; i < n; ++i) order[i] = i;
    const unsigned mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return (keys[a] & mask) < (keys[b] & mask); });
    for (unsigned i = 0; i < n; ++i) { keys_sorted[i] = keys[order[i]]; vals_sorted[i] = vals[order[i]]; }
}
void radix_sort_keys_u64(State &, unsigned long long *keys, unsigned long long *keys_sorted, unsigned n, int bits) {
    const unsigned long long mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    std::vector<unsigned long long> v(keys, keys + n);
    std::stable_sort(v.begin(), v.end(), [&](unsigned long long a, unsigned long long b) { return (a & mask) < (b & mask); });
    for (unsigned i = 0; i < n; ++i) keys_sorted[i] = v[i];
}
void unique_sorted_u64(State &, const unsigned long long *keys_sorted, unsigned long long *out, unsigned n, unsigned *d_count) {
    unsigned m = 0;
    for (unsigned i = 0; i < n; ++i) if (i == 0 || keys_sorted[i] != keys_sorted[i - 1]) out[m++] = keys_sorted[i];
    *d_count = m;
}
} // namespace ppf
extern "C" __attribute__((weak)) void print_rust(const char *) {}

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

// test_math.cpp -- CPU check of the backend's per-element math (csrc/math_*.cuh, compiled for the
// host in THIS test only) against the reference's own headers compiled for the host
// (oracle/_ref/libref_host.so, built from /root/reference by oracle/Makefile.ref).
// Both sides are built with -ffp-contract=off, so a bitwise match means the two
// sources perform the same fp32 operations in the same order; on the device the compiler then
// contracts the same expression trees the same way.
//
// Usage: test_math <path/to/libref_host.so> [n_cases]      exit code 0 = all checks passed
#include "../../ppf-contact-solver_b200/csrc/math_accd.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_contact.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_distance.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_elastic.cuh"
#include "../../ppf-contact-solver_b200/csrc/math_hinge.cuh"
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <random>

using namespace ppf;

static void *lib;
template <class F> static F sym(const char *n) {
    void *p = dlsym(lib, n);
    if (!p) { fprintf(stderr, "missing %s\n", n); exit(2); }
    return reinterpret_cast<F>(p);
}
static bool same_bits(float a, float b) {
    uint32_t x, y;
    memcpy(&x, &a, 4); memcpy(&y, &b, 4);
    return x == y || (a != a && b != b) || (a == 0.0f && b == 0.0f);
}
static std::mt19937 rng(12345);
static float uni(float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(rng); }
static V3 rv(float s) { return mk(uni(-s, s), uni(-s, s), uni(-s, s)); }
static void put(float *d, const V3 &v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails < 20) { fprintf(stderr, "FAIL %s:%d ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } ++fails; } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: test_math libref_host.so [n]\n"); return 2; }
    lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    const int N = argc > 2 ? atoi(argv[2]) : 20000;
    auto r_pt = sym<void (*)(const float *, const float *, const float *, const float *, float *)>("ref_point_triangle_coeff");
    auto r_ptu = sym<void (*)(const float *, const float *, const float *, const float *, float *)>("ref_point_triangle_coeff_unclassified");
    auto r_pe = sym<void (*)(const float *, const float *, const float *, float *)>("ref_point_edge_coeff");
    auto r_ee = sym<void (*)(const float *, const float *, const float *, const float *, float *)>("ref_edge_edge_coeff");
    typedef float (*ccd_fn)(const float *, const float *, float, float, float, float, unsigned *);
    auto r_ptccd = sym<ccd_fn>("ref_point_triangle_ccd");
    auto r_peccd = sym<ccd_fn>("ref_point_edge_ccd");
    auto r_ppccd = sym<ccd_fn>("ref_point_point_ccd");
    auto r_eeccd = sym<ccd_fn>("ref_edge_edge_ccd");
    auto r_isect = sym<int (*)(const float *, const float *, const float *, const float *, const float *)>("ref_edge_triangle_intersect");
    auto r_cubic = sym<void (*)(float, float, float, float *)>("ref_cubic");
    auto r_fric = sym<void (*)(const float *, const float *, const float *, float, float, float *, float *, float *)>("ref_friction");
    auto r_bw = sym<void (*)(const float *, float, float, float *, float *)>("ref_baraffwitkin");
    auto r_hinge = sym<void (*)(const float *, float, float *, float *, float *)>("ref_hinge");

    // ---- distance coefficients: bitwise ----
    for (int it = 0; it < N; ++it) {
        const float sc = it % 3 == 0 ? 1e-2f : 1.0f;
        V3 p = rv(sc), a = rv(sc), b = rv(sc), c = rv(sc);
        if (it % 7 == 0) c = a + 0.5f * (b - a); // degenerate triangle
        if (it % 11 == 0) b = a;                  // degenerate edge
        float P[3], A[3], B[3], C[3], o[4];
        put(P, p); put(A, a); put(B, b); put(C, c);
        r_pt(P, A, B, C, o);
        C3 m = point_triangle_coeff(p, a, b, c);
        CHECK(same_bits(o[0], m.a) && same_bits(o[1], m.b) && same_bits(o[2], m.c), "point_triangle_coeff it=%d (%g %g %g) vs (%g %g %g)", it, o[0], o[1], o[2], m.a, m.b, m.c);
        r_ptu(P, A, B, C, o);
        m = point_triangle_coeff_unclassified(p, a, b, c);
        CHECK(same_bits(o[0], m.a) && same_bits(o[1], m.b) && same_bits(o[2], m.c), "point_triangle_unclassified it=%d", it);
        r_pe(P, A, B, o);
        C2 e = point_edge_coeff(p, a, b);
        CHECK(same_bits(o[0], e.a) && same_bits(o[1], e.b), "point_edge_coeff it=%d", it);
        V3 d = rv(sc);
        if (it % 5 == 0) d = c + (b - a); // parallel edges
        float D[3];
        put(D, d);
        r_ee(A, B, C, D, o);
        C4 q = edge_edge_coeff(a, b, c, d);
        CHECK(same_bits(o[0], q.a) && same_bits(o[1], q.b) && same_bits(o[2], q.c) && same_bits(o[3], q.d), "edge_edge_coeff it=%d", it);
        int ri = r_isect(P, A, B, C, D);
        // our device version lives in contact.cu; the host copy of the same arithmetic is exercised through the oracle tests
        (void)ri;
    }
    // ---- ACCD: bitwise time of impact ----
    for (int it = 0; it < N / 4; ++it) {
        const float sc = 0.1f, mv = it % 2 ? 0.3f : 0.02f;
        float x0[12], x1[12];
        V3 pts[4], dsp[4];
        for (int k = 0; k < 4; ++k) { pts[k] = rv(sc); dsp[k] = rv(mv); }
        if (it % 3 == 0) { pts[0] = (1.0f / 3.0f) * (pts[1] + pts[2] + pts[3]) + mk(0.f, 0.f, 0.02f); dsp[0] = mk(0.f, 0.f, -0.2f); }
        for (int k = 0; k < 4; ++k) { put(x0 + 3 * k, pts[k]); put(x1 + 3 * k, pts[k] + dsp[k]); }
        const float offset = it % 5 == 0 ? 1e-3f : 0.0f, ghat = 1e-3f, max_t = 1.25f, eps = 1e-7f;
        unsigned rov = 0;
        CcdOverlap ov{0u, -1.f, -1.f};
        float rr = r_ptccd(x0, x1, offset, ghat, max_t, eps, &rov);
        float mm = point_triangle_ccd(pts[0], pts[0] + dsp[0], pts[1], pts[2], pts[3], pts[1] + dsp[1], pts[2] + dsp[2], pts[3] + dsp[3], offset, ghat, max_t, eps, ov);
        CHECK(same_bits(rr, mm) && rov == ov.flag, "point_triangle_ccd it=%d ref=%.9g mine=%.9g ov %u/%u", it, rr, mm, rov, ov.flag);
        ov = CcdOverlap{0u, -1.f, -1.f};
        rr = r_eeccd(x0, x1, offset, ghat, max_t, eps, &rov);
        mm = edge_edge_ccd(pts[0], pts[1], pts[2], pts[3], pts[0] + dsp[0], pts[1] + dsp[1], pts[2] + dsp[2], pts[3] + dsp[3], offset, ghat, max_t, eps, ov);
        CHECK(same_bits(rr, mm) && rov == ov.flag, "edge_edge_ccd it=%d ref=%.9g mine=%.9g", it, rr, mm);
        ov = CcdOverlap{0u, -1.f, -1.f};
        rr = r_peccd(x0, x1, offset, ghat, max_t, eps, &rov);
        mm = point_edge_ccd(pts[0], pts[0] + dsp[0], pts[1], pts[2], pts[1] + dsp[1], pts[2] + dsp[2], offset, ghat, max_t, eps, ov);
        CHECK(same_bits(rr, mm) && rov == ov.flag, "point_edge_ccd it=%d ref=%.9g mine=%.9g", it, rr, mm);
        ov = CcdOverlap{0u, -1.f, -1.f};
        rr = r_ppccd(x0, x1, offset, ghat, max_t, eps, &rov);
        mm = point_point_ccd(pts[0], pts[0] + dsp[0], pts[1], pts[1] + dsp[1], offset, ghat, max_t, eps, ov);
        CHECK(same_bits(rr, mm) && rov == ov.flag, "point_point_ccd it=%d ref=%.9g mine=%.9g", it, rr, mm);
    }
    // ---- barrier / friction: bitwise ----
    for (int it = 0; it < N; ++it) {
        const float ghat = uni(1e-4f, 1e-2f), offset = it % 2 ? 0.0f : uni(0.f, 1e-3f), g = offset + uni(0.f, 1.2f) * ghat;
        float o[3];
        r_cubic(g, ghat, offset, o);
        CHECK(same_bits(o[1], cubic_gradient(g, ghat, offset)) && same_bits(o[2], cubic_curvature(g, ghat, offset)), "cubic it=%d", it);
        V3 fc = rv(10.f), dx = rv(it % 2 ? 1e-6f : 1e-3f), n = normalized(rv(1.f));
        float FC[3], DX[3], NN[3], gr[3], hs[9], lam;
        put(FC, fc); put(DX, dx); put(NN, n);
        r_fric(FC, DX, NN, 0.5f, 1e-5f, gr, hs, &lam);
        FrictionOut fo = friction_eval(fc, dx, n, 0.5f, 1e-5f);
        bool ok = same_bits(lam, fo.lambda) && same_bits(gr[0], fo.grad.x) && same_bits(gr[1], fo.grad.y) && same_bits(gr[2], fo.grad.z);
        for (int k = 0; k < 9; ++k) ok = ok && same_bits(hs[k], fo.hess.m[k]);
        CHECK(ok, "friction it=%d", it);
    }
    // ---- Baraff-Witkin in F space and hinge: tolerance (algebra reorganised on our side) ----
    double worst_bw = 0, worst_h = 0, worst_hf = 0;
    for (int it = 0; it < N / 4; ++it) {
        // rest triangle in the plane, deformed copy
        const V3 X0 = mk(0, 0, 0), X1 = mk(uni(0.5f, 1.5f), uni(-0.2f, 0.2f), 0), X2 = mk(uni(-0.3f, 0.6f), uni(0.5f, 1.5f), 0);
        const float m00 = X1.x, m10 = X1.y, m01 = X2.x, m11 = X2.y, det = m00 * m11 - m01 * m10;
        const float inv[4] = {m11 / det, -m10 / det, -m01 / det, m00 / det};
        const float amp = it % 2 ? 0.3f : 0.02f;
        const V3 x0 = X0 + rv(amp), x1 = X1 + rv(amp), x2 = X2 + rv(amp);
        FaceElastic fe;
        const float mu = 3000.f, lmd = 8000.f;
        bw_face(x0, x1, x2, inv, mu, lmd, 1.0f, fe);
        // reference: F-space quantities, then the same chain rule in double
        const V3 d0 = x1 - x0, d1 = x2 - x0;
        float F[6];
        const V3 fu = d0 * inv[0] + d1 * inv[1], fv = d0 * inv[2] + d1 * inv[3];
        put(F, fu); put(F + 3, fv);
        float dedF[6], H[36];
        r_bw(F, mu, lmd, dedF, H);
        const double g[3][2] = {{-(double)inv[0] - inv[1], -(double)inv[2] - inv[3]}, {inv[0], inv[2]}, {inv[1], inv[3]}};
        double scale = 1e-30;
        for (int k = 0; k < 36; ++k) scale = std::max(scale, (double)fabsf(H[k]));
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < 3; ++r) {
                        double ref = 0;
                        for (int d = 0; d < 2; ++d)
                            for (int e = 0; e < 2; ++e) ref += g[a][d] * g[b][e] * H[(3 * e + c) * 6 + (3 * d + r)];
                        worst_bw = std::max(worst_bw, fabs(ref - fe.H[a][b].m[c * 3 + r]) / (scale * 4.0));
                    }
        for (int a = 0; a < 3; ++a)
            for (int r = 0; r < 3; ++r) {
                const double ref = g[a][0] * dedF[r] + g[a][1] * dedF[3 + r];
                worst_bw = std::max(worst_bw, fabs(ref - comp(fe.f[a], r)) / (1.0 + fabs(ref)));
            }
    }
    for (int it = 0; it < N / 4; ++it) {
        // hinge h = (h0,h1 shared edge; h2,h3 opposite), random fold
        V3 h[4];
        h[0] = mk(0, 0, 0) + rv(0.05f); h[1] = mk(1, 0, 0) + rv(0.05f);
        const float ang = uni(-2.5f, 2.5f);
        h[2] = mk(0.5f, 0.8f, 0) + rv(0.05f);
        h[3] = mk(0.5f, -0.8f * cosf(ang), 0.8f * sinf(ang)) + rv(0.05f);
        float x12[12], f12[12], H144[144], angle;
        for (int k = 0; k < 4; ++k) put(x12 + 3 * k, h[k]);
        const float rest = it % 2 ? 0.0f : 0.3f;
        r_hinge(x12, rest, f12, H144, &angle);
        const V3 X[4] = {h[2], h[1], h[0], h[3]};
        HingeElastic he;
        hinge_force_hessian(X, rest, he);
        double sc = 1e-30;
        for (int k = 0; k < 144; ++k) sc = std::max(sc, (double)fabsf(H144[k]));
        for (int k = 0; k < 144; ++k) worst_h = std::max(worst_h, fabs((double)H144[k] - he.H[k]) / sc);
        for (int k = 0; k < 4; ++k)
            for (int r = 0; r < 3; ++r) worst_hf = std::max(worst_hf, fabs((double)f12[3 * k + r] - comp(he.f[k], r)) / (1.0 + fabs(f12[3 * k + r])));
    }
    printf("worst relative deviation: baraff-witkin %.3e, hinge hessian %.3e, hinge force %.3e\n", worst_bw, worst_h, worst_hf);
    CHECK(worst_bw < 2e-4, "baraff-witkin deviates %.3e", worst_bw);
    CHECK(worst_h < 1e-5, "hinge hessian deviates %.3e", worst_h);
    CHECK(worst_hf < 1e-5, "hinge force deviates %.3e", worst_hf);
    printf("%s (%d failures)\n", fails ? "FAILED" : "OK", fails);
    return fails ? 1 : 0;
}

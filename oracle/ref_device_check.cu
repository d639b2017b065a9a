// ref_device_check.cu -- TEST INFRASTRUCTURE (oracle), not product code.
//
// GPU-side bitwise check of the backend's contact predicates / ACCD (csrc/math_*.cuh) against the
// reference's own device headers, both compiled by the same nvcc with FMA contraction ON (the
// default), i.e. with the arithmetic the two real backends execute.  The CPU twin of this check
// (tests/native/test_math.cpp) runs with contraction off; this one catches differences that only
// appear once the compiler fuses multiply-adds.
//
// Built here (where /root/reference exists) by oracle/Makefile.ref into oracle/_ref/ref_device_check;
// the binary travels to the GPU box.  Reference headers are included from where they lie.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "data.hpp"
#include "contact/distance.hpp"
#include "contact/accd.hpp"
#include "../ppf-contact-solver_b200/csrc/math_accd.cuh"
#include "../ppf-contact-solver_b200/csrc/math_distance.cuh"

namespace accd {
__device__ unsigned g_ccd_overlap = 0u, g_ccd_overlap_v0 = 0xFFFFFFFFu, g_ccd_overlap_v1 = 0xFFFFFFFFu, g_ccd_overlap_kind = 0xFFFFFFFFu;
__device__ float g_ccd_overlap_d2 = -1.0f, g_ccd_overlap_offset = -1.0f;
}
unsigned long long g_device_alloc_count = 0, g_device_free_count = 0;

__device__ inline Vec3f rv(const float *p) { return Vec3f(p[0], p[1], p[2]); }
__device__ inline ppf::V3 mv(const float *p) { return ppf::mk(p[0], p[1], p[2]); }

// in: 12 floats per case (4 points); out_ref / out_mine: 16 floats per case
__global__ void k_distance(const float *in, unsigned n, float *out_ref, float *out_mine) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + 12 * i;
    float *r = out_ref + 16 * i, *m = out_mine + 16 * i;
    {
        Vec3f c = distance::point_triangle_distance_coeff<float, float>(rv(p), rv(p + 3), rv(p + 6), rv(p + 9));
        r[0] = c[0]; r[1] = c[1]; r[2] = c[2];
        Vec3f e = c[0] * (rv(p) - rv(p + 3)) + c[1] * (rv(p) - rv(p + 6)) + c[2] * (rv(p) - rv(p + 9));
        r[3] = e.squaredNorm();
        ppf::C3 d = ppf::point_triangle_coeff(mv(p), mv(p + 3), mv(p + 6), mv(p + 9));
        m[0] = d.a; m[1] = d.b; m[2] = d.c;
        ppf::V3 f = ppf::lin3(d.a, mv(p) - mv(p + 3), d.b, mv(p) - mv(p + 6), d.c, mv(p) - mv(p + 9));
        m[3] = ppf::sqnorm(f);
    }
    {
        Vec2f c = distance::point_edge_distance_coeff<float, float>(rv(p), rv(p + 3), rv(p + 6));
        r[4] = c[0]; r[5] = c[1];
        Vec3f e = c[0] * (rv(p) - rv(p + 3)) + c[1] * (rv(p) - rv(p + 6));
        r[6] = e.squaredNorm();
        ppf::C2 d = ppf::point_edge_coeff(mv(p), mv(p + 3), mv(p + 6));
        m[4] = d.a; m[5] = d.b;
        ppf::V3 f = ppf::lin2(d.a, mv(p) - mv(p + 3), d.b, mv(p) - mv(p + 6));
        m[6] = ppf::sqnorm(f);
    }
    {
        Vec4f c = distance::edge_edge_distance_coeff<float, float>(rv(p), rv(p + 3), rv(p + 6), rv(p + 9));
        r[7] = c[0]; r[8] = c[1]; r[9] = c[2]; r[10] = c[3];
        Vec3f x0 = float(c[0]) * rv(p) + float(c[1]) * rv(p + 3);
        Vec3f x1 = float(c[2]) * rv(p + 6) + float(c[3]) * rv(p + 9);
        Vec3f e = x0 - x1;
        r[11] = e.squaredNorm();
        ppf::C4 d = ppf::edge_edge_coeff(mv(p), mv(p + 3), mv(p + 6), mv(p + 9));
        m[7] = d.a; m[8] = d.b; m[9] = d.c; m[10] = d.d;
        ppf::V3 a = ppf::lin2(d.a, mv(p), d.b, mv(p + 3)), b = ppf::lin2(d.c, mv(p + 6), d.d, mv(p + 9));
        ppf::V3 f = a - b;
        m[11] = ppf::sqnorm(f);
    }
    {
        Vec3f c = distance::point_triangle_distance_coeff_unclassified<float, float>(rv(p), rv(p + 3), rv(p + 6), rv(p + 9));
        r[12] = c[0]; r[13] = c[1]; r[14] = c[2];
        ppf::C3 d = ppf::point_triangle_coeff_unclassified(mv(p), mv(p + 3), mv(p + 6), mv(p + 9));
        m[12] = d.a; m[13] = d.b; m[14] = d.c;
        r[15] = m[15] = 0.0f;
    }
}

// in: 24 floats per case (4 points at t0, 4 points at t1); out: 4 floats per case (pt, ee, pe, pp)
__global__ void k_ccd(const float *in, unsigned n, float offset, float ghat, float *out_ref, float *out_mine) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *a = in + 24 * i, *b = a + 12;
    ParamSet prm;
    prm.line_search_max_t = 1.25f;
    prm.ccd_eps = 1e-7f;
    float *r = out_ref + 4 * i, *m = out_mine + 4 * i;
    r[0] = accd::point_triangle_ccd(rv(a), rv(b), rv(a + 3), rv(a + 6), rv(a + 9), rv(b + 3), rv(b + 6), rv(b + 9), offset, ghat, prm);
    r[1] = accd::edge_edge_ccd(rv(a), rv(a + 3), rv(a + 6), rv(a + 9), rv(b), rv(b + 3), rv(b + 6), rv(b + 9), offset, ghat, prm);
    r[2] = accd::point_edge_ccd(rv(a), rv(b), rv(a + 3), rv(a + 6), rv(b + 3), rv(b + 6), offset, ghat, prm);
    r[3] = accd::point_point_ccd(rv(a), rv(b), rv(a + 3), rv(b + 3), offset, ghat, prm);
    ppf::CcdOverlap ov{0u, -1.f, -1.f};
    m[0] = ppf::point_triangle_ccd(mv(a), mv(b), mv(a + 3), mv(a + 6), mv(a + 9), mv(b + 3), mv(b + 6), mv(b + 9), offset, ghat, 1.25f, 1e-7f, ov);
    m[1] = ppf::edge_edge_ccd(mv(a), mv(a + 3), mv(a + 6), mv(a + 9), mv(b), mv(b + 3), mv(b + 6), mv(b + 9), offset, ghat, 1.25f, 1e-7f, ov);
    m[2] = ppf::point_edge_ccd(mv(a), mv(b), mv(a + 3), mv(a + 6), mv(b + 3), mv(b + 6), offset, ghat, 1.25f, 1e-7f, ov);
    m[3] = ppf::point_point_ccd(mv(a), mv(b), mv(a + 3), mv(b + 3), offset, ghat, 1.25f, 1e-7f, ov);
}

static bool same_bits(float a, float b) {
    unsigned x, y;
    memcpy(&x, &a, 4); memcpy(&y, &b, 4);
    return x == y || (a != a && b != b) || (a == 0.0f && b == 0.0f);
}

int main(int argc, char **argv) {
    const unsigned n = argc > 1 ? (unsigned)atoi(argv[1]) : 400000;
    std::mt19937 rng(777);
    auto uni = [&](float lo, float hi) { return std::uniform_real_distribution<float>(lo, hi)(rng); };
    std::vector<float> in(size_t(24) * n);
    const float h = 2.0f / 48.0f;
    for (unsigned i = 0; i < n; ++i) {
        float *p = in.data() + size_t(24) * i;
        const int mode = i % 4;
        for (int k = 0; k < 4; ++k)
            for (int d = 0; d < 3; ++d) {
                float v;
                if (mode == 0) v = uni(-1.f, 1.f);
                else if (mode == 1) v = uni(-0.05f, 0.05f);
                else { // lattice points (regular-mesh-like symmetric configurations) with optional tiny noise
                    v = h * float(int(uni(-3.f, 3.99f))) + (d == 2 ? 0.3f * h * float(k == 0) : 0.0f);
                    if (mode == 3) v += uni(-1e-6f, 1e-6f);
                }
                p[3 * k + d] = v;
            }
        for (int k = 0; k < 12; ++k) p[12 + k] = p[k] + uni(-0.05f, 0.05f) * (mode == 1 ? 0.1f : 1.0f);
    }
    float *d_in, *d_r, *d_m;
    cudaMalloc(&d_in, in.size() * 4); cudaMalloc(&d_r, size_t(16) * n * 4); cudaMalloc(&d_m, size_t(16) * n * 4);
    cudaMemcpy(d_in, in.data(), in.size() * 4, cudaMemcpyHostToDevice);
    int bad = 0;
    { // distance (uses the first 12 floats of every 24)
        std::vector<float> in12(size_t(12) * n);
        for (unsigned i = 0; i < n; ++i) memcpy(&in12[size_t(12) * i], &in[size_t(24) * i], 48);
        float *d12;
        cudaMalloc(&d12, in12.size() * 4);
        cudaMemcpy(d12, in12.data(), in12.size() * 4, cudaMemcpyHostToDevice);
        k_distance<<<(n + 127) / 128, 128>>>(d12, n, d_r, d_m);
        std::vector<float> r(size_t(16) * n), m(size_t(16) * n);
        cudaMemcpy(r.data(), d_r, r.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(m.data(), d_m, m.size() * 4, cudaMemcpyDeviceToHost);
        int mism[16] = {0};
        for (size_t k = 0; k < r.size(); ++k)
            if (!same_bits(r[k], m[k])) { ++mism[k % 16]; ++bad; }
        printf("distance mismatches per output slot:");
        for (int k = 0; k < 16; ++k) printf(" %d", mism[k]);
        printf("  (n=%u)\n", n);
    }
    for (int pass = 0; pass < 2; ++pass) {
        const float offset = pass ? 1e-3f : 0.0f;
        k_ccd<<<(n + 127) / 128, 128>>>(d_in, n, offset, 1e-3f, d_r, d_m);
        std::vector<float> r(size_t(4) * n), m(size_t(4) * n);
        cudaMemcpy(r.data(), d_r, r.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(m.data(), d_m, m.size() * 4, cudaMemcpyDeviceToHost);
        int mism[4] = {0}, hits[4] = {0};
        for (size_t k = 0; k < r.size(); ++k) {
            if (!same_bits(r[k], m[k])) { ++mism[k % 4]; ++bad; }
            if (r[k] < 1.25f) ++hits[k % 4];
        }
        printf("ccd offset=%g mismatches (pt ee pe pp): %d %d %d %d   hits: %d %d %d %d\n", offset, mism[0], mism[1], mism[2], mism[3], hits[0], hits[1], hits[2], hits[3]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 3; }
    printf("%s (%d mismatching values)\n", bad ? "MISMATCH" : "OK", bad);
    return bad ? 1 : 0;
}

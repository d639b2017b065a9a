// vecmath.cuh -- minimal fp32 3-vector / 3x3 helpers shared by every kernel.
//
// All arithmetic is fp32 and every helper spells out ONE fixed evaluation order, because
// contact-pair parity with the reference is decided by fp32 predicates:
// the order below is the order the reference's SMat templates evaluate in
// (crates/ppf-cts-solver/src/cpp/linalg/smat.hpp:330-362 dot/cross/squaredNorm/normalize),
// so that the compiler's FMA contraction sees the same expression trees.
//
// Every function is PPF_HD so the same source builds for the device (product) and, in
// tests only, for the host (checked against the reference headers compiled for the host).
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define PPF_HD __host__ __device__ __forceinline__
#define PPF_D __device__ __forceinline__
#else
#define PPF_HD inline
#define PPF_D inline
#endif

namespace ppf {

struct V3 {
    float x, y, z;
};

PPF_HD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PPF_HD V3 operator+(const V3 &a, const V3 &b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
PPF_HD V3 operator-(const V3 &a, const V3 &b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
PPF_HD V3 operator-(const V3 &a) { return mk(-a.x, -a.y, -a.z); }
PPF_HD V3 operator*(float s, const V3 &a) { return mk(s * a.x, s * a.y, s * a.z); }
PPF_HD V3 operator*(const V3 &a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
PPF_HD V3 operator/(const V3 &a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
PPF_HD V3 &operator+=(V3 &a, const V3 &b) { a.x = a.x + b.x; a.y = a.y + b.y; a.z = a.z + b.z; return a; }
PPF_HD V3 &operator-=(V3 &a, const V3 &b) { a.x = a.x - b.x; a.y = a.y - b.y; a.z = a.z - b.z; return a; }
// NOTE: the accumulator starts from an explicit 0.0f exactly like SMat::dot (smat.hpp:330-335): with FMA
// contraction the first product is then ROUNDED (fma(a0,b0,0)) and the later ones are fused, which is a
// different rounding from `a0*b0 + a1*b1 + ...` (where the compiler fuses the first product instead).
// On the device the contraction is PINNED with explicit intrinsics: nvcc emits mul.f32 / add.f32 without a rounding
// modifier and ptxas decides per call site what to fuse, so the same source line can round differently in two kernels
// (round 2: one edge-edge pair in 380 000 of the twist scene flipped on exactly that).  The pinned shapes are the ones
// the reference's expressions get by default (checked bit for bit on a B200 by oracle/ref_device_check.cu).
PPF_HD float dot(const V3 &a, const V3 &b) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x)));
#else
    float acc = 0.0f;
    acc = acc + a.x * b.x;
    acc = acc + a.y * b.y;
    acc = acc + a.z * b.z;
    return acc;
#endif
}
// a x + b y: the first product fused into the sum, the second one rounded (LHS-first contraction of fadd(fmul, fmul))
PPF_HD float lin2(float a, float x, float b, float y) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, x, __fmul_rn(b, y));
#else
    return a * x + b * y;
#endif
}
// (a x + b y) + c z
PPF_HD float lin3(float a, float x, float b, float y, float c, float z) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(c, z, __fmaf_rn(a, x, __fmul_rn(b, y)));
#else
    return a * x + b * y + c * z;
#endif
}
PPF_HD V3 cross(const V3 &a, const V3 &b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PPF_HD float sqnorm(const V3 &a) { return dot(a, a); }
PPF_HD V3 lin2(float a, const V3 &x, float b, const V3 &y) { return mk(lin2(a, x.x, b, y.x), lin2(a, x.y, b, y.y), lin2(a, x.z, b, y.z)); }
PPF_HD V3 lin3(float a, const V3 &x, float b, const V3 &y, float c, const V3 &z) {
    return mk(lin3(a, x.x, b, y.x, c, z.x), lin3(a, x.y, b, y.y, c, z.y), lin3(a, x.z, b, y.z, c, z.z));
}
PPF_HD float norm(const V3 &a) { return sqrtf(sqnorm(a)); }
PPF_HD V3 normalized(const V3 &a) {
    const float n = norm(a);
    return n > 0.0f ? mk(a.x / n, a.y / n, a.z / n) : a;
}
PPF_HD float comp(const V3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
PPF_HD float min3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
PPF_HD float max3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

// 3x3 matrix, column-major m[c*3+r] (same byte layout as the ABI's Mat3x3f)
struct M3 {
    float m[9];
};
PPF_HD M3 m3_zero() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.0f; return r; }
PPF_HD M3 m3_identity() { M3 r = m3_zero(); r.m[0] = r.m[4] = r.m[8] = 1.0f; return r; }
PPF_HD M3 m3_outer(const V3 &a, const V3 &b) { // a b^T
    M3 r;
    r.m[0] = a.x * b.x; r.m[1] = a.y * b.x; r.m[2] = a.z * b.x;
    r.m[3] = a.x * b.y; r.m[4] = a.y * b.y; r.m[5] = a.z * b.y;
    r.m[6] = a.x * b.z; r.m[7] = a.y * b.z; r.m[8] = a.z * b.z;
    return r;
}
PPF_HD M3 operator+(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
PPF_HD M3 operator-(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
PPF_HD M3 operator*(float s, const M3 &a) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = s * a.m[i]; return r; }
PPF_HD M3 &operator+=(M3 &a, const M3 &b) { for (int i = 0; i < 9; ++i) a.m[i] = a.m[i] + b.m[i]; return a; }
// y = A x  (column-major; same accumulation order as UnrolledMat3x3f::operator*, solver.cu:214-220)
PPF_HD V3 mul(const M3 &a, const V3 &x) {
    return mk(a.m[0] * x.x + a.m[3] * x.y + a.m[6] * x.z, a.m[1] * x.x + a.m[4] * x.y + a.m[7] * x.z,
              a.m[2] * x.x + a.m[5] * x.y + a.m[8] * x.z);
}
// y = A^T x (UnrolledMat3x3f::operator^, solver.cu:221-227)
PPF_HD V3 mul_t(const M3 &a, const V3 &x) {
    return mk(a.m[0] * x.x + a.m[1] * x.y + a.m[2] * x.z, a.m[3] * x.x + a.m[4] * x.y + a.m[5] * x.z,
              a.m[6] * x.x + a.m[7] * x.y + a.m[8] * x.z);
}
PPF_HD M3 transpose(const M3 &a) {
    M3 r;
    for (int c = 0; c < 3; ++c)
        for (int rr = 0; rr < 3; ++rr) r.m[c + rr * 3] = a.m[rr + c * 3];
    return r;
}
// C = A B, accumulation order of smat.hpp operator* (k outer per column)
PPF_HD M3 matmul(const M3 &a, const M3 &b) {
    M3 r = m3_zero();
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 3; ++k) {
            const float bkc = b.m[k + c * 3];
            for (int rr = 0; rr < 3; ++rr) r.m[rr + c * 3] = r.m[rr + c * 3] + a.m[rr + k * 3] * bkc;
        }
    return r;
}
PPF_HD float det3(const M3 &a) { // smat.hpp determinant()
    const float *m = a.m;
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
}

} // namespace ppf

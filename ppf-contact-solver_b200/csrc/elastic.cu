// elastic.cu -- inertia/air and FEM (membrane + hinge bending) gradient/Hessian assembly into the
// topology-fixed 3x3 block matrix.
//
// Replaces energy::embed_momentum_force_hessian (crates/ppf-cts-solver/src/cpp/energy/energy.cu:700-716,
// per-vertex body 182-270) and energy::embed_elastic_force_hessian (energy.cu:718-794; face body
// 304-364, hinge body 448-553, rod stretch 272-302, rod bending 556-641).  One thread per element; the per-element math lives in
// math_elastic.cuh / math_hinge.cuh; blocks are scattered through the DataSet's precomputed slot
// tables (builder.rs:1368-1496) with fp32 atomics exactly like the reference's push_at
// (csrmat.cu:586-594), lower-triangle sentinels skipped.
#include "backend.hpp"
#include "math_elastic.cuh"
#include "math_damping.cuh"
#include "math_hinge.cuh"
#include "math_inflate.cuh"
#include "math_plasticity.cuh"
#include "math_rod.cuh"
#include "math_stitch.cuh"
#include "math_strain.cuh"
#include "math_torque.cuh"

namespace ppf {

__device__ __forceinline__ V3 ld3(const float4 *p, unsigned i) { const float4 v = p[i]; return mk(v.x, v.y, v.z); }
__device__ __forceinline__ void atomic_add3(float *dst, const V3 &v) {
    if (v.x) atomicAdd(dst + 0, v.x);
    if (v.y) atomicAdd(dst + 1, v.y);
    if (v.z) atomicAdd(dst + 2, v.z);
}
__device__ __forceinline__ void atomic_add_block(float *val, unsigned slot, const float *m) {
    if (slot == 0xFFFFFFFFu) return;
    float *dst = val + size_t(9) * slot;
#pragma unroll
    for (int i = 0; i < 9; ++i)
        if (m[i]) atomicAdd(dst + i, m[i]);
}

struct MomentumArgs {
    float dt, air_density, air_friction, iso_air, wind_w, fix_xz;
    float3 wind;
    int inactive_momentum;
};

__global__ void k_momentum(unsigned nv, const float4 *__restrict__ curr, const float4 *__restrict__ eval_x,
                           const float4 *__restrict__ target, const float4 *__restrict__ vcp, const float *__restrict__ varea,
                           const unsigned *__restrict__ fix_index, const unsigned *__restrict__ pull_index,
                           const PpfPullPair *__restrict__ pull, const unsigned *__restrict__ vf_off,
                           const unsigned *__restrict__ vf_idx, const uint4 *__restrict__ face, MomentumArgs a,
                           float *__restrict__ force, float *__restrict__ diag) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv || fix_index[i] != 0) return;
    const float mass = vcp[i].w, area = varea[i];
    const V3 x = ld3(curr, i), y = ld3(eval_x, i);
    V3 normal = mk(0.f, 0.f, 0.f);
    for (unsigned k = vf_off[i]; k < vf_off[i + 1]; ++k) {
        const uint4 f = face[vf_idx[k]];
        const V3 z0 = ld3(eval_x, f.x), z1 = ld3(eval_x, f.y), z2 = ld3(eval_x, f.z);
        normal += cross(z1 - z0, z2 - z0);
    }
    if (sqnorm(normal)) normal = normalized(normal);
    V3 f = mk(0.f, 0.f, 0.f);
    M3 H = m3_zero();
    const float idt2 = 1.0f / (a.dt * a.dt);
    const bool normal_nonzero = !(fabsf(normal.x) <= 1e-5f && fabsf(normal.y) <= 1e-5f && fabsf(normal.z) <= 1e-5f);
    if (normal_nonzero && a.air_density) {
        V3 wind = mk(0.f, 0.f, 0.f);
        if (!a.inactive_momentum) wind = a.wind_w * mk(a.wind.x, a.wind.y, a.wind.z);
        const V3 z = (y - x) - a.dt * wind;
        const M3 nn = m3_outer(normal, normal);
        const M3 P = m3_identity() - nn;
        const float c = area * a.air_density;
        f += c * ((normal * dot(normal, z)) * idt2 + (a.air_friction * idt2) * mul(P, z));
        H += c * (idt2 * nn + (a.air_friction * idt2) * P);
    }
    const unsigned pi = pull_index[i];
    if (pi) {
        const PpfPullPair pp = pull[pi - 1];
        f += pp.weight * (y - mk(pp.position.v[0], pp.position.v[1], pp.position.v[2]));
        H += pp.weight * m3_identity();
    } else {
        f += (mass * idt2) * (y - ld3(target, i));
        H += (mass * idt2) * m3_identity();
    }
    if (a.iso_air) {
        f += (a.iso_air * idt2) * (y - x);
        H += (a.iso_air * idt2) * m3_identity();
    }
    if (a.fix_xz && y.y > a.fix_xz) { // ParamSet.fix_xz: drag on the horizontal motion above the height fix_xz (energy.cu:260-266)
        const float tm = fminf(1.0f, y.y - a.fix_xz) * mass;
        f.x += (tm * (y.x - x.x)) / (a.dt * a.dt);
        f.z += (tm * (y.z - x.z)) / (a.dt * a.dt);
        H.m[0] += tm * (a.dt * a.dt);
        H.m[8] += tm * (a.dt * a.dt);
    }
    float *fo = force + size_t(3) * i; // one thread owns vertex i in this kernel
    fo[0] += f.x; fo[1] += f.y; fo[2] += f.z;
    float *dg = diag + size_t(9) * i;
#pragma unroll
    for (int k = 0; k < 9; ++k) dg[k] += H.m[k];
}

__global__ void __launch_bounds__(128) k_face(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ eval_x,
                                              const float4 *__restrict__ fmat, const float4 *__restrict__ inv_rest,
                                              const unsigned char *__restrict__ ffixed, const unsigned char *__restrict__ fskip,
                                              const unsigned *__restrict__ slots, float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || ffixed[i] || fskip[i]) return;
    const float4 m = fmat[i];
    if (!(m.x > 0.0f) || (int)m.w != (int)PPF_MODEL_BARAFFWITKIN) return; // the SVD-based models run in k_face_iso
    const uint4 f = face[i];
    const float4 ir = inv_rest[i];
    const float irv[4] = {ir.x, ir.y, ir.z, ir.w};
    FaceElastic fe;
    bw_face(ld3(eval_x, f.x), ld3(eval_x, f.y), ld3(eval_x, f.z), irv, m.x, m.y, m.z, fe);
    const unsigned vi[3] = {f.x, f.y, f.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], fe.f[a]);
    const unsigned *sl = slots + size_t(9) * i;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) atomic_add_block(val, sl[a * 3 + b], fe.H[a][b].m);
}

// inflation pressure on shell faces (energy.cu:366-392; math_inflate.cuh), same face gate as the membrane term
__global__ void __launch_bounds__(128) k_face_pressure(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ eval_x,
                                                       const float *__restrict__ pressure, const unsigned char *__restrict__ ffixed,
                                                       const unsigned char *__restrict__ fskip, const unsigned *__restrict__ slots,
                                                       float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || ffixed[i] || fskip[i]) return;
    const float P = pressure[i];
    if (!(P > 0.0f)) return;
    const uint4 f = face[i];
    const V3 x0 = ld3(eval_x, f.x), x1 = ld3(eval_x, f.y), x2 = ld3(eval_x, f.z);
    V3 g[3];
    inflate_gradient(P, x0, x1 - x0, x2 - x0, g);
    float H[81];
    inflate_hessian(P, x0, x1, x2, H);
    const unsigned vi[3] = {f.x, f.y, f.z};
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], g[a]);
    const unsigned *sl = slots + size_t(9) * i;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            float B[9];
            for (int cc = 0; cc < 3; ++cc)
                for (int r = 0; r < 3; ++r) B[cc * 3 + r] = H[(3 * b + cc) * 9 + 3 * a + r];
            atomic_add_block(val, sl[a * 3 + b], B);
        }
}

// shell faces with an SVD-based model: ARAP / StVK / SNHk (energy.cu:337-352; math_strain.cuh iso_face)
__global__ void __launch_bounds__(128) k_face_iso(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ eval_x,
                                                  const float4 *__restrict__ fmat, const float4 *__restrict__ inv_rest,
                                                  const unsigned char *__restrict__ ffixed, const unsigned char *__restrict__ fskip,
                                                  const unsigned *__restrict__ slots, float eig_eps, float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || ffixed[i] || fskip[i]) return;
    const float4 m = fmat[i];
    const int model = (int)m.w;
    if (!(m.x > 0.0f) || model == (int)PPF_MODEL_BARAFFWITKIN) return;
    const uint4 f = face[i];
    const float4 ir = inv_rest[i];
    const float irv[4] = {ir.x, ir.y, ir.z, ir.w};
    StrainFace sf;
    iso_face(ld3(eval_x, f.x), ld3(eval_x, f.y), ld3(eval_x, f.z), irv, model, m.x, m.y, m.z, eig_eps, sf);
    const unsigned vi[3] = {f.x, f.y, f.z};
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], sf.f[a]);
    const unsigned *sl = slots + size_t(9) * i;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) atomic_add_block(val, sl[a * 3 + b], sf.H[a][b].m);
}

__global__ void __launch_bounds__(128) k_hinge(unsigned nh, const uint4 *__restrict__ hinge, const float4 *__restrict__ eval_x,
                                               const float4 *__restrict__ hmat, const float *__restrict__ stiff,
                                               const unsigned char *__restrict__ hfixed, const unsigned *__restrict__ slots,
                                               float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nh || hfixed[i]) return;
    const float k = stiff[i];
    if (!(k > 0.0f)) return;
    const uint4 h = hinge[i];
    const unsigned vi[4] = {h.z, h.y, h.x, h.w}; // remap (dihedral_angle.hpp:15-17)
    V3 X[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) X[a] = ld3(eval_x, vi[a]);
    HingeElastic he;
    hinge_force_hessian(X, hmat[i].z, he);
#pragma unroll
    for (int a = 0; a < 4; ++a) atomic_add3(force + size_t(3) * vi[a], k * he.f[a]);
    const unsigned *sl = slots + size_t(16) * i;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            const unsigned slot = sl[a * 4 + b];
            if (slot == 0xFFFFFFFFu) continue;
            float blk[9];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) blk[c * 3 + r] = k * he.H[(3 * b + c) * 12 + (3 * a + r)];
            atomic_add_block(val, slot, blk);
        }
}

// ---- rods (energy.cu:272-302 stretch, 556-641 bending; math_rod.cuh) ----
__global__ void __launch_bounds__(128) k_rod(unsigned nrod, const uint2 *__restrict__ edge, const float4 *__restrict__ eval_x,
                                             const float2 *__restrict__ rod_par, const unsigned char *__restrict__ efixed,
                                             const unsigned *__restrict__ slots, float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrod || efixed[i]) return;
    const float2 par = rod_par[i]; // (rest length, stiffness * mass)
    if (!(par.y > 0.0f)) return;
    const uint2 e = edge[i];
    RodElastic re;
    rod_hook(ld3(eval_x, e.x), ld3(eval_x, e.y), par.x, par.y, re);
    atomic_add3(force + size_t(3) * e.x, re.f[0]);
    atomic_add3(force + size_t(3) * e.y, re.f[1]);
    float neg[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) neg[k] = -re.K.m[k];
    const unsigned *sl = slots + size_t(4) * i;
    atomic_add_block(val, sl[0], re.K.m);
    atomic_add_block(val, sl[1], neg);
    atomic_add_block(val, sl[2], neg);
    atomic_add_block(val, sl[3], re.K.m);
}

__global__ void __launch_bounds__(128) k_rod_bend(unsigned n, const uint4 *__restrict__ elem, const float2 *__restrict__ par,
                                                  const float4 *__restrict__ eval_x, const unsigned *__restrict__ fix_index,
                                                  const unsigned *__restrict__ slots, float *__restrict__ force, float *__restrict__ val) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 el = elem[t]; // (j, i, k): i is the bending vertex
    if (fix_index[el.y] != 0) return;
    const float2 pr = par[t]; // (coefficient, rest angle)
    StrandElastic se;
    strand_force_hessian(ld3(eval_x, el.x), ld3(eval_x, el.y), ld3(eval_x, el.z), pr.y, se);
    const unsigned vi[3] = {el.x, el.y, el.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], pr.x * se.f[a]);
    const unsigned *sl = slots + size_t(9) * t;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            const unsigned slot = sl[a * 3 + b];
            if (slot == 0xFFFFFFFFu) continue;
            float blk[9];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) blk[c * 3 + r] = pr.x * se.H[(3 * b + c) * 9 + (3 * a + r)];
            atomic_add_block(val, slot, blk);
        }
}

// ---- strain limiting (strainlimiting.cu:46-120, 250-307; math_strain.cuh) ----
__device__ __forceinline__ unsigned fixed_slot(const unsigned *__restrict__ off, const unsigned *__restrict__ col, unsigned i, unsigned j) { // i <= j
    for (unsigned k = off[i]; k < off[i + 1]; ++k) {
        const unsigned c = col[k];
        if (c == j) return k;
        if (c > j) break;
    }
    return 0xFFFFFFFFu;
}
__global__ void __launch_bounds__(128) k_strainlimit(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ eval_x,
                                                     const float4 *__restrict__ curr, const float4 *__restrict__ fmat,
                                                     const float2 *__restrict__ fstrain, const float4 *__restrict__ inv_rest,
                                                     const unsigned char *__restrict__ ffixed, const unsigned char *__restrict__ fskip,
                                                     const unsigned *__restrict__ fx_off, const unsigned *__restrict__ fx_col,
                                                     const float *__restrict__ fx_el, const unsigned *__restrict__ slots, float eig_eps, int barrier,
                                                     float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || ffixed[i] || fskip[i]) return;
    const float2 sl2 = fstrain[i];
    if (!(sl2.x > 0.0f)) return;
    const uint4 f = face[i];
    const unsigned vi[3] = {f.x, f.y, f.z};
    const float4 ir = inv_rest[i];
    const float irv[4] = {ir.x, ir.y, ir.z, ir.w};
    V3 xe[3], xc[3];
    for (int a = 0; a < 3; ++a) { xe[a] = ld3(eval_x, vi[a]); xc[a] = ld3(curr, vi[a]); }
    // cheap exit before gathering the nine elastic blocks: is the face stretched at all?
    {
        float F[3][2];
        deformation_grad32(xe[0], xe[1], xe[2], irv, F);
        Svd32 svd;
        jacobi_svd32(F, svd);
        if (!(fmaxf(svd.S[0], svd.S[1]) - 1.0f > 0.0f)) return;
    }
    M3 Hel[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            const unsigned ia = vi[a], ib = vi[b];
            const bool swap = ia > ib;
            const unsigned slot = fixed_slot(fx_off, fx_col, swap ? ib : ia, swap ? ia : ib);
            M3 m = m3_zero();
            if (slot != 0xFFFFFFFFu) {
                const float *p = fx_el + size_t(9) * slot;
                if (!swap) for (int e = 0; e < 9; ++e) m.m[e] = p[e];
                else for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) m.m[c * 3 + r] = p[r * 3 + c];
            }
            Hel[a][b] = m;
        }
    StrainFace sf;
    if (!strain_face(xe, xc, irv, sl2.x, sl2.y, fmat[i].z, Hel, eig_eps, sf, barrier)) return;
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], sf.f[a]);
    const unsigned *sl = slots + size_t(9) * i;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) atomic_add_block(val, sl[a * 3 + b], sf.H[a][b].m);
}
// per-face largest admissible step fraction; folded into *toi_bits (positive floats order like their bit patterns)
__global__ void __launch_bounds__(128) k_strain_toi(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ x0,
                                                    const float4 *__restrict__ x1, const float2 *__restrict__ fstrain,
                                                    const float4 *__restrict__ inv_rest, const unsigned char *__restrict__ ffixed,
                                                    const unsigned char *__restrict__ fskip, float max_t, unsigned *toi_bits) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    float t = max_t;
    if (i < nf && !ffixed[i] && !fskip[i]) {
        const float2 sl2 = fstrain[i];
        if (sl2.x > 0.0f) {
            const uint4 f = face[i];
            const float4 ir = inv_rest[i];
            const float irv[4] = {ir.x, ir.y, ir.z, ir.w};
            const V3 xa[3] = {ld3(x0, f.x), ld3(x0, f.y), ld3(x0, f.z)}, xb[3] = {ld3(x1, f.x), ld3(x1, f.y), ld3(x1, f.z)};
            t = strain_face_toi(xa, xb, irv, sl2.y, max_t);
        }
    }
    for (int o = 16; o > 0; o >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if ((threadIdx.x & 31) == 0 && t < max_t) atomicMin(toi_bits, __float_as_uint(fmaxf(t, 0.0f)));
}

// ---- plastic creep of the rest shape after a step (plasticity.cu; math_plasticity.cuh) ----
__global__ void __launch_bounds__(128) k_plastic_face(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ curr,
                                                      const float2 *__restrict__ fplast, float dt, float4 *__restrict__ inv_rest) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const float2 pl = fplast[i]; // (rate, threshold)
    if (!(pl.x > 0.0f)) return;
    const uint4 f = face[i];
    const float4 ir = inv_rest[i];
    float irv[4] = {ir.x, ir.y, ir.z, ir.w};
    if (face_plasticity(ld3(curr, f.x), ld3(curr, f.y), ld3(curr, f.z), irv, pl.y, plastic_alpha(pl.x, dt)))
        inv_rest[i] = make_float4(irv[0], irv[1], irv[2], irv[3]);
}
__global__ void __launch_bounds__(128) k_plastic_hinge(unsigned nh, const uint4 *__restrict__ hinge, const float4 *__restrict__ curr,
                                                       const float2 *__restrict__ hplast, const unsigned char *__restrict__ hfixed, float dt,
                                                       float4 *__restrict__ hmat) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nh || hfixed[i]) return;
    const float2 pl = hplast[i];
    if (!(pl.x > 0.0f)) return;
    const uint4 h = hinge[i];
    const V3 X[4] = {ld3(curr, h.z), ld3(curr, h.y), ld3(curr, h.x), ld3(curr, h.w)}; // remapped (dihedral_angle.hpp:15-17)
    float4 m = hmat[i]; // (length, area, rest angle, bend)
    m.z = hinge_plasticity(X, m.z, pl.y, plastic_alpha(pl.x, dt));
    hmat[i] = m;
}
__global__ void __launch_bounds__(128) k_plastic_rod_bend(unsigned n, const uint4 *__restrict__ elem, const float4 *__restrict__ curr,
                                                          const float2 *__restrict__ rb_plast, float dt, float2 *__restrict__ par) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float2 pl = rb_plast[t];
    if (!(pl.x > 0.0f)) return;
    const uint4 el = elem[t]; // (j, i, k)
    float2 pr = par[t];       // (coefficient, rest angle)
    pr.y = rod_bend_plasticity(ld3(curr, el.x), ld3(curr, el.y), ld3(curr, el.z), pr.y, pl.y, plastic_alpha(pl.x, dt));
    par[t] = pr;
}
void apply_plasticity(State &s, const PpfParamSet &prm) {
    cudaStream_t q = s.stream;
    if (s.has_fplast && s.nf) { k_plastic_face<<<div_up(s.nf, 128), 128, 0, q>>>(s.nf, s.face.p, s.curr.p, s.fplast.p, prm.dt, s.inv_rest.p); LAUNCH_COUNT(1); }
    if (s.has_hplast && s.nh) { k_plastic_hinge<<<div_up(s.nh, 128), 128, 0, q>>>(s.nh, s.hinge.p, s.curr.p, s.hplast.p, s.hfixed.p, prm.dt, s.hmat.p); LAUNCH_COUNT(1); }
    if (s.has_rb_plast && s.nrodbend) { k_plastic_rod_bend<<<div_up(s.nrodbend, 128), 128, 0, q>>>(s.nrodbend, s.rb_elem.p, s.curr.p, s.rb_plast.p, prm.dt, s.rb_par.p); LAUNCH_COUNT(1); }
}

// ---- stitches (energy.cu:796-886; math_stitch.cuh) ----
__global__ void __launch_bounds__(64) k_stitch(unsigned n, const PpfStitch *__restrict__ st, const float4 *__restrict__ eval_x,
                                               const float4 *__restrict__ vcp, const unsigned *__restrict__ slots, float length_factor,
                                               float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PpfStitch q = st[i];
    V3 x[6];
    float ghat[6], offset[6], w[6];
    for (int a = 0; a < 6; ++a) {
        const unsigned v = q.index.v[a];
        x[a] = ld3(eval_x, v);
        const float4 cp = vcp[v];
        ghat[a] = cp.x; offset[a] = cp.y; w[a] = q.weight.v[a];
    }
    StitchOut o;
    stitch_force_hessian(x, ghat, offset, w, q.stiffness, length_factor, o);
    for (int a = 0; a < 6; ++a) atomic_add3(force + size_t(3) * q.index.v[a], o.f[a]);
    const unsigned *sl = slots + size_t(36) * i;
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
            const unsigned slot = sl[a * 6 + b];
            if (slot == 0xFFFFFFFFu) continue;
            float blk[9];
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) blk[c * 3 + r] = o.H[(3 * b + c) * 18 + (3 * a + r)];
            atomic_add_block(val, slot, blk);
        }
}
void assemble_stitches(State &s, const PpfParamSet &prm) {
    if (!s.nstitch) return;
    k_stitch<<<div_up(s.nstitch, 64), 64, 0, s.stream>>>(s.nstitch, s.stitches.p, s.eval_x.p, s.vcp.p, s.stitch_slots.p, prm.stitch_length_factor, s.force.p, s.fx_val.p);
    LAUNCH_COUNT(1);
}

// ---- Rayleigh damping (energy.cu:88-180; math_damping.cuh) ----
// The element kernels above stay as they are; these kernels add only the damping DELTA of the damped elements:
//   membrane / rod stretch:  gradient += s K u,      Hessian += s K       (the reference scales K by (1 + s) in place)
//   hinge / rod bending:     gradient += s K_lag u,  Hessian += s K_lag   (K_lag = element Hessian at the start-of-step positions)
// with s = beta / dt and u = eval_x - curr.
__global__ void __launch_bounds__(128) k_damp_face(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ eval_x,
                                                   const float4 *__restrict__ curr, const float4 *__restrict__ fmat,
                                                   const float4 *__restrict__ inv_rest, const float *__restrict__ fdamp,
                                                   const unsigned char *__restrict__ ffixed, const unsigned char *__restrict__ fskip,
                                                   const unsigned *__restrict__ slots, float dt, float eig_eps, float *__restrict__ force,
                                                   float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf || ffixed[i] || fskip[i]) return;
    const float beta = fdamp[i];
    const float4 m = fmat[i];
    if (!(beta > 0.0f) || !(m.x > 0.0f)) return;
    const uint4 f = face[i];
    const unsigned vi[3] = {f.x, f.y, f.z};
    const float4 ir = inv_rest[i];
    const float irv[4] = {ir.x, ir.y, ir.z, ir.w};
    float K[81];
    if ((int)m.w == (int)PPF_MODEL_BARAFFWITKIN) {
        FaceElastic fe;
        bw_face(ld3(eval_x, f.x), ld3(eval_x, f.y), ld3(eval_x, f.z), irv, m.x, m.y, m.z, fe);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) K[(3 * b + c) * 9 + 3 * a + r] = fe.H[a][b].m[c * 3 + r];
    } else {
        StrainFace sf;
        iso_face(ld3(eval_x, f.x), ld3(eval_x, f.y), ld3(eval_x, f.z), irv, (int)m.w, m.x, m.y, m.z, eig_eps, sf);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) K[(3 * b + c) * 9 + 3 * a + r] = sf.H[a][b].m[c * 3 + r];
    }
    V3 u[3], df[3];
    for (int a = 0; a < 3; ++a) { u[a] = ld3(eval_x, vi[a]) - ld3(curr, vi[a]); df[a] = mk(0.f, 0.f, 0.f); }
    const float s = beta / dt;
    damping_gradient<3>(K, u, s, df);
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], df[a]);
    const unsigned *sl = slots + size_t(9) * i;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            if (sl[a * 3 + b] == 0xFFFFFFFFu) continue;
            float blk[9];
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) blk[c * 3 + r] = s * K[(3 * b + c) * 9 + 3 * a + r];
            atomic_add_block(val, sl[a * 3 + b], blk);
        }
}
__global__ void __launch_bounds__(128) k_damp_rod(unsigned nrod, const uint2 *__restrict__ edge, const float4 *__restrict__ eval_x,
                                                  const float4 *__restrict__ curr, const float2 *__restrict__ rod_par,
                                                  const float *__restrict__ rod_damp, const unsigned char *__restrict__ efixed,
                                                  const unsigned *__restrict__ slots, float dt, float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrod || efixed[i]) return;
    const float beta = rod_damp[i];
    const float2 par = rod_par[i];
    if (!(beta > 0.0f) || !(par.y > 0.0f)) return;
    const uint2 e = edge[i];
    RodElastic re;
    rod_hook(ld3(eval_x, e.x), ld3(eval_x, e.y), par.x, par.y, re);
    float K[36];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r)
        K[(3 * b + c) * 6 + 3 * a + r] = (a == b ? 1.0f : -1.0f) * re.K.m[c * 3 + r];
    const V3 u[2] = {ld3(eval_x, e.x) - ld3(curr, e.x), ld3(eval_x, e.y) - ld3(curr, e.y)};
    V3 df[2] = {mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 0.f)};
    const float s = beta / dt;
    damping_gradient<2>(K, u, s, df);
    atomic_add3(force + size_t(3) * e.x, df[0]);
    atomic_add3(force + size_t(3) * e.y, df[1]);
    float pos[9], neg[9];
    for (int k = 0; k < 9; ++k) { pos[k] = s * re.K.m[k]; neg[k] = -pos[k]; }
    const unsigned *sl = slots + size_t(4) * i;
    atomic_add_block(val, sl[0], pos);
    atomic_add_block(val, sl[1], neg);
    atomic_add_block(val, sl[2], neg);
    atomic_add_block(val, sl[3], pos);
}
__global__ void __launch_bounds__(128) k_damp_hinge(unsigned nh, const uint4 *__restrict__ hinge, const float4 *__restrict__ eval_x,
                                                    const float4 *__restrict__ curr, const float4 *__restrict__ hmat,
                                                    const float *__restrict__ stiff, const float *__restrict__ hdamp,
                                                    const unsigned char *__restrict__ hfixed, const unsigned *__restrict__ slots, float dt,
                                                    float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nh || hfixed[i]) return;
    const float beta = hdamp[i], k = stiff[i];
    if (!(beta > 0.0f) || !(k > 0.0f)) return;
    const uint4 h = hinge[i];
    const unsigned vi[4] = {h.z, h.y, h.x, h.w};
    V3 X[4], u[4], df[4];
    for (int a = 0; a < 4; ++a) { X[a] = ld3(curr, vi[a]); u[a] = ld3(eval_x, vi[a]) - X[a]; df[a] = mk(0.f, 0.f, 0.f); }
    HingeElastic he; // at the START-OF-STEP positions: K_lag
    hinge_force_hessian(X, hmat[i].z, he);
    const float s = beta / dt;
    damping_gradient<4>(he.H, u, s * k, df); // K_lag = k * H
    for (int a = 0; a < 4; ++a) atomic_add3(force + size_t(3) * vi[a], df[a]);
    const unsigned *sl = slots + size_t(16) * i;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            if (sl[a * 4 + b] == 0xFFFFFFFFu) continue;
            float blk[9];
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) blk[c * 3 + r] = s * (k * he.H[(3 * b + c) * 12 + (3 * a + r)]);
            atomic_add_block(val, sl[a * 4 + b], blk);
        }
}
__global__ void __launch_bounds__(128) k_damp_rod_bend(unsigned n, const uint4 *__restrict__ elem, const float2 *__restrict__ par,
                                                       const float *__restrict__ rb_damp, const float4 *__restrict__ eval_x,
                                                       const float4 *__restrict__ curr, const unsigned *__restrict__ fix_index,
                                                       const unsigned *__restrict__ slots, float dt, float *__restrict__ force,
                                                       float *__restrict__ val) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 el = elem[t];
    const float beta = rb_damp[t];
    if (fix_index[el.y] != 0 || !(beta > 0.0f)) return;
    const float2 pr = par[t];
    const unsigned vi[3] = {el.x, el.y, el.z};
    V3 X[3], u[3], df[3];
    for (int a = 0; a < 3; ++a) { X[a] = ld3(curr, vi[a]); u[a] = ld3(eval_x, vi[a]) - X[a]; df[a] = mk(0.f, 0.f, 0.f); }
    StrandElastic se; // K_lag
    strand_force_hessian(X[0], X[1], X[2], pr.y, se);
    const float s = beta / dt;
    damping_gradient<3>(se.H, u, s * pr.x, df);
    for (int a = 0; a < 3; ++a) atomic_add3(force + size_t(3) * vi[a], df[a]);
    const unsigned *sl = slots + size_t(9) * t;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            if (sl[a * 3 + b] == 0xFFFFFFFFu) continue;
            float blk[9];
            for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) blk[c * 3 + r] = s * (pr.x * se.H[(3 * b + c) * 9 + (3 * a + r)]);
            atomic_add_block(val, sl[a * 3 + b], blk);
        }
}

void assemble_damping(State &s, float dt, const PpfParamSet &prm) {
    if (!(dt > 0.0f)) return;
    cudaStream_t q = s.stream;
    if (s.has_fdamp && s.nf) {
        k_damp_face<<<div_up(s.nf, 128), 128, 0, q>>>(s.nf, s.face.p, s.eval_x.p, s.curr.p, s.fmat.p, s.inv_rest.p, s.fdamp.p, s.ffixed.p, s.fskip.p, s.face_slots.p,
                                                      dt, prm.eiganalysis_eps, s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
    if (s.has_rod_damp && s.nrod) {
        k_damp_rod<<<div_up(s.nrod, 128), 128, 0, q>>>(s.nrod, s.edge.p, s.eval_x.p, s.curr.p, s.rod_par.p, s.rod_damp.p, s.efixed.p, s.rod_slots.p, dt, s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
    if (s.has_hdamp && s.nh) {
        k_damp_hinge<<<div_up(s.nh, 128), 128, 0, q>>>(s.nh, s.hinge.p, s.eval_x.p, s.curr.p, s.hmat.p, s.hinge_stiff.p, s.hdamp.p, s.hfixed.p, s.hinge_slots.p, dt,
                                                       s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
    if (s.has_rb_damp && s.nrodbend) {
        k_damp_rod_bend<<<div_up(s.nrodbend, 128), 128, 0, q>>>(s.nrodbend, s.rb_elem.p, s.rb_par.p, s.rb_damp.p, s.eval_x.p, s.curr.p, s.fix_index.p, s.rb_slots.p,
                                                                dt, s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
}

// ---- rod strain limiting (strainlimiting.cu:122-245; math_rod.cuh) ----
__device__ __forceinline__ M3 ld_block(const float *__restrict__ val, unsigned slot) {
    M3 m;
    const float *p = val + size_t(9) * slot;
#pragma unroll
    for (int e = 0; e < 9; ++e) m.m[e] = p[e];
    return m;
}
__global__ void __launch_bounds__(128) k_rod_strain(unsigned nrod, const uint2 *__restrict__ edge, const float4 *__restrict__ eval_x,
                                                    const float4 *__restrict__ curr, const float2 *__restrict__ rod_sl,
                                                    const float4 *__restrict__ ecp, const unsigned char *__restrict__ efixed,
                                                    const unsigned *__restrict__ slots, const float *__restrict__ fx_el, int barrier,
                                                    float *__restrict__ force, float *__restrict__ val) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrod || efixed[i]) return;
    const float2 sl2 = rod_sl[i]; // (limit, drawn length)
    if (!(sl2.x > 0.0f)) return;
    const uint2 e = edge[i];
    const V3 x0 = ld3(eval_x, e.x), x1 = ld3(eval_x, e.y);
    if (!(norm(x1 - x0) / sl2.y - 1.0f > 0.0f)) return; // not stretched: skip the block gathers
    const unsigned *sl = slots + size_t(4) * i;
    const M3 H00 = ld_block(fx_el, sl[0]), H11 = ld_block(fx_el, sl[3]);
    const M3 H01 = sl[1] != 0xFFFFFFFFu ? ld_block(fx_el, sl[1]) : transpose(ld_block(fx_el, sl[2]));
    RodElastic re;
    if (!rod_strain(x0, x1, ld3(curr, e.x), ld3(curr, e.y), sl2.y, sl2.x, ecp[i].w, H00, H01, H11, re, barrier)) return;
    atomic_add3(force + size_t(3) * e.x, re.f[0]);
    atomic_add3(force + size_t(3) * e.y, re.f[1]);
    float neg[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) neg[k] = -re.K.m[k];
    atomic_add_block(val, sl[0], re.K.m);
    atomic_add_block(val, sl[1], neg);
    atomic_add_block(val, sl[2], neg);
    atomic_add_block(val, sl[3], re.K.m);
}
__global__ void __launch_bounds__(128) k_rod_strain_toi(unsigned nrod, const uint2 *__restrict__ edge, const float4 *__restrict__ x0,
                                                        const float4 *__restrict__ x1, const float2 *__restrict__ rod_sl,
                                                        const unsigned char *__restrict__ efixed, float max_t, unsigned *toi_bits) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    float t = max_t;
    if (i < nrod && !efixed[i]) {
        const float2 sl2 = rod_sl[i];
        if (sl2.x > 0.0f) {
            const uint2 e = edge[i];
            t = rod_strain_toi(ld3(x0, e.y) - ld3(x0, e.x), ld3(x1, e.y) - ld3(x1, e.x), sl2.y, sl2.x, max_t);
        }
    }
    for (int o = 16; o > 0; o >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, o));
    if ((threadIdx.x & 31) == 0 && t < max_t) atomicMin(toi_bits, __float_as_uint(fmaxf(t, 0.0f)));
}

void assemble_strainlimit(State &s, const PpfParamSet &prm) {
    if (s.has_strainlimit && s.nf) {
        k_strainlimit<<<div_up(s.nf, 128), 128, 0, s.stream>>>(s.nf, s.face.p, s.eval_x.p, s.curr.p, s.fmat.p, s.fstrain.p, s.inv_rest.p, s.ffixed.p, s.fskip.p,
                                                               s.fx_off.p, s.fx_col.p, s.fx_el.p, s.face_slots.p, prm.eiganalysis_eps, (int)prm.barrier, s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
    if (s.has_rod_strainlimit && s.nrod) {
        k_rod_strain<<<div_up(s.nrod, 128), 128, 0, s.stream>>>(s.nrod, s.edge.p, s.eval_x.p, s.curr.p, s.rod_sl.p, s.ecp.p, s.efixed.p, s.rod_slots.p,
                                                                s.fx_el.p, (int)prm.barrier, s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
}
// strainlimiting::rod_line_search: fraction of the step x0 -> x1 that keeps every rod edge below its strain limit
float rod_strain_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm) {
    if (!s.has_rod_strainlimit || !s.nrod) return 1.0f;
    unsigned *w = s.counters.p + 26;
    const float max_t = prm.line_search_max_t;
    unsigned init;
    memcpy(&init, &max_t, sizeof(float));
    CUDA_OK(cudaMemcpyAsync(w, &init, sizeof(unsigned), cudaMemcpyHostToDevice, s.stream));
    k_rod_strain_toi<<<div_up(s.nrod, 128), 128, 0, s.stream>>>(s.nrod, s.edge.p, x0, x1, s.rod_sl.p, s.efixed.p, max_t, w);
    LAUNCH_COUNT(1);
    unsigned h = 0;
    CUDA_OK(cudaMemcpyAsync(&h, w, sizeof(unsigned), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
    float t;
    memcpy(&t, &h, sizeof(float));
    return t / max_t;
}
// strainlimiting::line_search: fraction of the step x0 -> x1 that keeps every face below its strain limit
float strain_line_search(State &s, const float4 *x0, const float4 *x1, const PpfParamSet &prm) {
    if (!s.has_strainlimit || !s.nf) return 1.0f;
    unsigned *w = s.counters.p + 26;
    const float max_t = prm.line_search_max_t;
    unsigned init;
    memcpy(&init, &max_t, sizeof(float));
    CUDA_OK(cudaMemcpyAsync(w, &init, sizeof(unsigned), cudaMemcpyHostToDevice, s.stream));
    k_strain_toi<<<div_up(s.nf, 128), 128, 0, s.stream>>>(s.nf, s.face.p, x0, x1, s.fstrain.p, s.inv_rest.p, s.ffixed.p, s.fskip.p, max_t, w);
    LAUNCH_COUNT(1);
    unsigned h = 0;
    CUDA_OK(cudaMemcpyAsync(&h, w, sizeof(unsigned), cudaMemcpyDeviceToHost, s.stream));
    CUDA_OK(cudaStreamSynchronize(s.stream));
    float t;
    memcpy(&t, &h, sizeof(float));
    return t / max_t;
}

// ---- "max_sigma": largest stretch ratio at the start of the step (main.cu:424-488) ----
// shells: largest singular value of the deformation gradient x min(shrink_x, shrink_y); rods: length / drawn length
__global__ void __launch_bounds__(128) k_face_sigma(unsigned nf, const uint4 *__restrict__ face, const float4 *__restrict__ curr,
                                                    const float4 *__restrict__ inv_rest, const float *__restrict__ fshrink,
                                                    const unsigned char *__restrict__ ffixed, float *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    float v = 0.0f;
    const float shrink = fshrink[i]; // 0 for collider faces: they carry no elastic energy (main.cu:446-449)
    if (!ffixed[i] && shrink > 0.0f) {
        const uint4 f = face[i];
        const float4 ir = inv_rest[i];
        const float irv[4] = {ir.x, ir.y, ir.z, ir.w};
        float F[3][2];
        deformation_grad32(ld3(curr, f.x), ld3(curr, f.y), ld3(curr, f.z), irv, F);
        Svd32 svd;
        jacobi_svd32(F, svd);
        v = fmaxf(svd.S[0], svd.S[1]) * shrink;
    }
    out[i] = v;
}
__global__ void __launch_bounds__(128) k_rod_sigma(unsigned nrod, const uint2 *__restrict__ edge, const float4 *__restrict__ curr,
                                                   const float2 *__restrict__ rod_sl, const unsigned char *__restrict__ efixed, float *__restrict__ out) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrod) return;
    float v = 0.0f;
    const float l0 = rod_sl[i].y; // drawn length
    if (!efixed[i] && l0 > 0.0f) {
        const uint2 e = edge[i];
        v = norm(ld3(curr, e.y) - ld3(curr, e.x)) / l0;
    }
    out[i] = v;
}
float max_stretch_ratio(State &s) {
    float *tmp = reinterpret_cast<float *>(s.velocity.p); // 4 floats per vertex of scratch
    float m = 0.0f;
    if (s.nf && size_t(s.nf) <= size_t(4) * s.nv) {
        k_face_sigma<<<div_up(s.nf, 128), 128, 0, s.stream>>>(s.nf, s.face.p, s.curr.p, s.inv_rest.p, s.fshrink.p, s.ffixed.p, tmp);
        LAUNCH_COUNT(1);
        m = fmaxf(m, reduce_max_f(s, tmp, s.nf, 0.0f));
    }
    if (s.nrod && size_t(s.nrod) <= size_t(4) * s.nv) {
        k_rod_sigma<<<div_up(s.nrod, 128), 128, 0, s.stream>>>(s.nrod, s.edge.p, s.curr.p, s.rod_sl.p, s.efixed.p, tmp);
        LAUNCH_COUNT(1);
        m = fmaxf(m, reduce_max_f(s, tmp, s.nrod, 0.0f));
    }
    return m;
}

// torque pins (energy.cu:643-701 pre-pass, 221-246 per-vertex force): one thread per group walks its vertices (groups are few
// and small: a pinned handle), then one thread per torque vertex adds its force and the PSD part of its Hessian
__global__ void k_torque_groups(unsigned ng, const PpfTorqueGroup *__restrict__ groups, const PpfTorqueVertex *__restrict__ tv,
                                const float4 *__restrict__ eval_x, const float4 *__restrict__ vcp, float *__restrict__ result) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ng) return;
    const TorqueGroupResult r = torque_group(groups[g], tv, [&](unsigned v) { return ld3(eval_x, v); }, [&](unsigned v) { return vcp[v].w; });
    float *o = result + size_t(7) * g;
    o[0] = r.center.x; o[1] = r.center.y; o[2] = r.center.z; o[3] = r.axis.x; o[4] = r.axis.y; o[5] = r.axis.z; o[6] = r.inv_r_perp_sq_sum;
}
__global__ void k_torque_vertices(unsigned n, const PpfTorqueVertex *__restrict__ tv, const float *__restrict__ result,
                                  const float4 *__restrict__ eval_x, const unsigned *__restrict__ fix_index, float *__restrict__ force,
                                  float *__restrict__ diag) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const PpfTorqueVertex t = tv[j];
    if (fix_index[t.index] != 0) return; // the momentum term skips prescribed vertices altogether (energy.cu:709-714)
    const float *o = result + size_t(7) * t.group_id;
    TorqueGroupResult r;
    r.center = mk(o[0], o[1], o[2]); r.axis = mk(o[3], o[4], o[5]); r.inv_r_perp_sq_sum = o[6];
    V3 f;
    M3 H;
    torque_vertex(ld3(eval_x, t.index), t.magnitude, r, f, H);
    float *fo = force + size_t(3) * t.index;
    atomicAdd(fo + 0, f.x); atomicAdd(fo + 1, f.y); atomicAdd(fo + 2, f.z);
    float *dg = diag + size_t(9) * t.index;
    for (int k = 0; k < 9; ++k)
        if (H.m[k]) atomicAdd(dg + k, H.m[k]);
}

void assemble_momentum(State &s, float dt, const PpfParamSet &prm) {
    MomentumArgs a;
    a.dt = dt; a.air_density = prm.air_density; a.air_friction = prm.air_friction; a.iso_air = prm.isotropic_air_friction; a.fix_xz = prm.fix_xz;
    a.wind = make_float3(prm.wind.v[0], prm.wind.v[1], prm.wind.v[2]);
    a.inactive_momentum = prm.inactive_momentum;
    { // utility.cu get_wind_weight (float time)
        const float time = (float)prm.time;
        const float angle = 30.0f * time, t = 0.25f;
        a.wind_w = t * (0.5f * (1.0f + sinf(angle))) + (1.0f - t);
    }
    k_momentum<<<div_up(s.nv, 128), 128, 0, s.stream>>>(s.nv, s.curr.p, s.eval_x.p, s.target.p, s.vcp.p, s.varea.p, s.fix_index.p,
                                                        s.pull_index.p, s.pull.p, s.vf_off.p, s.vf_idx.p, s.face.p, a, s.force.p, s.diag.p);
    LAUNCH_COUNT(1);
    if (s.ntvert) {
        k_torque_groups<<<div_up(s.ntgroup, 64), 64, 0, s.stream>>>(s.ntgroup, s.tgroups.p, s.tverts.p, s.eval_x.p, s.vcp.p, s.tresult.p);
        k_torque_vertices<<<div_up(s.ntvert, 128), 128, 0, s.stream>>>(s.ntvert, s.tverts.p, s.tresult.p, s.eval_x.p, s.fix_index.p, s.force.p, s.diag.p);
        LAUNCH_COUNT(2);
    }
}

void assemble_elastic(State &s, float dt, const PpfParamSet &prm) {
    (void)dt;
    if (s.nrodbend) k_rod_bend<<<div_up(s.nrodbend, 128), 128, 0, s.stream>>>(s.nrodbend, s.rb_elem.p, s.rb_par.p, s.eval_x.p, s.fix_index.p,
                                                                             s.rb_slots.p, s.force.p, s.fx_val.p);
    if (s.nrod) k_rod<<<div_up(s.nrod, 128), 128, 0, s.stream>>>(s.nrod, s.edge.p, s.eval_x.p, s.rod_par.p, s.efixed.p, s.rod_slots.p,
                                                                s.force.p, s.fx_val.p);
    if (s.nrod) LAUNCH_COUNT(2);
    if (s.nf) k_face<<<div_up(s.nf, 128), 128, 0, s.stream>>>(s.nf, s.face.p, s.eval_x.p, s.fmat.p, s.inv_rest.p, s.ffixed.p, s.fskip.p,
                                                             s.face_slots.p, s.force.p, s.fx_val.p);
    if (s.nf && s.has_iso_faces) {
        k_face_iso<<<div_up(s.nf, 128), 128, 0, s.stream>>>(s.nf, s.face.p, s.eval_x.p, s.fmat.p, s.inv_rest.p, s.ffixed.p, s.fskip.p, s.face_slots.p,
                                                            prm.eiganalysis_eps, s.force.p, s.fx_val.p);
        LAUNCH_COUNT(1);
    }
    if (s.nf && s.has_pressure) {
        k_face_pressure<<<div_up(s.nf, 128), 128, 0, s.stream>>>(s.nf, s.face.p, s.eval_x.p, s.fpressure.p, s.ffixed.p, s.fskip.p, s.face_slots.p, s.force.p,
                                                                 s.fx_val.p);
        LAUNCH_COUNT(1);
    }
    if (s.nh) k_hinge<<<div_up(s.nh, 128), 128, 0, s.stream>>>(s.nh, s.hinge.p, s.eval_x.p, s.hmat.p, s.hinge_stiff.p, s.hfixed.p,
                                                              s.hinge_slots.p, s.force.p, s.fx_val.p);
    LAUNCH_COUNT(2);
}

} // namespace ppf

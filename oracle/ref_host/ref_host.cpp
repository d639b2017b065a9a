// ref_host.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Compiles the reference's header-only per-element device math for the HOST, straight
// from /root/reference (no source is copied), and exposes it as a flat C API so the CPU
// test-suite can check our restatements against the real reference arithmetic without a
// GPU.  With a plain host compiler the CUDA qualifiers (__device__, __host__) reduce to
// ignored attributes, so the templates instantiate as ordinary functions.
//
// Built by oracle/Makefile.ref into oracle/_ref/libref_host.so.  Headers used:
//   contact/distance.hpp, contact/accd.hpp, contact/intersect_core.hpp, barrier/cubic.hpp,
//   energy/model/{friction,baraffwitkin,dihedral_angle,hook}.hpp, linalg/eigsolve.hpp
#include <cmath>
#include <cstring>
#include <limits>
using std::max;
using std::min;
using std::isfinite;
#include "data.hpp"
#include "contact/distance.hpp"
#include "contact/accd.hpp"
#include "contact/intersect_core.hpp"
#include "barrier/cubic.hpp"
#include "barrier/quadratic.hpp"
#include "barrier/logarithm.hpp"
#include "energy/model/friction.hpp"
#include "energy/model/baraffwitkin.hpp"
#include "energy/model/dihedral_angle.hpp"
#include "energy/model/hook.hpp"
#include "energy/model/arap.hpp"
#include "energy/model/stvk.hpp"
#include "energy/model/snhk.hpp"
#include "linalg/eigsolve.hpp"

// storage for the device globals accd.hpp declares extern (contact.cu:30-38)
namespace accd {
unsigned g_ccd_overlap = 0u, g_ccd_overlap_v0 = 0xFFFFFFFFu, g_ccd_overlap_v1 = 0xFFFFFFFFu, g_ccd_overlap_kind = 0xFFFFFFFFu;
float g_ccd_overlap_d2 = -1.0f, g_ccd_overlap_offset = -1.0f;
}
unsigned long long g_device_alloc_count = 0, g_device_free_count = 0;

static Vec3f v3(const float *p) { return Vec3f(p[0], p[1], p[2]); }

extern "C" {
void ref_point_triangle_coeff(const float *p, const float *t0, const float *t1, const float *t2, float *out) {
    Vec3f c = distance::point_triangle_distance_coeff<float, float>(v3(p), v3(t0), v3(t1), v3(t2));
    for (int k = 0; k < 3; ++k) out[k] = c[k];
}
void ref_point_triangle_coeff_unclassified(const float *p, const float *t0, const float *t1, const float *t2, float *out) {
    Vec3f c = distance::point_triangle_distance_coeff_unclassified<float, float>(v3(p), v3(t0), v3(t1), v3(t2));
    for (int k = 0; k < 3; ++k) out[k] = c[k];
}
void ref_point_edge_coeff(const float *p, const float *e0, const float *e1, float *out) {
    Vec2f c = distance::point_edge_distance_coeff<float, float>(v3(p), v3(e0), v3(e1));
    out[0] = c[0]; out[1] = c[1];
}
void ref_edge_edge_coeff(const float *a0, const float *a1, const float *b0, const float *b1, float *out) {
    Vec4f c = distance::edge_edge_distance_coeff<float, float>(v3(a0), v3(a1), v3(b0), v3(b1));
    for (int k = 0; k < 4; ++k) out[k] = c[k];
}
static ParamSet mkparam(float max_t, float ccd_eps) {
    ParamSet p;
    std::memset(&p, 0, sizeof(p));
    p.line_search_max_t = max_t;
    p.ccd_eps = ccd_eps;
    return p;
}
// x0: 4 points at t=0 (p,t0,t1,t2), x1: same at t=1
float ref_point_triangle_ccd(const float *x0, const float *x1, float offset, float ghat, float max_t, float ccd_eps, unsigned *overlap) {
    accd::g_ccd_overlap = 0;
    ParamSet prm = mkparam(max_t, ccd_eps);
    float r = accd::point_triangle_ccd(v3(x0), v3(x1), v3(x0 + 3), v3(x0 + 6), v3(x0 + 9), v3(x1 + 3), v3(x1 + 6), v3(x1 + 9), offset, ghat, prm);
    if (overlap) *overlap = accd::g_ccd_overlap;
    return r;
}
float ref_point_edge_ccd(const float *x0, const float *x1, float offset, float ghat, float max_t, float ccd_eps, unsigned *overlap) {
    accd::g_ccd_overlap = 0;
    ParamSet prm = mkparam(max_t, ccd_eps);
    float r = accd::point_edge_ccd(v3(x0), v3(x1), v3(x0 + 3), v3(x0 + 6), v3(x1 + 3), v3(x1 + 6), offset, ghat, prm);
    if (overlap) *overlap = accd::g_ccd_overlap;
    return r;
}
float ref_point_point_ccd(const float *x0, const float *x1, float offset, float ghat, float max_t, float ccd_eps, unsigned *overlap) {
    accd::g_ccd_overlap = 0;
    ParamSet prm = mkparam(max_t, ccd_eps);
    float r = accd::point_point_ccd(v3(x0), v3(x1), v3(x0 + 3), v3(x1 + 3), offset, ghat, prm);
    if (overlap) *overlap = accd::g_ccd_overlap;
    return r;
}
// x0: ea0, ea1, eb0, eb1
float ref_edge_edge_ccd(const float *x0, const float *x1, float offset, float ghat, float max_t, float ccd_eps, unsigned *overlap) {
    accd::g_ccd_overlap = 0;
    ParamSet prm = mkparam(max_t, ccd_eps);
    float r = accd::edge_edge_ccd(v3(x0), v3(x0 + 3), v3(x0 + 6), v3(x0 + 9), v3(x1), v3(x1 + 3), v3(x1 + 6), v3(x1 + 9), offset, ghat, prm);
    if (overlap) *overlap = accd::g_ccd_overlap;
    return r;
}
int ref_edge_triangle_intersect(const float *a0, const float *a1, const float *b0, const float *b1, const float *b2) {
    // contact.cu:2745-2759 wrapper over intersect_core.hpp
    Vec3f d1 = v3(b1) - v3(b0), d2 = v3(b2) - v3(b0), e0 = v3(a0) - v3(b0), e1 = v3(a1) - v3(b0);
    const float zero[3] = {0, 0, 0};
    const float e0a[3] = {e0[0], e0[1], e0[2]}, e1a[3] = {e1[0], e1[1], e1[2]};
    const float d1a[3] = {d1[0], d1[1], d1[2]}, d2a[3] = {d2[0], d2[1], d2[2]};
    return ppf_isect::edge_triangle_intersect<float>(e0a, e1a, zero, d1a, d2a) ? 1 : 0;
}
void ref_cubic(float g, float ghat, float offset, float *out3) {
    out3[0] = cubic::energy(g, ghat, offset);
    out3[1] = cubic::gradient(g, ghat, offset);
    out3[2] = cubic::curvature(g, ghat, offset);
}
// friction: force_contact[3], dx[3], normal[3] -> grad[3], hess[9] (column-major)
void ref_friction(const float *fc, const float *dx, const float *n, float mu, float min_dx, float *grad, float *hess, float *lambda) {
    Friction f(v3(fc), v3(dx), v3(n), mu, min_dx);
    Vec3f g = f.gradient();
    Mat3x3f H = f.hessian();
    for (int k = 0; k < 3; ++k) grad[k] = g[k];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) hess[c * 3 + r] = H(r, c);
    *lambda = f.lambda;
}
// Baraff-Witkin in F space: F column-major 3x2 -> dedF (6, column-major 3x2), d2edF2 (36, column-major 6x6)
void ref_baraffwitkin(const float *F6, float mu, float lmd, float *dedF, float *d2edF2) {
    Mat3x2f F;
    for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) F(r, c) = F6[c * 3 + r];
    Mat3x2f g = BaraffWitkin::stretch_gradient(F, mu) + BaraffWitkin::shear_gradient(F, lmd);
    Mat6x6f H = BaraffWitkin::stretch_hessian(F, mu) + BaraffWitkin::shear_hessian(F, lmd);
    for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) dedF[c * 3 + r] = g(r, c);
    for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) d2edF2[c * 6 + r] = H(r, c);
}
// hinge: x[12] = positions of hinge vertices h0..h3 (ORIGINAL hinge order) -> force (3x4 col-major, in the
// REMAPPED order the reference scatters with), hess (12x12 col-major, remapped order), remapped index perm
void ref_hinge(const float *x12, float rest_angle, float *force12, float *hess144, float *angle) {
    Vec3f pts[4] = {v3(x12), v3(x12 + 3), v3(x12 + 6), v3(x12 + 9)};
    Vec<Vec3f> vertex;
    vertex.data = pts;
    vertex.size = 4;
    vertex.allocated = 4;
    Vec4u hinge(0u, 1u, 2u, 3u);
    Mat3x4f f;
    Mat12x12f H;
    dihedral_angle::face_compute_force_hessian(vertex, hinge, rest_angle, f, H);
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) force12[c * 3 + r] = f(r, c);
    for (int c = 0; c < 12; ++c) for (int r = 0; r < 12; ++r) hess144[c * 12 + r] = H(r, c);
    Vec4u hm = dihedral_angle::remap(Vec4u(0u, 1u, 2u, 3u));
    *angle = dihedral_angle::face_dihedral_angle(pts[hm[0]], pts[hm[1]], pts[hm[2]], pts[hm[3]]);
}
// symmetric 3x3 eigen (column-major in) -> ascending eigenvalues, eigenvectors (columns)
void ref_symm3x3(const float *A9, float *val3, float *vec9) {
    Mat3x3f A, Q;
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) A(r, c) = A9[c * 3 + r];
    Vec3f l;
    linalg::eig::symm3x3(A, l, Q);
    for (int k = 0; k < 3; ++k) val3[k] = l[k];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) vec9[c * 3 + r] = Q(r, c);
}
// rod stretch (hook.hpp): x6 = the two end points -> gradient (3x2 col-major), hess (6x6 col-major)
void ref_hook(const float *x6, float l0, float weight, float *grad6, float *hess36) {
    Mat3x2f g;
    Mat6x6f H;
    hook::make_diff_table(v3(x6), v3(x6 + 3), l0, weight, g, H);
    for (int c = 0; c < 2; ++c) for (int r = 0; r < 3; ++r) grad6[c * 3 + r] = g(r, c);
    for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) hess36[c * 6 + r] = H(r, c);
}
// rod bending (dihedral_angle.hpp strand_compute_force_hessian): x9 = (x0, x1, x2) -> force (3x3 col-major), hess (9x9 col-major)
void ref_strand(const float *x9, float rest_angle, float *force9, float *hess81) {
    Mat3x3f f;
    Mat9x9f H;
    dihedral_angle::strand_compute_force_hessian(v3(x9), v3(x9 + 3), v3(x9 + 6), rest_angle, f, H);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) force9[c * 3 + r] = f(r, c);
    for (int c = 0; c < 9; ++c) for (int r = 0; r < 9; ++r) hess81[c * 9 + r] = H(r, c);
}
// isotropic shell models: table of derivatives w.r.t. the two singular values; model 0 ARAP, 1 StVK, 3 SNHk
// out5 = deda[0], deda[1], d2(0,0), d2(0,1), d2(1,1)
void ref_iso_table2(int model, const float *a2, float mu, float lmd, float *out5) {
    Vec2f a(a2[0], a2[1]);
    DiffTable2 t = model == 0 ? ARAP::make_diff_table2(a, mu, lmd) : (model == 1 ? StVK::make_diff_table2(a, mu, lmd) : SNHk::make_diff_table2(a, mu, lmd));
    out5[0] = t.deda[0]; out5[1] = t.deda[1]; out5[2] = t.d2ed2a(0, 0); out5[3] = t.d2ed2a(0, 1); out5[4] = t.d2ed2a(1, 1);
}
// the three barriers of ParamSet.barrier: kind 0 cubic, 1 quad, 2 log -> gradient, curvature
void ref_barrier(int kind, float g, float ghat, float offset, float *out2) {
    if (kind == 0) { out2[0] = cubic::gradient(g, ghat, offset); out2[1] = cubic::curvature(g, ghat, offset); }
    else if (kind == 1) { out2[0] = quadratic::gradient(g, ghat, offset); out2[1] = quadratic::curvature(g, ghat, offset); }
    else { out2[0] = logarithm::gradient(g, ghat, offset); out2[1] = logarithm::curvature(g, ghat, offset); }
}
}

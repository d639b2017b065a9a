// math_accd.cuh -- additive conservative CCD (ACCD) time-of-impact per primitive pair (fp32).
//
// The per-pair result is deterministic fp32, and the line search takes the minimum over all
// pairs, so the step fraction of a Newton iteration matches the reference bit for bit only
// if this routine repeats the reference's arithmetic.  It therefore follows, operation by
// operation, crates/ppf-cts-solver/src/cpp/contact/accd.hpp:
//   centerize 65-74, rescale 77-120, max_relative_u 142-151, directional_advance 242-282,
//   park_floor 289, ccd_helper 292-471, distance functors 473-561, *_ccd drivers 565-696.
// Columns are stored as x[c][r] (the reference's column-major 3xC matrices).
#pragma once
#include "math_distance.cuh"

namespace ppf {

struct CcdOverlap { // what the reference latches in g_ccd_overlap* (accd.hpp:23-45)
    unsigned flag;
    float d2, offset;
};

enum CcdKind { CCD_PT = 0, CCD_EE = 1, CCD_PE = 2, CCD_PP = 3 };

template <int C> struct Cols {
    float v[C][3];
};

template <int C> PPF_HD void accd_centerize(Cols<C> &x) {
    float mov[3] = {0.0f, 0.0f, 0.0f};
    const float scale = float(1.0 / C);
    for (int k = 0; k < C; ++k)
        for (int r = 0; r < 3; ++r) mov[r] = mov[r] + scale * x.v[k][r];
    for (int k = 0; k < C; ++k)
        for (int r = 0; r < 3; ++r) x.v[k][r] = x.v[k][r] - mov[r];
}

template <int C> PPF_HD float accd_rescale(Cols<C> &x, Cols<C> &dx, float max_t, CcdOverlap &ov) {
    float m0 = fabsf(x.v[0][0]);
    float m1 = fabsf(x.v[0][0] + max_t * dx.v[0][0]);
    for (int k = 0; k < C; ++k)
        for (int r = 0; r < 3; ++r) {
            if (k == 0 && r == 0) continue;
            const float a = fabsf(x.v[k][r]);
            if (m0 < a) m0 = a;
            const float b = fabsf(x.v[k][r] + max_t * dx.v[k][r]);
            if (m1 < b) m1 = b;
        }
    const float max_entry = m0 > m1 ? m0 : (m1 > m0 ? m1 : m0); // std::max(a,b): b if a<b else a
    float scale = 0.99f / max_entry;
    if (!(scale > 0.0f) || !isfinite(scale)) {
        ov.flag = 1u; ov.d2 = 0.0f; ov.offset = -1.0f; // report_coincident()
        return 0.0f;
    }
    float scaled = 1.0f;
    const float s = 8.0f;
    while (true) {
        if (scale > s) {
            for (int k = 0; k < C; ++k)
                for (int r = 0; r < 3; ++r) { x.v[k][r] = x.v[k][r] * s; dx.v[k][r] = dx.v[k][r] * s; }
            scaled *= s;
            scale /= s;
        } else {
            for (int k = 0; k < C; ++k)
                for (int r = 0; r < 3; ++r) { x.v[k][r] = x.v[k][r] * scale; dx.v[k][r] = dx.v[k][r] * scale; }
            scaled *= scale;
            break;
        }
    }
    return scaled;
}

template <int C, int SPLIT> PPF_HD float accd_max_relative_u(const Cols<C> &u) {
    float max_u = 0.0f;
    for (int i = 0; i < SPLIT; ++i)
        for (int j = SPLIT; j < C; ++j) {
            const float d0 = u.v[i][0] - u.v[j][0], d1 = u.v[i][1] - u.v[j][1], d2 = u.v[i][2] - u.v[j][2];
            float sq = 0.0f; // SMat::squaredNorm accumulation order (see vecmath.cuh dot)
            sq = sq + d0 * d0; sq = sq + d1 * d1; sq = sq + d2 * d2;
            max_u = max_u < sq ? sq : max_u; // std::max<float>(max_u, sq)
        }
    return sqrtf(max_u);
}

PPF_HD V3 colv(const float *c) { return mk(c[0], c[1], c[2]); }

// squared distance between the two primitives stored in the columns of x; `dir` receives the
// (unnormalised) closest-point difference (accd.hpp:473-561)
template <int KIND, int C> PPF_HD float accd_sqdist(const Cols<C> &x, V3 &dir) {
    if (KIND == CCD_PT) {
        const V3 p = colv(x.v[0]), t0 = colv(x.v[1]), t1 = colv(x.v[2]), t2 = colv(x.v[3]);
        const C3 c = point_triangle_coeff_unclassified(p, t0, t1, t2);
        dir = c.a * (t0 - p) + c.b * (t1 - p) + c.c * (t2 - p);
        return dot(dir, dir);
    } else if (KIND == CCD_EE) {
        const V3 p0 = colv(x.v[0]), p1 = colv(x.v[1]), q0 = colv(x.v[2]), q1 = colv(x.v[3]);
        const C4 c = edge_edge_coeff(p0, p1, q0, q1);
        const V3 a = c.a * p0 + c.b * p1;
        const V3 b = c.c * q0 + c.d * q1;
        dir = b - a;
        return dot(dir, dir);
    } else if (KIND == CCD_PE) {
        const V3 p = colv(x.v[0]), q0 = colv(x.v[1]), q1 = colv(x.v[2]);
        const C2 c = point_edge_coeff_unclassified(p, q0, q1);
        const V3 q = c.a * q0 + c.b * q1;
        dir = q - p;
        return dot(dir, dir);
    } else {
        const V3 p = colv(x.v[0]), q = colv(x.v[1]);
        dir = q - p;
        return dot(dir, dir);
    }
}

template <int C, int SPLIT>
PPF_HD float accd_directional_advance(const Cols<C> &x, const Cols<C> &dx, const V3 &w, float w_norm, float dip,
                                      float park, float max_t) {
    const float round_slack = 2.0f * FLT_EPSILON;
    const float coord_slack = 6.2e-7f;
    const V3 w_abs = mk(fabsf(w.x), fabsf(w.y), fabsf(w.z));
    const float dip_w = dip * w_norm;
    const float park_w = park * w_norm;
    float step = FLT_MAX;
    for (int i = 0; i < SPLIT; ++i)
        for (int j = SPLIT; j < C; ++j) {
            const V3 gap = mk(x.v[j][0] - x.v[i][0], x.v[j][1] - x.v[i][1], x.v[j][2] - x.v[i][2]);
            const V3 rel = mk(dx.v[i][0] - dx.v[j][0], dx.v[i][1] - dx.v[j][1], dx.v[i][2] - dx.v[j][2]);
            const V3 gap_abs = mk(fabsf(gap.x), fabsf(gap.y), fabsf(gap.z));
            const V3 rel_abs = mk(fabsf(rel.x), fabsf(rel.y), fabsf(rel.z));
            const float err = round_slack * (dot(w_abs, gap_abs) + max_t * dot(w_abs, rel_abs) + park_w) + coord_slack * w_norm;
            const float projected = dot(w, gap) - err;
            if (!(projected - dip_w > 0.0f)) return 0.0f;
            const float rate = dot(w, rel);
            if (rate > 0.0f) step = fminf(step, fmaxf(0.0f, projected - park_w) / rate);
        }
    return step;
}

PPF_HD float accd_park_floor(float ghat) { return 1e-2f * ghat; }

template <int KIND, int C, int SPLIT>
PPF_HD float accd_helper(const Cols<C> &x0, const Cols<C> &dx, float u_max, float offset, float eps, float floor_gap,
                         float max_t, CcdOverlap &ov) {
    Cols<C> x = x0;
    V3 w;
    float lower_t = 0.0f;
    float toi = 0.0f;
    float d2 = accd_sqdist<KIND, C>(x0, w);
    const float target_squared = offset * offset;
    if (!(d2 > target_squared)) {
        ov.flag = 1u; ov.d2 = d2; ov.offset = offset;
        return 0.0f;
    }
    const float clearance = sqrtf(d2) - offset;
    const float floor_clear = fmaxf(2.0f * eps, floor_gap);
    const float park = offset + fminf(floor_clear, fmaxf(clearance - 0.5f * eps, 0.5f * clearance));
    const float overshoot = fmaxf(0.0f, (park - offset) - 0.5f * floor_clear);
    const float dip = park - overshoot;
    const float park_squared = fminf(park * park, nextafterf(d2, 0.0f));
    const float inv_u_max = 1.0f / u_max;
    const unsigned max_probes = 4096u;
    unsigned probes = 0;
    while (true) {
        if (!(d2 > park_squared)) break;
        lower_t = toi;
        const float dist = sqrtf(d2);
        float step = (dist - dip) * inv_u_max;
        if (toi + step <= max_t) {
            step = fmaxf(step, accd_directional_advance<C, SPLIT>(x, dx, w, dist * (1.0f + 2.0f * FLT_EPSILON), dip, park, max_t));
        }
        toi += step;
        if (toi > max_t) return max_t;
        else if (++probes > max_probes) return lower_t;
        for (int k = 0; k < C; ++k)
            for (int r = 0; r < 3; ++r) x.v[k][r] = x0.v[k][r] + toi * dx.v[k][r];
        d2 = accd_sqdist<KIND, C>(x, w);
    }
    float upper_t = toi;
    float window = upper_t - lower_t;
    while (true) {
        toi = 0.5f * (upper_t + lower_t);
        for (int k = 0; k < C; ++k)
            for (int r = 0; r < 3; ++r) x.v[k][r] = x0.v[k][r] + toi * dx.v[k][r];
        V3 tmp;
        d2 = accd_sqdist<KIND, C>(x, tmp);
        if (d2 > park_squared) lower_t = toi;
        else upper_t = toi;
        const float new_window = upper_t - lower_t;
        if (new_window == window) break;
        window = new_window;
    }
    if (!(lower_t > 0.0f)) {
        ov.flag = 1u; ov.d2 = d2; ov.offset = offset;
        return 0.0f;
    }
    return lower_t;
}

template <int KIND, int C, int SPLIT>
PPF_HD float accd_run(Cols<C> &x0, Cols<C> &dx, float offset, float ghat, float max_t, float ccd_eps, CcdOverlap &ov) {
    accd_centerize<C>(x0);
    accd_centerize<C>(dx);
    const float scale = accd_rescale<C>(x0, dx, max_t, ov);
    if (!(scale > 0.0f)) return 0.0f;
    const float u_max = accd_max_relative_u<C, SPLIT>(dx);
    if (u_max) {
        return accd_helper<KIND, C, SPLIT>(x0, dx, u_max, scale * offset, scale * ccd_eps, scale * accd_park_floor(ghat), max_t, ov);
    }
    return max_t;
}

PPF_HD void set_col(float *c, const V3 &v) { c[0] = v.x; c[1] = v.y; c[2] = v.z; }

// accd.hpp:565-596
PPF_HD float point_triangle_ccd(const V3 &p0, const V3 &p1, const V3 &t00, const V3 &t01, const V3 &t02, const V3 &t10,
                                const V3 &t11, const V3 &t12, float offset, float ghat, float max_t, float ccd_eps,
                                CcdOverlap &ov) {
    Cols<4> x0, dx;
    set_col(x0.v[0], p0); set_col(x0.v[1], t00); set_col(x0.v[2], t01); set_col(x0.v[3], t02);
    set_col(dx.v[0], p1 - p0); set_col(dx.v[1], t10 - t00); set_col(dx.v[2], t11 - t01); set_col(dx.v[3], t12 - t02);
    return accd_run<CCD_PT, 4, 1>(x0, dx, offset, ghat, max_t, ccd_eps, ov);
}
// accd.hpp:598-627
PPF_HD float point_edge_ccd(const V3 &p0, const V3 &p1, const V3 &e00, const V3 &e01, const V3 &e10, const V3 &e11,
                            float offset, float ghat, float max_t, float ccd_eps, CcdOverlap &ov) {
    Cols<3> x0, dx;
    set_col(x0.v[0], p0); set_col(x0.v[1], e00); set_col(x0.v[2], e01);
    set_col(dx.v[0], p1 - p0); set_col(dx.v[1], e10 - e00); set_col(dx.v[2], e11 - e01);
    return accd_run<CCD_PE, 3, 1>(x0, dx, offset, ghat, max_t, ccd_eps, ov);
}
// accd.hpp:629-655
PPF_HD float point_point_ccd(const V3 &p0, const V3 &p1, const V3 &q0, const V3 &q1, float offset, float ghat,
                             float max_t, float ccd_eps, CcdOverlap &ov) {
    Cols<2> x0, dx;
    set_col(x0.v[0], p0); set_col(x0.v[1], q0);
    set_col(dx.v[0], p1 - p0); set_col(dx.v[1], q1 - q0);
    return accd_run<CCD_PP, 2, 1>(x0, dx, offset, ghat, max_t, ccd_eps, ov);
}
// accd.hpp:657-692
PPF_HD float edge_edge_ccd(const V3 &ea00, const V3 &ea01, const V3 &eb00, const V3 &eb01, const V3 &ea10,
                           const V3 &ea11, const V3 &eb10, const V3 &eb11, float offset, float ghat, float max_t,
                           float ccd_eps, CcdOverlap &ov) {
    Cols<4> x0, dx;
    set_col(x0.v[0], ea00); set_col(x0.v[1], ea01); set_col(x0.v[2], eb00); set_col(x0.v[3], eb01);
    set_col(dx.v[0], ea10 - ea00); set_col(dx.v[1], ea11 - ea01); set_col(dx.v[2], eb10 - eb00); set_col(dx.v[3], eb11 - eb01);
    return accd_run<CCD_EE, 4, 2>(x0, dx, offset, ghat, max_t, ccd_eps, ov);
}

} // namespace ppf

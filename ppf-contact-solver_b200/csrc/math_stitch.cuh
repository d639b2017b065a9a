// math_stitch.cuh -- one stitch (seam spring between two barycentric points of the mesh), fp32.
//
// Restates energy::embed_stitch_force_hessian (crates/ppf-cts-solver/src/cpp/energy/energy.cu:796-886): six vertex slots,
// the first three (weights summing to one) locate the source point, the last three the target; the spring acts on
// t = z0 - z1 with rest length l0 = (sum of w_i ghat_i) / 2 and its length capped at stitch_length_factor l0 + the two
// sides' largest contact offsets; gradient w_a s_a (l / l0 - 1) n and Hessian c0 g g^T + c1 dtdx^T dtdx like the Hookean rod
// (hook.hpp), both scaled by the stitch's own stiffness.  s_a = +1 for the source slots, -1 for the target slots.
#pragma once
#include "vecmath.cuh"

namespace ppf {

struct StitchOut {
    V3 f[6];
    float H[18 * 18]; // column-major
};

// x[6]: positions of the six slots, ghat[6] / offset[6]: the vertices' contact gap and offset, w[6]: the weights
PPF_HD void stitch_force_hessian(const V3 *x, const float *ghat, const float *offset, const float *w, float stiffness, float length_factor, StitchOut &o) {
    const float l0 = (w[0] * ghat[0] + w[1] * ghat[1] + w[2] * ghat[2] + w[3] * ghat[3] + w[4] * ghat[4] + w[5] * ghat[5]) / 2.0f;
    const float source_offset = fmaxf(fmaxf(offset[0], offset[1]), offset[2]);
    const float target_offset = fmaxf(fmaxf(offset[3], offset[4]), offset[5]);
    const float l_cap = length_factor * l0 + source_offset + target_offset;
    const float s = 1.0f / 6.0f;
    const V3 cog = s * x[0] + s * x[1] + s * x[2] + s * x[3] + s * x[4] + s * x[5];
    const V3 z0 = w[0] * (x[0] - cog) + w[1] * (x[1] - cog) + w[2] * (x[2] - cog);
    const V3 z1 = w[3] * (x[3] - cog) + w[4] * (x[4] - cog) + w[5] * (x[5] - cog);
    const V3 t = z0 - z1;
    const float l = fminf(l_cap, norm(t));
    const V3 n = t / l;
    const float sw[6] = {w[0], w[1], w[2], -w[3], -w[4], -w[5]}; // dtdx = [w0 I, w1 I, w2 I, -w3 I, -w4 I, -w5 I]
    const V3 dedt = (l / l0 - 1.0f) * n;
    float g[18];
    for (int a = 0; a < 6; ++a) { g[3 * a] = sw[a] * n.x; g[3 * a + 1] = sw[a] * n.y; g[3 * a + 2] = sw[a] * n.z; }
    const float r = (l - l0) / l;
    const float c0 = fmaxf(0.0f, 1.0f - r) / l0;
    const float c1 = fmaxf(0.0f, r / l0);
    for (int a = 0; a < 6; ++a) o.f[a] = stiffness * (sw[a] * dedt);
    for (int j = 0; j < 18; ++j)
        for (int i = 0; i < 18; ++i) {
            const float dd = (i % 3 == j % 3) ? (c1 * sw[i / 3]) * sw[j / 3] : 0.0f; // (c1 dtdx^T) dtdx
            o.H[j * 18 + i] = stiffness * ((c0 * g[i]) * g[j] + dd);
        }
}

} // namespace ppf

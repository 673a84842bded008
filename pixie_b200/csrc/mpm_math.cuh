// 3x3 fp32 helpers for the MPM kernels: row-major mat33 (warp_utils.py mat33 layout), a one-sided
// Jacobi SVD with the sign convention of wp.svd3 (U, V rotations; the last singular value carries
// the sign of det F), and the constitutive functions of mpm_utils.py:10-279.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pixie {
namespace mpm {

struct M3 { float m[9]; };   // row-major: m[3*r + c]
struct V3 { float x, y, z; };

__device__ __forceinline__ M3 m3_zero() { M3 a; for (int i = 0; i < 9; ++i) a.m[i] = 0.f; return a; }
__device__ __forceinline__ M3 m3_ident() { M3 a = m3_zero(); a.m[0] = a.m[4] = a.m[8] = 1.f; return a; }
__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 c;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            c.m[3 * r + k] = a.m[3 * r] * b.m[k] + a.m[3 * r + 1] * b.m[3 + k] + a.m[3 * r + 2] * b.m[6 + k];
    return c;
}
__device__ __forceinline__ M3 m3_mul_t(const M3& a, const M3& b) {   // a * b^T
    M3 c;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            c.m[3 * r + k] = a.m[3 * r] * b.m[3 * k] + a.m[3 * r + 1] * b.m[3 * k + 1] + a.m[3 * r + 2] * b.m[3 * k + 2];
    return c;
}
__device__ __forceinline__ M3 m3_t(const M3& a) {
    M3 c;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) c.m[3 * r + k] = a.m[3 * k + r];
    return c;
}
__device__ __forceinline__ float m3_det(const M3& a) {
    return a.m[0] * (a.m[4] * a.m[8] - a.m[5] * a.m[7]) - a.m[1] * (a.m[3] * a.m[8] - a.m[5] * a.m[6]) +
           a.m[2] * (a.m[3] * a.m[7] - a.m[4] * a.m[6]);
}
__device__ __forceinline__ V3 m3_mulv(const M3& a, const V3& v) {
    return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
// U * diag(s) * V^T
__device__ __forceinline__ M3 m3_usvt(const M3& U, const V3& s, const M3& V) {
    M3 us = U;
#pragma unroll
    for (int r = 0; r < 3; ++r) { us.m[3 * r] *= s.x; us.m[3 * r + 1] *= s.y; us.m[3 * r + 2] *= s.z; }
    return m3_mul_t(us, V);
}

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix: F = U diag(sig) V^T.
// U and V are proper rotations, |sig| sorted descending, sig.z has the sign of det(F)
// (the convention of wp.svd3 / McAdams et al. that mpm_utils.py relies on).
__device__ __forceinline__ void svd3(const M3& F, M3& U, V3& sig, M3& V) {
    // columns of B (= F V) and of V
    float b[3][3], v[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) { b[c][r] = F.m[3 * r + c]; v[c][r] = (r == c) ? 1.f : 0.f; }
    // Rotation of a column pair: with d = be - al, g = 2 ga, h = sqrt(d^2 + g^2) the textbook
    //   zeta = d / g, t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), c = 1 / sqrt(1 + t^2), s = c t
    // is c^2 = (1 + |d| / h) / 2, s = sign(d) g / (2 h c): two rsqrt (one Newton step each: c^2 + s^2 = 1 to 3e-7) instead of
    // three IEEE divisions and two square roots (~60 instructions per rotation, r02 ncu). A sweep that rotates nothing ends
    // the iteration (a pair counts as orthogonal below |cos| = 3e-7 ~ 5 ulp; the fixed five sweeps kept rotating rounding noise).
#pragma unroll 1
    for (int sweep = 0; sweep < 5; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pair = 0; pair < 3; ++pair) {
            const int p = (pair == 2) ? 1 : 0;
            const int q = (pair == 0) ? 1 : 2;
            const float al = b[p][0] * b[p][0] + b[p][1] * b[p][1] + b[p][2] * b[p][2];
            const float be = b[q][0] * b[q][0] + b[q][1] * b[q][1] + b[q][2] * b[q][2];
            const float ga = b[p][0] * b[q][0] + b[p][1] * b[q][1] + b[p][2] * b[q][2];
            if (fabsf(ga) > 1e-20f && ga * ga > 1e-13f * al * be) {
                rotated = true;
                const float d = be - al, g = 2.f * ga;
                const float hh = d * d + g * g;
                float rh = rsqrtf(hh);
                rh = rh * (1.5f - 0.5f * hh * rh * rh);
                const float c2 = 0.5f + 0.5f * fabsf(d) * rh;
                float rc = rsqrtf(c2);
                rc = rc * (1.5f - 0.5f * c2 * rc * rc);
                const float cs = c2 * rc;
                const float sn = copysignf(0.5f, d) * g * rh * rc;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float bp = b[p][r], bq = b[q][r];
                    b[p][r] = cs * bp - sn * bq;
                    b[q][r] = sn * bp + cs * bq;
                    const float vp = v[p][r], vq = v[q][r];
                    v[p][r] = cs * vp - sn * vq;
                    v[q][r] = sn * vp + cs * vq;
                }
            }
        }
        if (!rotated) break;
    }
    float n[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) n[c] = b[c][0] * b[c][0] + b[c][1] * b[c][1] + b[c][2] * b[c][2];
    // sort columns by descending norm; a swap with a negation keeps det(V) = +1
#define PIXIE_SWAPCOL(i, j)                                                   \
    if (n[i] < n[j]) {                                                        \
        float tn = n[i]; n[i] = n[j]; n[j] = tn;                              \
        for (int r = 0; r < 3; ++r) {                                         \
            float tb = b[i][r]; b[i][r] = b[j][r]; b[j][r] = -tb;             \
            float tv = v[i][r]; v[i][r] = v[j][r]; v[j][r] = -tv;             \
        }                                                                     \
    }
    PIXIE_SWAPCOL(0, 1)
    PIXIE_SWAPCOL(0, 2)
    PIXIE_SWAPCOL(1, 2)
#undef PIXIE_SWAPCOL
    const float s0 = sqrtf(n[0]), s1 = sqrtf(n[1]);
    float u0[3], u1[3], u2[3];
    if (s0 > 1e-30f) { for (int r = 0; r < 3; ++r) u0[r] = b[0][r] / s0; }
    else { u0[0] = 1.f; u0[1] = 0.f; u0[2] = 0.f; }
    // second column: remove any residual component along u0 before normalising
    {
        float d = u0[0] * b[1][0] + u0[1] * b[1][1] + u0[2] * b[1][2];
        float w0 = b[1][0] - d * u0[0], w1 = b[1][1] - d * u0[1], w2 = b[1][2] - d * u0[2];
        float nn = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
        if (nn > 1e-30f && s1 > 1e-30f) { u1[0] = w0 / nn; u1[1] = w1 / nn; u1[2] = w2 / nn; }
        else {
            // any unit vector orthogonal to u0
            if (fabsf(u0[0]) < 0.9f) { w0 = 0.f; w1 = -u0[2]; w2 = u0[1]; }
            else { w0 = -u0[2]; w1 = 0.f; w2 = u0[0]; }
            nn = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
            u1[0] = w0 / nn; u1[1] = w1 / nn; u1[2] = w2 / nn;
        }
    }
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
    u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
    u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    sig.x = s0;
    sig.y = s1;
    sig.z = u2[0] * b[2][0] + u2[1] * b[2][1] + u2[2] * b[2][2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        U.m[3 * r] = u0[r]; U.m[3 * r + 1] = u1[r]; U.m[3 * r + 2] = u2[r];
        V.m[3 * r] = v[0][r]; V.m[3 * r + 1] = v[1][r]; V.m[3 * r + 2] = v[2][r];
    }
}

// Rotation factor R = U V^T of the polar decomposition F = R S, by scaled Newton iteration
// R <- (g R + (g R)^-T) / 2 (Higham). kirchoff_stress_FCR (mpm_utils.py:10-17) uses the SVD only through
// R = U V^T, so for det F > 0 this is the same matrix at ~1/8 of the instructions of svd3.
// Returns false (caller falls back to svd3) if det F <= 0 or the iteration does not settle.
__device__ __forceinline__ bool polar_rotation(const M3& F, M3& R) {
    R = F;
    bool last = false;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        // cofactor matrix: inv(R)^T = cof(R) / det(R); det(R) is the first row of R times the first row of cof(R)
        M3 cof;
        cof.m[0] = R.m[4] * R.m[8] - R.m[5] * R.m[7];
        cof.m[1] = R.m[5] * R.m[6] - R.m[3] * R.m[8];
        cof.m[2] = R.m[3] * R.m[7] - R.m[4] * R.m[6];
        cof.m[3] = R.m[2] * R.m[7] - R.m[1] * R.m[8];
        cof.m[4] = R.m[0] * R.m[8] - R.m[2] * R.m[6];
        cof.m[5] = R.m[1] * R.m[6] - R.m[0] * R.m[7];
        cof.m[6] = R.m[1] * R.m[5] - R.m[2] * R.m[4];
        cof.m[7] = R.m[2] * R.m[3] - R.m[0] * R.m[5];
        cof.m[8] = R.m[0] * R.m[4] - R.m[1] * R.m[3];
        const float det = R.m[0] * cof.m[0] + R.m[1] * cof.m[1] + R.m[2] * cof.m[2];
        if (!(det > 1e-12f)) return false;
        // scaling g = det^(-1/3) speeds up the first steps; g -> 1 as R approaches a rotation
        const float g = (it < 2 && !last) ? rcbrtf(det) : 1.0f;
        const float a = 0.5f * g, b = __fdividef(0.5f, g * det);
        float delta = 0.f;
        M3 N;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            N.m[i] = a * R.m[i] + b * cof.m[i];
            delta = fmaxf(delta, fabsf(N.m[i] - R.m[i]));
        }
        R = N;
        if (last) return true;
        // quadratic convergence: once a step moves entries by < 3e-4 the next (unscaled) step is exact to fp32
        if (delta < 3e-4f) last = true;
    }
    return false;
}

// ---- Kirchhoff stresses (mpm_utils.py:10-86); tau = P F^T
__device__ __forceinline__ M3 stress_fcr_R(const M3& F, const M3& R, float J, float mu, float lam) {
    M3 d;
    for (int i = 0; i < 9; ++i) d.m[i] = F.m[i] - R.m[i];
    M3 s = m3_mul_t(d, F);
    const float a = 2.f * mu, p = lam * J * (J - 1.f);
    for (int i = 0; i < 9; ++i) s.m[i] *= a;
    s.m[0] += p; s.m[4] += p; s.m[8] += p;
    return s;
}
__device__ __forceinline__ M3 stress_fcr(const M3& F, const M3& U, const M3& V, float J, float mu, float lam) {
    const M3 R = m3_mul_t(U, V);
    M3 d;
    for (int i = 0; i < 9; ++i) d.m[i] = F.m[i] - R.m[i];
    M3 s = m3_mul_t(d, F);
    const float a = 2.f * mu, p = lam * J * (J - 1.f);
    for (int i = 0; i < 9; ++i) s.m[i] *= a;
    s.m[0] += p; s.m[4] += p; s.m[8] += p;
    return s;
}
__device__ __forceinline__ M3 stress_stvk(const M3& F, const M3& U, const M3& V, V3 sig, float mu, float lam) {
    sig.x = fmaxf(sig.x, 0.01f); sig.y = fmaxf(sig.y, 0.01f); sig.z = fmaxf(sig.z, 0.01f);
    const float e0 = logf(sig.x), e1 = logf(sig.y), e2 = logf(sig.z);
    const float tr = e0 + e1 + e2;
    const V3 tau = {2.f * mu * e0 + lam * tr, 2.f * mu * e1 + lam * tr, 2.f * mu * e2 + lam * tr};
    return m3_mul_t(m3_usvt(U, tau, V), F);
}
__device__ __forceinline__ M3 stress_drucker_prager(const M3& F, const M3& U, const M3& V, V3 sig, float mu, float lam) {
    const float l0 = logf(sig.x), l1 = logf(sig.y), l2 = logf(sig.z);
    const float tr = l0 + l1 + l2;
    const V3 c = {2.f * mu * l0 * (1.f / sig.x) + lam * tr * (1.f / sig.x),
                  2.f * mu * l1 * (1.f / sig.y) + lam * tr * (1.f / sig.y),
                  2.f * mu * l2 * (1.f / sig.z) + lam * tr * (1.f / sig.z)};
    return m3_mul_t(m3_usvt(U, c, V), F);
}
__device__ __forceinline__ M3 stress_water(float J, float bulk) {
    const float pressure = -bulk * (powf(J, -1.1f) - 1.f);
    M3 s = m3_zero();
    s.m[0] = s.m[4] = s.m[8] = J * pressure;
    return s;
}

// ---- return mappings (mpm_utils.py:89-279). Each returns the elastic deformation gradient and may
// mutate the particle's yield stress / Lame parameters exactly where the reference does.
__device__ __forceinline__ M3 return_von_mises(const M3& Ft, float mu, float lam, float& yield, float hardening, float xi,
                                               bool with_damage, float softening, float& mu_io, float& lam_io) {
    M3 U, V; V3 so;
    svd3(Ft, U, so, V);
    const V3 sg = {fmaxf(so.x, 0.01f), fmaxf(so.y, 0.01f), fmaxf(so.z, 0.01f)};
    float e0 = logf(sg.x), e1 = logf(sg.y), e2 = logf(sg.z);
    const float tr = e0 + e1 + e2, temp = tr / 3.f;
    const float t0 = 2.f * mu * e0 + lam * tr, t1 = 2.f * mu * e1 + lam * tr, t2 = 2.f * mu * e2 + lam * tr;
    const float st = t0 + t1 + t2;
    const float c0 = t0 - st / 3.f, c1 = t1 - st / 3.f, c2 = t2 - st / 3.f;
    if (sqrtf(c0 * c0 + c1 * c1 + c2 * c2) > yield) {
        if (with_damage && yield <= 0.f) return Ft;
        const float h0 = e0 - temp, h1 = e1 - temp, h2 = e2 - temp;
        const float hn = sqrtf(h0 * h0 + h1 * h1 + h2 * h2) + 1e-6f;
        const float dg = hn - yield / (2.f * mu);
        const float k = dg / hn;
        e0 -= k * h0; e1 -= k * h1; e2 -= k * h2;
        if (with_damage) {
            yield = yield - softening * sqrtf((k * h0) * (k * h0) + (k * h1) * (k * h1) + (k * h2) * (k * h2));
            if (yield <= 0.f) { mu_io = 0.f; lam_io = 0.f; }
        }
        const M3 Fe = m3_usvt(U, {expf(e0), expf(e1), expf(e2)}, V);
        if (hardening == 1.f) yield = yield + 2.f * mu_io * xi * dg;
        return Fe;
    }
    return Ft;
}
__device__ __forceinline__ M3 return_viscoplastic(const M3& Ft, float mu, float yield, float plastic_viscosity, float dt) {
    M3 U, V; V3 so;
    svd3(Ft, U, so, V);
    const V3 sg = {fmaxf(so.x, 0.01f), fmaxf(so.y, 0.01f), fmaxf(so.z, 0.01f)};
    const float b0 = sg.x * sg.x, b1 = sg.y * sg.y, b2 = sg.z * sg.z;
    const float e0 = logf(sg.x), e1 = logf(sg.y), e2 = logf(sg.z);
    const float tr = e0 + e1 + e2;
    const float h0 = e0 - tr / 3.f, h1 = e1 - tr / 3.f, h2 = e2 - tr / 3.f;
    const float s0 = 2.f * mu * h0, s1 = 2.f * mu * h1, s2 = 2.f * mu * h2;
    const float sn = sqrtf(s0 * s0 + s1 * s1 + s2 * s2);
    const float y = sn - sqrtf(2.f / 3.f) * yield;
    if (y > 0.f) {
        const float mu_hat = mu * (b0 + b1 + b2) / 3.f;
        const float snew = sn - y / (1.f + plastic_viscosity / (2.f * mu_hat * dt));
        const float r = snew / sn;
        const float k = 1.f / (2.f * mu);
        return m3_usvt(U, {expf(k * (r * s0) + tr / 3.f), expf(k * (r * s1) + tr / 3.f), expf(k * (r * s2) + tr / 3.f)}, V);
    }
    return Ft;
}
__device__ __forceinline__ M3 return_sand(const M3& Ft, float mu, float lam, float alpha) {
    M3 U, V; V3 sg;
    svd3(Ft, U, sg, V);
    const float e0 = logf(fmaxf(fabsf(sg.x), 1e-14f)), e1 = logf(fmaxf(fabsf(sg.y), 1e-14f)), e2 = logf(fmaxf(fabsf(sg.z), 1e-14f));
    const float tr = e0 + e1 + e2;
    const float h0 = e0 - tr / 3.f, h1 = e1 - tr / 3.f, h2 = e2 - tr / 3.f;
    const float hn = sqrtf(h0 * h0 + h1 * h1 + h2 * h2);
    const float dg = hn + (3.f * lam + 2.f * mu) / (2.f * mu) * tr * alpha;
    if (dg <= 0.f) return Ft;
    if (tr > 0.f) return m3_mul_t(U, V);
    const float k = dg / hn;
    return m3_usvt(U, {expf(e0 - h0 * k), expf(e1 - h1 * k), expf(e2 - h2 * k)}, V);
}

}  // namespace mpm
}  // namespace pixie

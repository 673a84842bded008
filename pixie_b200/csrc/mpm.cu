// MLS-MPM / APIC substep for sm_100a.  Replaces MPM_Simulator_WARP.p2g2p
// (third_party/PhysGaussian/mpm_solver_warp/mpm_solver_warp.py:514-637) and the Warp kernels it
// launches (mpm_utils.py:295-588, BC closures mpm_solver_warp.py:785-1179).
//
// Two paths:
//   default ("fused", mpm_fused.cuh): a private cell-sorted SoA copy of the particle state and two launches per
//     substep (particle kernel: g2p(i) + BCs/stress/p2g(i+1); grid kernel over the particles' node box), replayed from
//     a CUDA graph of 50 substeps with the simulation clock on the device;
//   direct (this file; slab-decomposed runs, PIXIE_MPM_DIRECT=1): four launches per substep on the caller's arrays
//     mpm_stress  : [impulse / Dirichlet particle BCs] -> return mapping + Kirchhoff stress
//     mpm_scatter : warp-aggregated p2g (one red.global.add.v4.f32 per run and node: grid node = float4 {mv.xyz, m})
//     mpm_grid    : normalise + gravity + damping + every grid BC (from a device BC table, registration
//                   order) -> grid_v; clears the {mv, m} node it just consumed (zero_grid fused away)
//     mpm_g2p     : gather, x/v/C/F_trial update, optional covariance update; thread 0 advances the
//                   simulation clock and moves the cuboid colliders (the reference's host-side
//                   `modify`, mpm_solver_warp.py:899-905, and `self.time += dt`, :637)
#include "mpm.cuh"
#include "mpm_math.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/cub.cuh>

namespace pixie {

using namespace mpm;

namespace {

constexpr int kMaxBC = 96;     // release_particles_sequentially registers 50 modifiers on its own

struct DevBC {
    int kind;
    float point[3], normal[3], size[3], velocity[3];
    float start_time, end_time, friction;
    int surface_type, reset;
    float h1[3], h2[3], hhr[2];
    float rotation_scale, translation_scale;
    const int* mask;
};

struct DevState {
    // particles
    float *x, *v, *F, *F_trial, *C, *stress, *R, *cov, *init_cov;
    float *vol, *mass, *density, *E, *nu, *mu, *lam, *bulk, *yield_stress;
    int *material, *selection;
    const int* order;   // thread i of p2g handles particle order[i] (cell-sorted, possibly a few substeps stale), or nullptr
    // grid
    float4* grid_mv;    // {momentum.xyz, mass}
    float4* grid_v;     // {velocity.xyz, 0}
    // clock + BCs
    double* time;
    DevBC* bcs;
    int n_bc;
    // scalars
    int n, n_grid;
    int x_begin, x_end;       // grid planes updated by mpm_grid_kernel
    float dx, inv_dx;
    float gx, gy, gz;
    float rpic_damping, grid_v_damping_scale, alpha, hardening, xi, plastic_viscosity, softening;
    int update_cov_with_F;
    int scatter_slices;       // 3: one thread per (particle, x-slice of the stencil); 1: one thread per particle
};

__device__ __forceinline__ M3 load_m3(const float* p, int i) {
    M3 a;
#pragma unroll
    for (int k = 0; k < 9; ++k) a.m[k] = p[(size_t)i * 9 + k];
    return a;
}
__device__ __forceinline__ void store_m3(float* p, int i, const M3& a) {
#pragma unroll
    for (int k = 0; k < 9; ++k) p[(size_t)i * 9 + k] = a.m[k];
}

struct Weights { int bx, by, bz; float fx[3]; float w[3][3]; float dw[3][3]; };   // [axis][node]

__device__ __forceinline__ Weights bspline_t(float inv_dx, float px, float py, float pz) {
    Weights W;
    const float g[3] = {px * inv_dx, py * inv_dx, pz * inv_dx};
    int b[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        b[a] = (int)(g[a] - 0.5f);                      // wp.int truncates toward zero (mpm_utils.py:344-346)
        const float fx = g[a] - (float)b[a];
        W.fx[a] = fx;
        const float wa = 1.5f - fx, wb = fx - 1.0f, wc = fx - 0.5f;
        W.w[a][0] = wa * wa * 0.5f;
        W.w[a][1] = 0.f - wb * wb + 0.75f;
        W.w[a][2] = wc * wc * 0.5f;
        W.dw[a][0] = fx - 1.5f;
        W.dw[a][1] = -2.0f * (fx - 1.0f);
        W.dw[a][2] = fx - 0.5f;
    }
    W.bx = b[0]; W.by = b[1]; W.bz = b[2];
    return W;
}
__device__ __forceinline__ Weights bspline(const DevState& s, float px, float py, float pz) { return bspline_t(s.inv_dx, px, py, pz); }

// ------------------------------------------------------------------------------------------ p2g
// Substep part 1, one thread per particle: pre-p2g particle operations, return mapping, stress.
// Writes v (if a BC changed it), F, stress (and yield_stress / mu / lam where a return map updates them).
// The scatter itself is mpm_scatter_kernel: at 1e5 particles a single kernel that does both is one long dependent
// instruction stream on ~20 warps per SM (r01 ncu: 4600 instructions per thread, issue slots 36 % used).
__global__ void __launch_bounds__(128)
mpm_stress_kernel(const DevState s, const float dt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    const bool live = true;
    const float time = (float)(*s.time);
    float vx = s.v[3 * p], vy = s.v[3 * p + 1], vz = s.v[3 * p + 2];
    const float px = s.x[3 * p], py = s.x[3 * p + 1], pz = s.x[3 * p + 2];
    const float mass = s.mass[p];

    // ---- pre-p2g particle operations: all impulses first, then all velocity modifiers
    //      (mpm_solver_warp.py:528-547)
    bool v_dirty = false;
    for (int k = 0; k < s.n_bc; ++k) {
        const DevBC& bc = s.bcs[k];
        if (bc.kind != PIXIE_BC_IMPULSE) continue;
        if (time >= bc.start_time && time < bc.end_time && bc.mask[p] == 1) {
            vx = vx + (bc.velocity[0] / mass) * dt;      // apply_force :1015-1027 (force stored in velocity[])
            vy = vy + (bc.velocity[1] / mass) * dt;
            vz = vz + (bc.velocity[2] / mass) * dt;
            v_dirty = true;
        }
    }
    for (int k = 0; k < s.n_bc; ++k) {
        const DevBC& bc = s.bcs[k];
        if (bc.kind == PIXIE_BC_VELOCITY_TRANSLATION) {
            if (time >= bc.start_time && time < bc.end_time && bc.mask[p] == 1) {
                vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2];
                v_dirty = true;
            }
        } else if (bc.kind == PIXIE_BC_VELOCITY_ROTATION) {
            if (time >= bc.start_time && time < bc.end_time && bc.mask[p] == 1) {
                // modify_particle_v_before_p2g :1137-1179
                const float ox = px - bc.point[0], oy = py - bc.point[1], oz = pz - bc.point[2];
                const float on = ox * bc.normal[0] + oy * bc.normal[1] + oz * bc.normal[2];
                const float hx = ox - on * bc.normal[0], hy = oy - on * bc.normal[1], hz = oz - on * bc.normal[2];
                const float hd = sqrtf(hx * hx + hy * hy + hz * hz);
                const float cosine = (ox * bc.h1[0] + oy * bc.h1[1] + oz * bc.h1[2]) / hd;
                float theta = acosf(cosine);
                if (!(ox * bc.h2[0] + oy * bc.h2[1] + oz * bc.h2[2] > 0.f)) theta = -theta;
                const float a1 = -hd * sinf(theta) * bc.rotation_scale;
                const float a2 = hd * cosf(theta) * bc.rotation_scale;
                const float av = bc.translation_scale;
                vx = a1 * bc.h1[0] + a2 * bc.h2[0] + av * bc.normal[0];
                vy = a1 * bc.h1[1] + a2 * bc.h2[1] + av * bc.normal[1];
                vz = a1 * bc.h1[2] + a2 * bc.h2[2] + av * bc.normal[2];
                v_dirty = true;
            }
        }
    }
    if (v_dirty && live) { s.v[3 * p] = vx; s.v[3 * p + 1] = vy; s.v[3 * p + 2] = vz; }

    if (s.selection[p] != 0) return;
    const bool contrib = true;

    // ---- compute_stress_from_F_trial (mpm_utils.py:467-526)
    const int material = s.material[p];
    float mu = s.mu[p], lam = s.lam[p];
    const M3 Ft = load_m3(s.F_trial, p);
    M3 F = Ft;
    if (material == 1) {
        float ys = s.yield_stress[p];
        const float ys0 = ys;
        F = return_von_mises(Ft, mu, lam, ys, s.hardening, s.xi, false, 0.f, mu, lam);
        if (ys != ys0 && contrib) s.yield_stress[p] = ys;
    } else if (material == 2) {
        F = return_sand(Ft, mu, lam, s.alpha);
    } else if (material == 3) {
        F = return_viscoplastic(Ft, mu, s.yield_stress[p], s.plastic_viscosity, dt);
    } else if (material == 5) {
        float ys = s.yield_stress[p];
        const float ys0 = ys, mu0 = mu;
        F = return_von_mises(Ft, mu, lam, ys, s.hardening, s.xi, true, s.softening, mu, lam);
        if (ys != ys0 && contrib) s.yield_stress[p] = ys;
        if (mu != mu0 && contrib) { s.mu[p] = mu; s.lam[p] = lam; }
    }
    if (contrib) store_m3(s.F, p, F);
    const float J = m3_det(F);
    M3 tau = m3_zero();
    if (material == 6) {
        tau = stress_water(J, s.bulk[p]);
    } else if (material == 0 || material == 5) {
        // fixed-corotated stress needs only R = U V^T: Newton polar iteration, SVD only if it does not converge
        M3 R;
        if (polar_rotation(F, R)) tau = stress_fcr_R(F, R, J, mu, lam);
        else { M3 U, V; V3 sig; svd3(F, U, sig, V); tau = stress_fcr(F, U, V, J, mu, lam); }
    } else if (material >= 1 && material <= 3) {
        M3 U, V; V3 sig;
        svd3(F, U, sig, V);
        if (material == 1 || material == 3) tau = stress_stvk(F, U, V, sig, mu, lam);
        else tau = stress_drucker_prager(F, U, V, sig, mu, lam);
    }
    {   // enforce symmetry
        const M3 tt = m3_t(tau);
#pragma unroll
        for (int i = 0; i < 9; ++i) tau.m[i] = (tau.m[i] + tt.m[i]) / 2.0f;
    }
    if (contrib) store_m3(s.stress, p, tau);
}

// Substep part 2: scatter of slice i = blockIdx.y of the 3x3x3 stencil (9 nodes) with warp-aggregated atomics.
// Threads walk the particles in cell order (`s.order`), so the lanes of a warp hold runs of particles with the SAME base
// cell = the same target nodes; each run (chopped at 8 lanes) is summed with 3 segmented shuffle steps per value and only
// the run's first lane issues the red.global.add.v4. Runs are found from the keys the lanes compute THIS substep, so a
// stale order costs efficiency, never correctness. (mpm_utils.py:338-394)
template <int kSlices>
__global__ void __launch_bounds__(128)
mpm_scatter_kernel(const DevState s, const float dt) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = tid < s.n;
    const int p = live ? (s.order ? s.order[tid] : tid) : 0;
    const bool contrib = live && s.selection[p] == 0;
    const float vx = s.v[3 * p], vy = s.v[3 * p + 1], vz = s.v[3 * p + 2];
    const float px = s.x[3 * p], py = s.x[3 * p + 1], pz = s.x[3 * p + 2];
    const float mass = s.mass[p];
    const M3 tau = load_m3(s.stress, p);
    const Weights W = bspline(s, px, py, pz);
    M3 C = load_m3(s.C, p);
    {
        const float r = s.rpic_damping;
        const M3 Ct = m3_t(C);
        M3 Cn;
#pragma unroll
        for (int i = 0; i < 9; ++i) Cn.m[i] = (1.0f - r) * C.m[i] + r / 2.0f * (C.m[i] - Ct.m[i]);
        C = (r < -0.001f) ? m3_zero() : Cn;
    }
    const float vol = s.vol[p];
    const int n = s.n_grid;

    // runs of equal base cell among consecutive lanes, chopped at 8
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int key = contrib ? (W.bx * n + W.by) * n + W.bz : -1 - lane;
    const int kprev = __shfl_up_sync(full, key, 1);
    bool head = (lane == 0) || (key != kprev);
    unsigned H = __ballot_sync(full, head);
    const int hl = 31 - __clz(H & (0xffffffffu >> (31 - lane)));     // head lane of my run
    head = head || (((lane - hl) & 7) == 0);
    H = __ballot_sync(full, head);
    const unsigned above = H & ~((2u << lane) - 1u);                  // heads strictly above this lane
    const int seg_end = above ? (__ffs(above) - 2) : 31;              // last lane of my segment
    const bool c1 = lane + 1 <= seg_end, c2 = lane + 2 <= seg_end, c4 = lane + 4 <= seg_end;
    auto segsum = [&](float v) {
        float t = __shfl_down_sync(full, v, 1); if (c1) v += t;
        t = __shfl_down_sync(full, v, 2); if (c2) v += t;
        t = __shfl_down_sync(full, v, 4); if (c4) v += t;
        return v;
    };
#pragma unroll
    for (int ii = 0; ii < (kSlices == 1 ? 3 : 1); ++ii) {
        const int i = kSlices == 1 ? ii : (int)blockIdx.y;
        // slice weights selected without dynamic indexing (keeps W in registers)
        const float w0i = i == 0 ? W.w[0][0] : (i == 1 ? W.w[0][1] : W.w[0][2]);
        const float dw0i = i == 0 ? W.dw[0][0] : (i == 1 ? W.dw[0][1] : W.dw[0][2]);
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ix = W.bx + i, iy = W.by + j, iz = W.bz + k;
                // the reference indexes out of bounds here (no checks); we drop the node. Same verdict for a whole run.
                const bool inb = contrib && (unsigned)ix < (unsigned)n && (unsigned)iy < (unsigned)n && (unsigned)iz < (unsigned)n;
                const V3 dpos = {((float)i - W.fx[0]) * s.dx, ((float)j - W.fx[1]) * s.dx, ((float)k - W.fx[2]) * s.dx};
                const float weight = w0i * W.w[1][j] * W.w[2][k];
                const V3 dweight = {dw0i * W.w[1][j] * W.w[2][k] * s.inv_dx,
                                    w0i * W.dw[1][j] * W.w[2][k] * s.inv_dx,
                                    w0i * W.w[1][j] * W.dw[2][k] * s.inv_dx};
                const V3 sd = m3_mulv(tau, dweight);
                const V3 cd = m3_mulv(C, dpos);
                const float wm = inb ? weight * mass : 0.f;
                float ax = inb ? wm * (vx + cd.x) + dt * (-vol * sd.x) : 0.f;
                float ay = inb ? wm * (vy + cd.y) + dt * (-vol * sd.y) : 0.f;
                float az = inb ? wm * (vz + cd.z) + dt * (-vol * sd.z) : 0.f;
                float aw = wm;
                ax = segsum(ax); ay = segsum(ay); az = segsum(az); aw = segsum(aw);
                if (head && inb) {
                    float* node = reinterpret_cast<float*>(s.grid_mv + ((size_t)ix * n + iy) * n + iz);
                    ptx::red_add_v4(node, ax, ay, az, aw);
                }
            }
    }
}

// ------------------------------------------------------------------------------------------ grid
__global__ void __launch_bounds__(256)
mpm_grid_kernel(const DevState s, const float dt) {
    const int n = s.n_grid;
    const size_t first = (size_t)s.x_begin * n * n, total = (size_t)s.x_end * n * n;
    const size_t idx = first + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int gz = (int)(idx % n), gy = (int)((idx / n) % n), gx = (int)(idx / ((size_t)n * n));
    const float time = (float)(*s.time);
    const float4 mv = s.grid_mv[idx];
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (mv.w > 1e-15f) {                                   // grid_normalization_and_gravity :398-409
        const float inv = 1.0f / mv.w;
        vx = mv.x * inv + dt * s.gx;
        vy = mv.y * inv + dt * s.gy;
        vz = mv.z * inv + dt * s.gz;
    }
    if (s.grid_v_damping_scale < 1.0f) {                   // add_damping_via_grid :583-588 (only if < 1)
        vx *= s.grid_v_damping_scale; vy *= s.grid_v_damping_scale; vz *= s.grid_v_damping_scale;
    }
    for (int k = 0; k < s.n_bc; ++k) {
        const DevBC& bc = s.bcs[k];
        const bool active = time >= bc.start_time && time < bc.end_time;
        if (bc.kind == PIXIE_BC_SURFACE_COLLIDER) {        // :785-840
            if (active) {
                const float ox = (float)gx * s.dx - bc.point[0], oy = (float)gy * s.dx - bc.point[1], oz = (float)gz * s.dx - bc.point[2];
                const float dotp = ox * bc.normal[0] + oy * bc.normal[1] + oz * bc.normal[2];
                if (dotp < 0.0f) {
                    if (bc.surface_type == 11) {
                        const float zz = (float)gz * s.dx;
                        if (zz < 0.4f || zz > 0.53f) { vx = 0.f; vy = 0.f; vz = 0.f; }
                        else { vx = vx * 0.3f; vy = 0.0f * 0.3f; vz = vz * 0.3f; }
                    } else {
                        // sticky -> 0; slip / separate: the reference computes the projected velocity and
                        // then overwrites the node with zero (:838-840)
                        vx = 0.f; vy = 0.f; vz = 0.f;
                    }
                }
            }
        } else if (bc.kind == PIXIE_BC_CUBOID) {           // :874-897
            if (active) {
                const float ox = (float)gx * s.dx - bc.point[0], oy = (float)gy * s.dx - bc.point[1], oz = (float)gz * s.dx - bc.point[2];
                if (fabsf(ox) < bc.size[0] && fabsf(oy) < bc.size[1] && fabsf(oz) < bc.size[2]) {
                    vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2];
                }
            } else if (bc.reset == 1) {
                if (time < bc.end_time + 15.0f * dt) { vx = 0.f; vy = 0.f; vz = 0.f; }
            }
        } else if (bc.kind == PIXIE_BC_BOUNDING_BOX) {     // :917-974
            if (active) {
                const int padding = 3;
                if (gx < padding && vx < 0.f) vx = 0.f;
                if (gx >= n - padding && vx > 0.f) vx = 0.f;
                if (gy < padding && vy < 0.f) vy = 0.f;
                if (gy >= n - padding && vy > 0.f) vy = 0.f;
                if (gz < padding && vz < 0.f) vz = 0.f;
                if (gz >= n - padding && vz > 0.f) vz = 0.f;
            }
        }
    }
    s.grid_v[idx] = make_float4(vx, vy, vz, 0.f);
    if (mv.x != 0.f || mv.y != 0.f || mv.z != 0.f || mv.w != 0.f) s.grid_mv[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------ g2p
__global__ void __launch_bounds__(128)
mpm_g2p_kernel(const DevState s, const float dt, const double dt_d) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < s.n && s.selection[p] == 0) {
        const float px = s.x[3 * p], py = s.x[3 * p + 1], pz = s.x[3 * p + 2];
        const Weights W = bspline(s, px, py, pz);
        const int n = s.n_grid;
        float nvx = 0.f, nvy = 0.f, nvz = 0.f;
        M3 nC = m3_zero(), nF = m3_zero();
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int ix = W.bx + i, iy = W.by + j, iz = W.bz + k;
                    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((unsigned)ix < (unsigned)n && (unsigned)iy < (unsigned)n && (unsigned)iz < (unsigned)n)
                        gv = s.grid_v[((size_t)ix * n + iy) * n + iz];
                    const float dp[3] = {(float)i - W.fx[0], (float)j - W.fx[1], (float)k - W.fx[2]};
                    const float weight = W.w[0][i] * W.w[1][j] * W.w[2][k];
                    const float dwv[3] = {W.dw[0][i] * W.w[1][j] * W.w[2][k] * s.inv_dx,
                                          W.w[0][i] * W.dw[1][j] * W.w[2][k] * s.inv_dx,
                                          W.w[0][i] * W.w[1][j] * W.dw[2][k] * s.inv_dx};
                    nvx = nvx + gv.x * weight; nvy = nvy + gv.y * weight; nvz = nvz + gv.z * weight;
                    const float cw = weight * s.inv_dx * 4.0f;
                    const float g3[3] = {gv.x, gv.y, gv.z};
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            nC.m[3 * r + c] = nC.m[3 * r + c] + (g3[r] * dp[c]) * cw;
                            nF.m[3 * r + c] = nF.m[3 * r + c] + g3[r] * dwv[c];
                        }
                }
        s.v[3 * p] = nvx; s.v[3 * p + 1] = nvy; s.v[3 * p + 2] = nvz;
        s.x[3 * p] = px + dt * nvx; s.x[3 * p + 1] = py + dt * nvy; s.x[3 * p + 2] = pz + dt * nvz;
        store_m3(s.C, p, nC);
        M3 A = m3_ident();
#pragma unroll
        for (int i = 0; i < 9; ++i) A.m[i] += nF.m[i] * dt;
        store_m3(s.F_trial, p, m3_mul(A, load_m3(s.F, p)));
        if (s.update_cov_with_F) {                          // update_cov :315-335
            float* cv = s.cov + (size_t)p * 6;
            M3 cn;
            cn.m[0] = cv[0]; cn.m[1] = cv[1]; cn.m[2] = cv[2]; cn.m[3] = cv[1]; cn.m[4] = cv[3]; cn.m[5] = cv[4];
            cn.m[6] = cv[2]; cn.m[7] = cv[4]; cn.m[8] = cv[5];
            const M3 a = m3_mul(nF, cn), b = m3_mul_t(cn, nF);
            M3 c1;
#pragma unroll
            for (int i = 0; i < 9; ++i) c1.m[i] = cn.m[i] + dt * (a.m[i] + b.m[i]);
            cv[0] = c1.m[0]; cv[1] = c1.m[1]; cv[2] = c1.m[2]; cv[3] = c1.m[4]; cv[4] = c1.m[5]; cv[5] = c1.m[8];
        }
    }
    // ---- substep epilogue: nothing in this kernel reads the clock or the BC table
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double t = *s.time;
        for (int k = 0; k < s.n_bc; ++k) {
            DevBC& bc = s.bcs[k];
            if (bc.kind == PIXIE_BC_CUBOID && t >= (double)bc.start_time && t < (double)bc.end_time) {
                // modify(): Python-float arithmetic, stored back as fp32 (mpm_solver_warp.py:899-905)
                bc.point[0] = (float)((double)bc.point[0] + dt_d * (double)bc.velocity[0]);
                bc.point[1] = (float)((double)bc.point[1] + dt_d * (double)bc.velocity[1]);
                bc.point[2] = (float)((double)bc.point[2] + dt_d * (double)bc.velocity[2]);
            }
        }
        *s.time = t + dt_d;
    }
}

// ------------------------------------------------------------------------------------------ setup kernels
__global__ void mpm_mu_lam_kernel(const DevState s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    const float E = s.E[p], nu = s.nu[p];
    s.mu[p] = E / (2.0f * (1.0f + nu));
    s.lam[p] = E * nu / ((1.0f + nu) * (1.0f - 2.0f * nu));
}
__global__ void mpm_bulk_kernel(const DevState s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    s.bulk[p] = s.lam[p] + 2.f / 3.f * s.mu[p];
}
__global__ void mpm_mass_kernel(const DevState s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    s.mass[p] = s.density[p] * s.vol[p];
}
__global__ void mpm_cov_from_F_kernel(const DevState s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    const M3 F = load_m3(s.F_trial, p);
    const float* ic = s.init_cov + (size_t)p * 6;
    M3 c0;
    c0.m[0] = ic[0]; c0.m[1] = ic[1]; c0.m[2] = ic[2]; c0.m[3] = ic[1]; c0.m[4] = ic[3]; c0.m[5] = ic[4];
    c0.m[6] = ic[2]; c0.m[7] = ic[4]; c0.m[8] = ic[5];
    const M3 c = m3_mul_t(m3_mul(F, c0), F);
    float* o = s.cov + (size_t)p * 6;
    o[0] = c.m[0]; o[1] = c.m[1]; o[2] = c.m[2]; o[3] = c.m[4]; o[4] = c.m[5]; o[5] = c.m[8];
}
__global__ void mpm_R_from_F_kernel(const DevState s) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    const M3 F = load_m3(s.F_trial, p);
    M3 U, V; V3 sig;
    svd3(F, U, sig, V);
    // svd3 already returns proper rotations, so the det < 0 fix-ups of compute_R_from_F (:568-576) are no-ops
    const M3 R = m3_mul_t(U, V);
    store_m3(s.R, p, m3_t(R));
}
__global__ void mpm_additional_params_kernel(const DevState s, const float* __restrict__ boxes, int n_boxes) {
    __shared__ float sb[128 * 10];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    float px = 0, py = 0, pz = 0;
    if (p < s.n) { px = s.x[3 * p]; py = s.x[3 * p + 1]; pz = s.x[3 * p + 2]; }
    int hit = -1;
    float hv[4] = {0, 0, 0, 0};
    for (int b0 = 0; b0 < n_boxes; b0 += 128) {
        const int nb = min(128, n_boxes - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * 10; i += blockDim.x) sb[i] = boxes[(size_t)b0 * 10 + i];
        __syncthreads();
        for (int b = 0; b < nb; ++b) {
            const float* B = sb + b * 10;
            if (px > B[0] - B[3] && px < B[0] + B[3] && py > B[1] - B[4] && py < B[1] + B[4] &&
                pz > B[2] - B[5] && pz < B[2] + B[5]) {
                hit = b0 + b; hv[0] = B[6]; hv[1] = B[7]; hv[2] = B[8]; hv[3] = B[9];   // later boxes override
            }
        }
    }
    if (p < s.n && hit >= 0) {
        s.E[p] = hv[0]; s.nu[p] = hv[1]; s.density[p] = hv[2]; s.material[p] = (int)hv[3];
    }
}
__global__ void mpm_select_box_kernel(const DevState s, float3 point, float3 size, int* mask) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    const float ox = s.x[3 * p] - point.x, oy = s.x[3 * p + 1] - point.y, oz = s.x[3 * p + 2] - point.z;
    mask[p] = (fabsf(ox) < size.x && fabsf(oy) < size.y && fabsf(oz) < size.z) ? 1 : 0;
}
__global__ void mpm_select_cyl_kernel(const DevState s, float3 point, float3 normal, float half_height, float radius, int* mask) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= s.n) return;
    const float ox = s.x[3 * p] - point.x, oy = s.x[3 * p + 1] - point.y, oz = s.x[3 * p + 2] - point.z;
    const float on = ox * normal.x + oy * normal.y + oz * normal.z;
    const float vd = fabsf(on);
    const float hx = ox - on * normal.x, hy = oy - on * normal.y, hz = oz - on * normal.z;
    const float hd = sqrtf(hx * hx + hy * hy + hz * hz);
    mask[p] = (vd < half_height && hd < radius) ? 1 : 0;
}

#include "mpm_fused.cuh"

// base-cell key of every live particle (+ identity index), input of the radix sort that produces DevState::order
__global__ void mpm_cell_key_kernel(const float* __restrict__ x, int n, float inv_dx, int n_grid, int* __restrict__ keys, int* __restrict__ idx) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const Weights W = bspline_t(inv_dx, x[3 * p], x[3 * p + 1], x[3 * p + 2]);
    const int bx = min(max(W.bx, 0), n_grid - 1), by = min(max(W.by, 0), n_grid - 1), bz = min(max(W.bz, 0), n_grid - 1);
    keys[p] = (bx * n_grid + by) * n_grid + bz;
    idx[p] = p;
}

}  // namespace

// ============================================================================================ host
struct Mpm {
    int n = 0, n_grid = 0;
    int n_active = 0;                  // particles [0, n_active) are live (slab mode migrates particles between ranks)
    int x_begin = 0, x_end = 0;        // grid planes this instance updates ([0, n_grid) unless slab-decomposed)
    bool grid_borrowed = false;        // grid_mv belongs to the caller (pixie_mpm_bind_grid)
    float grid_lim = 1.f;
    void* fields[PIXIE_MPM_FIELD_COUNT] = {nullptr};
    pixie_mpm_params params{};
    std::vector<DevBC> bcs;
    float4* grid_mv = nullptr;
    float4* grid_v = nullptr;
    double* d_time = nullptr;          // direct path clock
    DevBC* d_bcs = nullptr;
    // direct path: CUDA graph of a batch of substeps, keyed by dt and the state snapshot it was captured with
    cudaGraphExec_t graph = nullptr;
    double graph_dt = 0;
    bool graph_valid = false;
    // fused path: a few cached graphs keyed by (substep count, clock parity, dt)
    static constexpr int kGraphSlots = 4;
    struct GraphSlot { cudaGraphExec_t exec = nullptr; int count = 0, parity = 0, launches = 0; double dt = 0; } graphs[kGraphSlots];
    int graph_next = 0;
    std::string error;

    // ---- cell order (radix sort of base-cell keys): indirection of the direct path, physical order of the fused path
    int *cell_order = nullptr, *cell_keys = nullptr, *cell_keys_sorted = nullptr, *cell_idx = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    bool order_valid = false;
    int steps_since_order = 0;

    // ---- fused path (mpm_fused.cuh): private cell-sorted SoA copy of the particle state
    bool fused = true;                 // false: four-kernel path on the caller's arrays (slab-decomposed runs, PIXIE_MPM_DIRECT=1)
    struct FsBuf { float* f = nullptr; int *material = nullptr, *selection = nullptr, *perm = nullptr; } fs[2];
    int cap = 0;
    int* d_box = nullptr;              // [6] node box swept by the grid kernel
    double* tslots = nullptr;          // [2] clock, by substep parity
    float* pts = nullptr;              // [2][kMaxBC][3] collider points (the cuboid ones move), by substep parity
    int tpar = 0;
    bool internal_valid = false;       // sorted state mirrors the caller's arrays (+ steps taken since)
    bool user_stale = false;           // sorted state is ahead of the caller's arrays
    int steps_since_sort = 0;
    // ---- slab-decomposed runs on the fused path: exchange buffer = [SlabFlags][grid_mv], neighbours' buffers, overlap totals
    uint8_t* xbuf = nullptr;           // owns grid_mv (+ the flags block in front of it): one allocation, one IPC handle
    bool slab = false;
    int slab_x0 = 0, slab_x1 = 0, slab_slack = 1;
    const uint8_t* peer_xbuf[2] = {nullptr, nullptr};
    float4* ov_total[2] = {nullptr, nullptr};
    int ov_lo[2] = {0, 0}, ov_hi[2] = {0, 0};
    bool g2p_pending = false;          // slab phases: the gather of the last finished substep has not run yet
    float slab_dt = 0.f;               // ... and the dt it has to use
    int agg = 2;                       // log2 of the longest aggregated run in the scatter (PIXIE_MPM_AGG): 2 measured best at 100k/64^3
    long long launches = 0;            // kernels of this library launched for this handle (bench.py's gpu_launches)
};

static constexpr int kGraphSteps = 25;         // direct path
static constexpr int kFusedGraphSteps = 50;    // fused path: substeps per graph replay
static constexpr int kMinGraphSteps = 4;       // shorter batches are launched directly
static constexpr int kResortEvery = 100;       // substeps between re-sorts; CFL keeps a particle within ~a cell of its slot far longer
static constexpr int kBoxMargin = 2;           // nodes added around the particles' node box at every sort

static DevState make_state(Mpm* m) {
    DevState s{};
    auto f = [&](int id) { return reinterpret_cast<float*>(m->fields[id]); };
    s.x = f(PIXIE_MPM_X); s.v = f(PIXIE_MPM_V); s.F = f(PIXIE_MPM_F); s.F_trial = f(PIXIE_MPM_F_TRIAL);
    s.C = f(PIXIE_MPM_C); s.stress = f(PIXIE_MPM_STRESS); s.R = f(PIXIE_MPM_R); s.cov = f(PIXIE_MPM_COV);
    s.init_cov = f(PIXIE_MPM_INIT_COV); s.vol = f(PIXIE_MPM_VOL); s.mass = f(PIXIE_MPM_MASS);
    s.density = f(PIXIE_MPM_DENSITY); s.E = f(PIXIE_MPM_E); s.nu = f(PIXIE_MPM_NU); s.mu = f(PIXIE_MPM_MU);
    s.lam = f(PIXIE_MPM_LAM); s.bulk = f(PIXIE_MPM_BULK); s.yield_stress = f(PIXIE_MPM_YIELD);
    s.material = reinterpret_cast<int*>(m->fields[PIXIE_MPM_MATERIAL]);
    s.selection = reinterpret_cast<int*>(m->fields[PIXIE_MPM_SELECTION]);
    s.order = m->order_valid ? m->cell_order : nullptr;
    s.grid_mv = m->grid_mv; s.grid_v = m->grid_v; s.time = m->d_time; s.bcs = m->d_bcs; s.n_bc = (int)m->bcs.size();
    s.n = m->n_active; s.n_grid = m->n_grid;
    s.x_begin = m->x_begin; s.x_end = m->x_end;
    // dx, inv_dx exactly as mpm_solver_warp.py:61-66 (Python doubles rounded to fp32 members)
    s.dx = (float)((double)m->grid_lim / (double)m->n_grid);
    s.inv_dx = (float)((double)m->n_grid / (double)m->grid_lim);
    const pixie_mpm_params& q = m->params;
    s.gx = q.gravity[0]; s.gy = q.gravity[1]; s.gz = q.gravity[2];
    s.rpic_damping = q.rpic_damping; s.grid_v_damping_scale = q.grid_v_damping_scale; s.alpha = q.alpha;
    s.hardening = q.hardening; s.xi = q.xi; s.plastic_viscosity = q.plastic_viscosity; s.softening = q.softening;
    s.update_cov_with_F = q.update_cov_with_F;
    s.scatter_slices = 1;
    return s;
}

void mpm_destroy(Mpm* m);
int mpm_sync(Mpm* m, cudaStream_t st);
static void fused_launch(Mpm* m, bool do_g2p, bool do_p2g, bool write_all, float dt, cudaStream_t st);

// ---------------------------------------------------------------------------------- sort scratch (both paths)
static int sort_alloc(Mpm* m) {
    if (m->cell_order) return 0;
    const size_t cap = (size_t)m->n;
    if (cudaMalloc(&m->cell_order, cap * sizeof(int)) != cudaSuccess || cudaMalloc(&m->cell_keys, cap * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&m->cell_keys_sorted, cap * sizeof(int)) != cudaSuccess || cudaMalloc(&m->cell_idx, cap * sizeof(int)) != cudaSuccess) {
        m->error = "cudaMalloc failed (cell order)"; return 1;
    }
    cub::DeviceRadixSort::SortPairs(nullptr, m->cub_bytes, m->cell_keys, m->cell_keys_sorted, m->cell_idx, m->cell_order, (int)cap, 0, 32, 0);
    if (cudaMalloc(&m->cub_tmp, m->cub_bytes) != cudaSuccess) { m->error = "cudaMalloc failed (sort scratch)"; return 1; }
    return 0;
}
static int key_bits(const Mpm* m) {
    int bits = 1;
    while ((1ll << bits) < (long long)m->n_grid * m->n_grid * m->n_grid) ++bits;
    return bits;
}

// ---------------------------------------------------------------------------------- fused path: host
static FsUser fs_user(Mpm* m) {
    FsUser u{};
    auto f = [&](int id) { return reinterpret_cast<float*>(m->fields[id]); };
    u.x = f(PIXIE_MPM_X); u.v = f(PIXIE_MPM_V); u.C = f(PIXIE_MPM_C); u.F = f(PIXIE_MPM_F); u.Ft = f(PIXIE_MPM_F_TRIAL);
    u.stress = f(PIXIE_MPM_STRESS); u.mass = f(PIXIE_MPM_MASS); u.vol = f(PIXIE_MPM_VOL); u.mu = f(PIXIE_MPM_MU); u.lam = f(PIXIE_MPM_LAM);
    u.bulk = f(PIXIE_MPM_BULK); u.ys = f(PIXIE_MPM_YIELD); u.cov = f(PIXIE_MPM_COV);
    u.material = reinterpret_cast<int*>(m->fields[PIXIE_MPM_MATERIAL]);
    u.selection = reinterpret_cast<int*>(m->fields[PIXIE_MPM_SELECTION]);
    return u;
}

static int fused_alloc(Mpm* m) {
    if (m->fs[0].f) return 0;
    m->cap = (m->n + 31) / 32 * 32;
    const size_t cap = (size_t)m->cap;
    bool ok = true;
    for (int b = 0; b < 2 && ok; ++b) {
        Mpm::FsBuf& s = m->fs[b];
        ok = cudaMalloc(&s.f, (size_t)FS_NFLOAT * cap * sizeof(float)) == cudaSuccess && cudaMalloc(&s.material, cap * sizeof(int)) == cudaSuccess &&
             cudaMalloc(&s.selection, cap * sizeof(int)) == cudaSuccess && cudaMalloc(&s.perm, cap * sizeof(int)) == cudaSuccess;
        if (ok) {
            cudaMemset(s.f, 0, (size_t)FS_NFLOAT * cap * sizeof(float));
            cudaMemset(s.material, 0, cap * sizeof(int)); cudaMemset(s.selection, 0, cap * sizeof(int)); cudaMemset(s.perm, 0, cap * sizeof(int));
        }
    }
    if (!ok) { m->error = "cudaMalloc failed (sorted particle state)"; return 1; }
    return sort_alloc(m);
}

// node box of the particles at `x` (+ margin) into m->d_box
static void fused_box(Mpm* m, const float* x, long long stride_comp, long long stride_part, cudaStream_t st) {
    const int init[6] = {m->n_grid, m->n_grid, m->n_grid, 0, 0, 0};
    cudaMemcpyAsync(m->d_box, init, sizeof(init), cudaMemcpyHostToDevice, st);
    const float inv_dx = (float)((double)m->n_grid / (double)m->grid_lim);
    fs_box_kernel<<<148, 256, 0, st>>>(x, stride_comp, stride_part, m->n_active, inv_dx, m->n_grid, kBoxMargin, m->d_box, 0);
    fs_box_kernel<<<1, 32, 0, st>>>(x, stride_comp, stride_part, m->n_active, inv_dx, m->n_grid, kBoxMargin, m->d_box, 1);
    m->launches += 2;
}

static int fused_sort(Mpm* m, const float* x, long long stride_comp, long long stride_part, cudaStream_t st) {
    m->steps_since_sort = 0;
    if (m->n_active <= 0) return 0;
    const float inv_dx = (float)((double)m->n_grid / (double)m->grid_lim);
    fs_key_kernel<<<(m->n_active + 255) / 256, 256, 0, st>>>(x, stride_comp, stride_part, m->n_active, inv_dx, m->n_grid, m->cell_keys, m->cell_idx);
    size_t bytes = m->cub_bytes;
    if (cub::DeviceRadixSort::SortPairs(m->cub_tmp, bytes, m->cell_keys, m->cell_keys_sorted, m->cell_idx, m->cell_order, m->n_active, 0, key_bits(m), st) != cudaSuccess) {
        m->error = "radix sort failed"; return 1;
    }
    m->steps_since_sort = 0;
    m->launches += 1;          // + the radix sort passes of cub (library kernels, not counted)
    return 0;
}

// caller's arrays -> sorted state
static int fused_gather_from_user(Mpm* m, cudaStream_t st) {
    if (fused_alloc(m)) return 1;
    const FsUser u = fs_user(m);
    if (fused_sort(m, u.x, 1, 3, st)) return 1;
    Mpm::FsBuf& d = m->fs[0];
    if (m->n_active > 0) fs_gather_kernel<<<(m->n_active + 255) / 256, 256, 0, st>>>(u, m->cell_order, m->n_active, m->cap, d.f, d.material, d.selection, d.perm,
                                                          m->params.update_cov_with_F ? 1 : 0);
    fused_box(m, u.x, 1, 3, st);
    m->launches += 1;
    m->internal_valid = true;
    m->user_stale = false;
    return cudaGetLastError() != cudaSuccess;
}

// re-sort of the live sorted state (between two launches of the particle kernel)
static int fused_resort(Mpm* m, cudaStream_t st) {
    Mpm::FsBuf& a = m->fs[0];
    Mpm::FsBuf& b = m->fs[1];
    if (fused_sort(m, a.f + (size_t)FS_X * m->cap, m->cap, 1, st)) return 1;
    if (m->n_active <= 0) return 0;
    fs_permute_kernel<<<(m->n_active + 255) / 256, 256, 0, st>>>(a.f, a.material, a.selection, a.perm, m->cell_order, m->n_active, m->cap, b.f, b.material,
                                                           b.selection, b.perm);
    // back into buffer 0: the captured graph and the launch arguments keep pointing at it
    const size_t cap = (size_t)m->cap;
    cudaMemcpyAsync(a.f, b.f, (size_t)FS_NFLOAT * cap * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(a.material, b.material, cap * sizeof(int), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(a.selection, b.selection, cap * sizeof(int), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(a.perm, b.perm, cap * sizeof(int), cudaMemcpyDeviceToDevice, st);
    fused_box(m, a.f + (size_t)FS_X * m->cap, m->cap, 1, st);
    m->launches += 1;
    return cudaGetLastError() != cudaSuccess;
}

// sorted state -> caller's arrays (if it is ahead); afterwards the caller may mutate its arrays, so the sorted copy is
// considered out of date.
int mpm_sync(Mpm* m, cudaStream_t st) {
    if (!m->fused) return 0;
    if (m->g2p_pending && m->internal_valid) {        // slab phases: finish the last substep (gather) before anything is read
        fused_launch(m, true, false, true, m->slab_dt, st);
        m->g2p_pending = false;
    }
    if (m->user_stale) {
        const Mpm::FsBuf& s = m->fs[0];
        if (m->n_active > 0) fs_unsort_kernel<<<(m->n_active + 255) / 256, 256, 0, st>>>(fs_user(m), s.perm, m->n_active, m->cap, s.f, m->params.update_cov_with_F ? 1 : 0);
        m->user_stale = false;
        m->launches += 1;
        if (cudaGetLastError() != cudaSuccess) { m->error = "unsort launch failed"; return 1; }
    }
    m->internal_valid = false;
    return 0;
}

// Launch with programmatic stream serialisation (PDL): the kernel may be scheduled while its predecessor in the stream
// (or captured graph) is still draining; both kernels of the substep chain wait for it with griddepcontrol.wait.
template <typename... KArgs, typename... Args>
static void pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    static const bool pdl = !(getenv("PIXIE_MPM_PDL") && atoi(getenv("PIXIE_MPM_PDL")) == 0);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static FusedState fused_state(Mpm* m) {
    FusedState t{};
    const Mpm::FsBuf& s = m->fs[0];
    t.f = s.f; t.material = s.material; t.selection = s.selection; t.perm = s.perm;
    t.cap = m->cap; t.n = m->n_active;
    t.grid_v = m->grid_v; t.grid_mv = m->grid_mv; t.box = m->d_box;
    t.bcs = m->d_bcs; t.n_bc = (int)m->bcs.size();
    t.n_particle_bc = 0;
    for (const DevBC& b : m->bcs) {
        if (b.kind < PIXIE_BC_IMPULSE) continue;
        if (t.n_particle_bc < kInlinePBC) {
            ParticleBC& q = t.pbc[t.n_particle_bc];
            q.kind = b.kind; q.start_time = b.start_time; q.end_time = b.end_time; q.mask = b.mask;
            q.rotation_scale = b.rotation_scale; q.translation_scale = b.translation_scale;
            for (int a = 0; a < 3; ++a) { q.velocity[a] = b.velocity[a]; q.point[a] = b.point[a]; q.normal[a] = b.normal[a]; q.h1[a] = b.h1[a]; q.h2[a] = b.h2[a]; }
        }
        ++t.n_particle_bc;
    }
    t.n_pbc_inline = t.n_particle_bc <= kInlinePBC ? t.n_particle_bc : -1;
    t.n_grid = m->n_grid;
    t.dx = (float)((double)m->grid_lim / (double)m->n_grid);
    t.inv_dx = (float)((double)m->n_grid / (double)m->grid_lim);
    const pixie_mpm_params& q = m->params;
    t.rpic_damping = q.rpic_damping; t.alpha = q.alpha; t.hardening = q.hardening; t.xi = q.xi;
    t.plastic_viscosity = q.plastic_viscosity; t.softening = q.softening;
    t.update_cov_with_F = q.update_cov_with_F;
    if (m->slab) {
        SlabFlags* fl = reinterpret_cast<SlabFlags*>(m->xbuf);
        t.slab_step = &fl->step; t.slab_err = &fl->error;
        t.base_lo = m->peer_xbuf[0] ? m->slab_x0 - m->slab_slack : -(1 << 30);
        t.base_hi = m->peer_xbuf[1] ? m->slab_x1 + m->slab_slack : (1 << 30);
    }
    return t;
}

static void fused_launch(Mpm* m, bool do_g2p, bool do_p2g, bool write_all, float dt, cudaStream_t st) {
    FusedState t = fused_state(m);
    t.do_g2p = do_g2p; t.do_p2g = do_p2g; t.write_all = write_all;
    t.time = m->tslots + m->tpar;                 // clock of the substep whose stress / scatter runs in this launch
    // 88 registers per thread: 32-thread blocks pack 23 per SM (736 threads), so 100k particles are one wave on 148 SMs
    static const int B = [] { const char* e = getenv("PIXIE_MPM_FUSED_BLOCK"); const int b = e ? atoi(e) : kFusedThreads; return (b == 32 || b == 64 || b == 128) ? b : kFusedThreads; }();
    const int blocks = (std::max(m->n_active, 1) + B - 1) / B;
    static const bool hoist = !(getenv("PIXIE_MPM_HOIST") && atoi(getenv("PIXIE_MPM_HOIST")) == 0);   // r02 A/B: 23.5 vs 24.1 us
    void (*kern)(const FusedState, const float) = mpm_fused_kernel<2, false>;
    if (hoist) kern = m->agg == 1 ? mpm_fused_kernel<1, true> : (m->agg == 3 ? mpm_fused_kernel<3, true> : mpm_fused_kernel<2, true>);
    else kern = m->agg == 0 ? mpm_fused_kernel<0, false> : (m->agg == 1 ? mpm_fused_kernel<1, false> : (m->agg == 3 ? mpm_fused_kernel<3, false> : mpm_fused_kernel<2, false>));
    pdl_launch(kern, dim3(blocks), dim3(B), st, t, dt);
    m->launches += 1;
}

static void gridbox_launch(Mpm* m, float dt, double dt_d, cudaStream_t st) {
    GridBoxArgs g{};
    g.grid_mv = m->grid_mv; g.grid_v = m->grid_v; g.box = m->d_box;
    g.time_in = m->tslots + m->tpar; g.time_out = m->tslots + (m->tpar ^ 1);
    g.pts_in = m->pts + (size_t)m->tpar * kMaxBC * 3; g.pts_out = m->pts + (size_t)(m->tpar ^ 1) * kMaxBC * 3;
    g.bcs = m->d_bcs; g.n_bc = (int)m->bcs.size();
    g.n_grid = m->n_grid; g.x_begin = m->x_begin; g.x_end = m->x_end;
    if (m->slab) {
        g.mine = reinterpret_cast<SlabFlags*>(m->xbuf);
        for (int sd = 0; sd < 2; ++sd) {
            g.peer[sd] = reinterpret_cast<const SlabFlags*>(m->peer_xbuf[sd]);
            g.total[sd] = m->ov_total[sd]; g.ov_lo[sd] = m->ov_lo[sd]; g.ov_hi[sd] = m->ov_hi[sd];
        }
    }
    g.dx = (float)((double)m->grid_lim / (double)m->n_grid);
    const pixie_mpm_params& q = m->params;
    g.gx = q.gravity[0]; g.gy = q.gravity[1]; g.gz = q.gravity[2]; g.grid_v_damping_scale = q.grid_v_damping_scale;
    // enough blocks for two per SM; the kernel strides over the (usually much smaller than n_grid^3) node box
    pdl_launch(mpm_gridbox_kernel, dim3(296), dim3(256), st, g, dt, dt_d);
    m->launches += 1;
    m->tpar ^= 1;
}

static void halo_launch(Mpm* m, cudaStream_t st) {
    HaloArgs a{};
    a.mine = reinterpret_cast<SlabFlags*>(m->xbuf);
    a.grid_mv = m->grid_mv; a.box = m->d_box; a.n_grid = m->n_grid;
    for (int sd = 0; sd < 2; ++sd) {
        a.peer[sd] = reinterpret_cast<const SlabFlags*>(m->peer_xbuf[sd]);
        a.peer_mv[sd] = m->peer_xbuf[sd] ? reinterpret_cast<const float4*>(m->peer_xbuf[sd] + sizeof(SlabFlags)) : nullptr;
        a.total[sd] = m->ov_total[sd]; a.ov_lo[sd] = m->ov_lo[sd]; a.ov_hi[sd] = m->ov_hi[sd];
    }
    pdl_launch(mpm_halo_kernel, dim3(148), dim3(256), st, a);
    m->launches += 1;
}

// `count` substeps as: scatter(0) | [halo(0)] | grid(0) | g2p(0)+scatter(1) | ... | grid(count-1) | g2p(count-1)
static void fused_batch(Mpm* m, int count, float dt, double dt_d, cudaStream_t st) {
    fused_launch(m, false, true, count == 1 || m->slab, dt, st);
    for (int i = 0; i < count; ++i) {
        if (m->slab) halo_launch(m, st);
        gridbox_launch(m, dt, dt_d, st);
        if (i + 1 < count) fused_launch(m, true, true, i + 2 == count || m->slab, dt, st);
        else fused_launch(m, true, false, true, dt, st);
    }
}

// CUDA graph of `count` substeps starting at clock parity `m->tpar` (cached: the 50-substep batch of long rollouts and, for
// slab runs, the chunk between two particle migrations)
static cudaGraphExec_t fused_graph(Mpm* m, int count, float dt, double dt_d) {
    for (auto& g : m->graphs)
        if (g.exec && g.count == count && g.parity == m->tpar && g.dt == dt_d) return g.exec;
    Mpm::GraphSlot& slot = m->graphs[m->graph_next];
    m->graph_next = (m->graph_next + 1) % Mpm::kGraphSlots;
    if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
    cudaStream_t cs;
    cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    cudaGraph_t g = nullptr;
    const int par0 = m->tpar;
    const long long launches0 = m->launches;
    bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        fused_batch(m, count, dt, dt_d, cs);
        ok = cudaStreamEndCapture(cs, &g) == cudaSuccess && g;
    }
    slot.launches = (int)(m->launches - launches0);
    m->tpar = par0;                                    // capture did not run anything
    m->launches = launches0;
    if (ok) ok = cudaGraphInstantiate(&slot.exec, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if (!ok) { cudaGetLastError(); slot.exec = nullptr; return nullptr; }
    slot.count = count; slot.parity = par0; slot.dt = dt_d;
    return slot.exec;
}

static int mpm_step_fused(Mpm* m, int n_substeps, double dt_d, cudaStream_t st) {
    const float dt = (float)dt_d;
    if (!m->internal_valid && fused_gather_from_user(m, st)) return 1;
    if (!m->graph_valid) {                              // parameters / BCs / bindings changed: captured launches are stale
        for (auto& g : m->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        m->graph_valid = true;
    }
    int done = 0;
    while (done < n_substeps) {
        if (m->steps_since_sort >= kResortEvery && fused_resort(m, st)) return 1;
        int count = std::min(n_substeps - done, kFusedGraphSteps);
        count = std::min(count, std::max(1, kResortEvery - m->steps_since_sort));
        cudaGraphExec_t g = count >= kMinGraphSteps ? fused_graph(m, count, dt, dt_d) : nullptr;
        if (g) {
            if (cudaGraphLaunch(g, st) != cudaSuccess) { m->error = "cudaGraphLaunch failed"; return 1; }
            for (auto& sl : m->graphs) if (sl.exec == g) m->launches += sl.launches;
            if (count & 1) m->tpar ^= 1;               // the replay advanced the clock `count` times
        } else {
            fused_batch(m, count, dt, dt_d, st);
        }
        done += count;
        m->steps_since_sort += count;
    }
    m->user_stale = true;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { m->error = std::string("kernel launch failed: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

// leave the fused path for good (slab-decomposed runs drive the four-kernel path on the caller's arrays)
static int switch_to_direct(Mpm* m) {
    if (!m->fused) return 0;
    if (mpm_sync(m, 0)) return 1;
    cudaDeviceSynchronize();
    double t = 0;
    cudaMemcpy(&t, m->tslots + m->tpar, sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(m->d_time, &t, sizeof(double), cudaMemcpyHostToDevice);
    // moved collider points back into the BC table the direct kernels read
    std::vector<float> pts((size_t)kMaxBC * 3);
    cudaMemcpy(pts.data(), m->pts + (size_t)m->tpar * kMaxBC * 3, pts.size() * sizeof(float), cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < m->bcs.size(); ++k) {
        for (int a = 0; a < 3; ++a) m->bcs[k].point[a] = pts[3 * k + a];
        cudaMemcpy(m->d_bcs + k, &m->bcs[k], sizeof(DevBC), cudaMemcpyHostToDevice);
    }
    m->fused = false;
    m->graph_valid = false;
    return 0;
}

Mpm* mpm_create(int n_particles, int n_grid, float grid_lim, std::string& err) {
    if (n_particles <= 0 || n_grid <= 0) { err = "n_particles and n_grid must be positive"; return nullptr; }
    auto* m = new Mpm();
    m->n = n_particles; m->n_active = n_particles; m->n_grid = n_grid; m->grid_lim = grid_lim;
    m->x_begin = 0; m->x_end = n_grid;
    m->params.n_grid = n_grid; m->params.grid_lim = grid_lim;
    m->params.grid_v_damping_scale = 1.1f;                 // mpm_solver_warp.py:92
    {
        // friction_angle 25 deg default (:83-86), evaluated like the reference (float math on 3.14159265)
        const double sin_phi = sin(25.0 / 180.0 * 3.14159265);
        m->params.alpha = (float)(sqrt(2.0 / 3.0) * 2.0 * sin_phi / (3.0 - sin_phi));
    }
    m->params.softening = 0.1f;
    const size_t nodes = (size_t)n_grid * n_grid * n_grid;
    if (cudaMalloc(&m->xbuf, sizeof(SlabFlags) + nodes * sizeof(float4)) != cudaSuccess ||
        cudaMalloc(&m->grid_v, nodes * sizeof(float4)) != cudaSuccess ||
        cudaMalloc(&m->d_time, sizeof(double)) != cudaSuccess ||
        cudaMalloc(&m->d_bcs, kMaxBC * sizeof(DevBC)) != cudaSuccess ||
        cudaMalloc(&m->d_box, 6 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&m->tslots, 2 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&m->pts, (size_t)2 * kMaxBC * 3 * sizeof(float)) != cudaSuccess) {
        err = "cudaMalloc failed (no CUDA device?)";
        delete m;
        return nullptr;
    }
    m->grid_mv = reinterpret_cast<float4*>(m->xbuf + sizeof(SlabFlags));
    cudaMemset(m->xbuf, 0, sizeof(SlabFlags) + nodes * sizeof(float4));
    cudaMemset(m->grid_v, 0, nodes * sizeof(float4));
    cudaMemset(m->d_time, 0, sizeof(double));
    cudaMemset(m->tslots, 0, 2 * sizeof(double));
    cudaMemset(m->pts, 0, (size_t)2 * kMaxBC * 3 * sizeof(float));
    m->fused = getenv("PIXIE_MPM_DIRECT") == nullptr;
    if (const char* a = getenv("PIXIE_MPM_AGG")) m->agg = std::min(3, std::max(0, atoi(a)));
    return m;
}

void mpm_destroy(Mpm* m) {
    if (!m) return;
    cudaFree(m->cell_order); cudaFree(m->cell_keys); cudaFree(m->cell_keys_sorted); cudaFree(m->cell_idx); cudaFree(m->cub_tmp);
    for (int b = 0; b < 2; ++b) { cudaFree(m->fs[b].f); cudaFree(m->fs[b].material); cudaFree(m->fs[b].selection); cudaFree(m->fs[b].perm); }
    cudaFree(m->d_box); cudaFree(m->tslots); cudaFree(m->pts);
    if (m->graph) cudaGraphExecDestroy(m->graph);
    for (auto& g : m->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    cudaFree(m->xbuf);
    cudaFree(m->ov_total[0]); cudaFree(m->ov_total[1]);
    cudaFree(m->grid_v); cudaFree(m->d_time); cudaFree(m->d_bcs);
    delete m;
}

int mpm_bind(Mpm* m, int field, void* ptr) {
    if (field < 0 || field >= PIXIE_MPM_FIELD_COUNT) { m->error = "bad field id"; return 1; }
    if (mpm_sync(m, 0)) return 1;          // flush results into the arrays bound so far before one of them changes
    m->fields[field] = ptr;
    m->graph_valid = m->fused ? m->graph_valid : false;   // the fused graph only references the private copy
    return 0;
}
int mpm_set_params(Mpm* m, const pixie_mpm_params& p) {
    if (mpm_sync(m, 0)) return 1;
    if (p.n_grid != m->n_grid) {
        // set_parameters_dict re-allocates the grids when n_grid changes (mpm_solver_warp.py:318-343)
        if (m->slab) { m->error = "n_grid cannot change in slab mode"; return 1; }
        cudaDeviceSynchronize();
        if (!m->grid_borrowed) cudaFree(m->xbuf);
        cudaFree(m->grid_v);
        m->xbuf = nullptr; m->grid_borrowed = false;
        const size_t nodes = (size_t)p.n_grid * p.n_grid * p.n_grid;
        if (cudaMalloc(&m->xbuf, sizeof(SlabFlags) + nodes * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&m->grid_v, nodes * sizeof(float4)) != cudaSuccess) { m->error = "cudaMalloc failed"; return 1; }
        m->grid_mv = reinterpret_cast<float4*>(m->xbuf + sizeof(SlabFlags));
        cudaMemset(m->xbuf, 0, sizeof(SlabFlags) + nodes * sizeof(float4));
        cudaMemset(m->grid_v, 0, nodes * sizeof(float4));
        m->n_grid = p.n_grid;
        m->x_begin = 0; m->x_end = p.n_grid;
    }
    m->grid_lim = p.grid_lim;
    m->params = p;
    m->graph_valid = false;
    return 0;
}
int mpm_add_bc(Mpm* m, const pixie_mpm_bc& b) {
    if ((int)m->bcs.size() >= kMaxBC) { m->error = "too many boundary conditions (limit " + std::to_string(kMaxBC) + ")"; return 1; }
    if (b.kind >= PIXIE_BC_IMPULSE && !b.mask_dev) { m->error = "particle BC needs a mask"; return 1; }
    DevBC d{};
    d.kind = b.kind;
    for (int i = 0; i < 3; ++i) {
        d.point[i] = b.point[i]; d.normal[i] = b.normal[i]; d.size[i] = b.size[i]; d.velocity[i] = b.velocity[i];
        d.h1[i] = b.horizontal_axis_1[i]; d.h2[i] = b.horizontal_axis_2[i];
    }
    d.hhr[0] = b.half_height_and_radius[0]; d.hhr[1] = b.half_height_and_radius[1];
    d.start_time = b.start_time; d.end_time = b.end_time; d.friction = b.friction;
    d.surface_type = b.surface_type; d.reset = b.reset;
    d.rotation_scale = b.rotation_scale; d.translation_scale = b.translation_scale;
    d.mask = b.mask_dev;
    if (mpm_sync(m, 0)) return 1;
    m->bcs.push_back(d);
    // append in place: the device tables also hold the *moved* cuboid positions of earlier BCs
    const size_t k = m->bcs.size() - 1;
    cudaMemcpy(m->d_bcs + k, &d, sizeof(DevBC), cudaMemcpyHostToDevice);
    cudaMemcpy(m->pts + 3 * k, d.point, 3 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(m->pts + (size_t)kMaxBC * 3 + 3 * k, d.point, 3 * sizeof(float), cudaMemcpyHostToDevice);
    m->graph_valid = false;
    return 0;
}
int mpm_clear_bcs(Mpm* m) { m->bcs.clear(); m->graph_valid = false; return 0; }
int mpm_set_time(Mpm* m, double t) {
    const double both[2] = {t, t};
    if (cudaMemcpy(m->tslots, both, sizeof(both), cudaMemcpyHostToDevice) != cudaSuccess) return 1;
    return cudaMemcpy(m->d_time, &t, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess;
}
int mpm_get_time(Mpm* m, double* t) {
    const double* src = m->fused ? m->tslots + m->tpar : m->d_time;
    return cudaMemcpy(t, src, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess;
}

static int check_bound(Mpm* m) {
    static const int need[] = {PIXIE_MPM_X, PIXIE_MPM_V, PIXIE_MPM_F, PIXIE_MPM_F_TRIAL, PIXIE_MPM_C, PIXIE_MPM_STRESS,
                               PIXIE_MPM_VOL, PIXIE_MPM_MASS, PIXIE_MPM_MU, PIXIE_MPM_LAM, PIXIE_MPM_BULK,
                               PIXIE_MPM_YIELD, PIXIE_MPM_MATERIAL, PIXIE_MPM_SELECTION};
    for (int id : need)
        if (!m->fields[id]) { m->error = "field " + std::to_string(id) + " is not bound"; return 1; }
    if (m->params.update_cov_with_F && !m->fields[PIXIE_MPM_COV]) { m->error = "cov not bound"; return 1; }
    return 0;
}

static constexpr int kReorderEvery = 100;   // direct path: substeps between re-sorts of the order indirection

// (Re)builds Mpm::cell_order from the current positions. Stream-ordered, no host sync: the graph reads the same buffer.
static int mpm_build_cell_order(Mpm* m, cudaStream_t st) {
    const int n = m->n_active;
    if (n <= 0) { m->order_valid = false; return 0; }
    if (sort_alloc(m)) return 1;
    const float inv_dx = (float)((double)m->n_grid / (double)m->grid_lim);
    mpm_cell_key_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(m->fields[PIXIE_MPM_X]), n, inv_dx, m->n_grid, m->cell_keys, m->cell_idx);
    size_t bytes = m->cub_bytes;
    if (cub::DeviceRadixSort::SortPairs(m->cub_tmp, bytes, m->cell_keys, m->cell_keys_sorted, m->cell_idx, m->cell_order, n, 0, key_bits(m), st) != cudaSuccess) {
        m->error = "radix sort failed"; return 1;
    }
    m->steps_since_order = 0;
    if (!m->order_valid) { m->order_valid = true; m->graph_valid = false; }   // DevState::order changes from null to the buffer
    return 0;
}

// Threads per block of the per-particle kernels of the direct path
static int particle_block() {
    static const int b = getenv("PIXIE_MPM_BLOCK") ? atoi(getenv("PIXIE_MPM_BLOCK")) : 64;
    return (b == 32 || b == 64 || b == 128) ? b : 64;
}

static void launch_substep(const DevState& s, float dt, double dt_d, cudaStream_t st) {
    const int n = s.n;
    const size_t nodes = (size_t)(s.x_end - s.x_begin) * s.n_grid * s.n_grid;
    if (n > 0) {
        const int B = particle_block();
        mpm_stress_kernel<<<(n + B - 1) / B, B, 0, st>>>(s, dt);
        mpm_scatter_kernel<1><<<(n + B - 1) / B, B, 0, st>>>(s, dt);
    }
    mpm_grid_kernel<<<(unsigned)((nodes + 255) / 256), 256, 0, st>>>(s, dt);
    mpm_g2p_kernel<<<(std::max(n, 1) + particle_block() - 1) / particle_block(), particle_block(), 0, st>>>(s, dt, dt_d);
}

int mpm_step(Mpm* m, int n_substeps, double dt_d, cudaStream_t st) {
    if (check_bound(m)) return 1;
    if (n_substeps <= 0) return 0;
    if (m->fused) return mpm_step_fused(m, n_substeps, dt_d, st);
    const float dt = (float)dt_d;
    if ((!m->order_valid || m->steps_since_order >= kReorderEvery) && mpm_build_cell_order(m, st)) return 1;
    const DevState s = make_state(m);
    int done = 0;
    if (n_substeps >= kGraphSteps) {
        if (!m->graph_valid || m->graph_dt != dt_d) {
            if (m->graph) { cudaGraphExecDestroy(m->graph); m->graph = nullptr; }
            cudaStream_t cs;
            cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
            cudaGraph_t g = nullptr;
            bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                for (int i = 0; i < kGraphSteps; ++i) launch_substep(s, dt, dt_d, cs);
                ok = cudaStreamEndCapture(cs, &g) == cudaSuccess && g;
            }
            if (ok) ok = cudaGraphInstantiate(&m->graph, g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            if (!ok) { cudaGetLastError(); m->graph = nullptr; }
            m->graph_valid = ok;
            m->graph_dt = dt_d;
        }
        if (m->graph_valid) {
            while (n_substeps - done >= kGraphSteps) {
                if (m->steps_since_order >= kReorderEvery && mpm_build_cell_order(m, st)) return 1;
                if (cudaGraphLaunch(m->graph, st) != cudaSuccess) { m->error = "cudaGraphLaunch failed"; return 1; }
                done += kGraphSteps;
                m->launches += 4 * kGraphSteps;
                m->steps_since_order += kGraphSteps;
            }
        }
    }
    for (; done < n_substeps; ++done) { launch_substep(s, dt, dt_d, st); ++m->steps_since_order; m->launches += 4; }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { m->error = std::string("kernel launch failed: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

#define PIXIE_SIMPLE_LAUNCH(kernel)                                                         \
    if (mpm_sync(m, st)) return 1;                                                          \
    const DevState s = make_state(m);                                                       \
    kernel<<<(m->n + 255) / 256, 256, 0, st>>>(s);                                          \
    const cudaError_t e = cudaGetLastError();                                               \
    if (e != cudaSuccess) { m->error = cudaGetErrorString(e); return 1; }                   \
    return 0;

int mpm_compute_mu_lam(Mpm* m, cudaStream_t st) { PIXIE_SIMPLE_LAUNCH(mpm_mu_lam_kernel) }
int mpm_compute_bulk(Mpm* m, cudaStream_t st) { PIXIE_SIMPLE_LAUNCH(mpm_bulk_kernel) }
int mpm_compute_mass(Mpm* m, cudaStream_t st) { PIXIE_SIMPLE_LAUNCH(mpm_mass_kernel) }
int mpm_compute_cov_from_F(Mpm* m, cudaStream_t st) { PIXIE_SIMPLE_LAUNCH(mpm_cov_from_F_kernel) }
int mpm_compute_R_from_F(Mpm* m, cudaStream_t st) { PIXIE_SIMPLE_LAUNCH(mpm_R_from_F_kernel) }

int mpm_apply_additional_params(Mpm* m, const float* boxes_host, int n_boxes, cudaStream_t st) {
    if (n_boxes <= 0) return 0;
    if (mpm_sync(m, st)) return 1;
    float* d = nullptr;
    if (cudaMalloc(&d, (size_t)n_boxes * 10 * 4) != cudaSuccess) { m->error = "cudaMalloc failed"; return 1; }
    cudaMemcpyAsync(d, boxes_host, (size_t)n_boxes * 10 * 4, cudaMemcpyDefault, st);    // host or device source (UVA)
    const DevState s = make_state(m);
    mpm_additional_params_kernel<<<(m->n + 127) / 128, 128, 0, st>>>(s, d, n_boxes);
    cudaStreamSynchronize(st);
    cudaFree(d);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { m->error = cudaGetErrorString(e); return 1; }
    return 0;
}
int mpm_select_box(Mpm* m, const float* point, const float* size, int* mask, cudaStream_t st) {
    if (mpm_sync(m, st)) return 1;
    const DevState s = make_state(m);
    mpm_select_box_kernel<<<(m->n + 255) / 256, 256, 0, st>>>(s, make_float3(point[0], point[1], point[2]),
                                                              make_float3(size[0], size[1], size[2]), mask);
    return cudaGetLastError() != cudaSuccess;
}
int mpm_select_cylinder(Mpm* m, const float* point, const float* normal, float hh, float radius, int* mask, cudaStream_t st) {
    if (mpm_sync(m, st)) return 1;
    const DevState s = make_state(m);
    mpm_select_cyl_kernel<<<(m->n + 255) / 256, 256, 0, st>>>(s, make_float3(point[0], point[1], point[2]),
                                                              make_float3(normal[0], normal[1], normal[2]), hh, radius, mask);
    return cudaGetLastError() != cudaSuccess;
}
// ---- spatially sharded runs (BASELINE config 5): the caller owns the {mv,m} grid, exchanges ghost planes between
//      scatter and finish, and migrates particles by shrinking / growing the live prefix of the bound arrays.
int mpm_bind_grid(Mpm* m, void* mv4) {
    if (switch_to_direct(m)) return 1;
    m->grid_mv = reinterpret_cast<float4*>(mv4);            // the handle's own buffer (xbuf) stays allocated, unused
    m->grid_borrowed = true;
    m->graph_valid = false;
    return 0;
}
int mpm_set_slab(Mpm* m, int x_begin, int x_end) {
    if (x_begin < 0 || x_end > m->n_grid || x_begin >= x_end) { m->error = "bad slab range"; return 1; }
    if (switch_to_direct(m)) return 1;
    m->x_begin = x_begin; m->x_end = x_end;
    m->graph_valid = false;
    return 0;
}
int mpm_set_active_count(Mpm* m, int n_active) {
    if (n_active < 0 || n_active > m->n) { m->error = "active count exceeds the bound capacity"; return 1; }
    if (mpm_sync(m, 0)) return 1;          // results of the old live prefix go back first; the next step re-reads the arrays
    m->n_active = n_active;
    m->order_valid = false;       // the order lists exactly the live prefix
    m->graph_valid = false;
    return 0;
}
int mpm_substep_scatter(Mpm* m, double dt_d, cudaStream_t st) {
    if (switch_to_direct(m)) return 1;
    if (check_bound(m)) return 1;
    const DevState s = make_state(m);
    if (s.n > 0) {
        const int B = particle_block();
        mpm_stress_kernel<<<(s.n + B - 1) / B, B, 0, st>>>(s, (float)dt_d);
        mpm_scatter_kernel<1><<<(s.n + B - 1) / B, B, 0, st>>>(s, (float)dt_d);
    }
    return cudaGetLastError() != cudaSuccess;
}
int mpm_substep_finish(Mpm* m, double dt_d, cudaStream_t st) {
    if (switch_to_direct(m)) return 1;
    if (check_bound(m)) return 1;
    const DevState s = make_state(m);
    const size_t nodes = (size_t)(s.x_end - s.x_begin) * s.n_grid * s.n_grid;
    mpm_grid_kernel<<<(unsigned)((nodes + 255) / 256), 256, 0, st>>>(s, (float)dt_d);
    mpm_g2p_kernel<<<(std::max(s.n, 1) + particle_block() - 1) / particle_block(), particle_block(), 0, st>>>(s, (float)dt_d, dt_d);
    return cudaGetLastError() != cudaSuccess;
}
// ---- slab mode of the fused path (BASELINE config 5). The exchange buffer [SlabFlags][grid_mv] of each handle is made
//      visible to its x-neighbours (cudaIpc between processes, plain pointers inside one process); scatter, overlap
//      exchange and grid update then chain on the device with flag handshakes, no host in the loop.
int mpm_exchange_buffer(Mpm* m, void** base, size_t* bytes) {
    *base = m->xbuf;
    *bytes = sizeof(SlabFlags) + (size_t)m->n_grid * m->n_grid * m->n_grid * sizeof(float4);
    return 0;
}
int mpm_slab_attach(Mpm* m, int x0, int x1, int slack, const void* left_xbuf, const void* right_xbuf) {
    if (!m->fused) { m->error = "slab_attach needs the default (fused) path"; return 1; }
    if (x0 < 0 || x1 > m->n_grid || x0 >= x1 || slack < 0) { m->error = "bad slab range"; return 1; }
    if ((left_xbuf || right_xbuf) && (x1 - x0) < 2 + 2 * slack) { m->error = "slab narrower than 2 + 2*slack planes"; return 1; }
    if (mpm_sync(m, 0)) return 1;
    const int n = m->n_grid;
    m->slab = true; m->slab_x0 = x0; m->slab_x1 = x1; m->slab_slack = slack;
    m->peer_xbuf[0] = reinterpret_cast<const uint8_t*>(left_xbuf);
    m->peer_xbuf[1] = reinterpret_cast<const uint8_t*>(right_xbuf);
    // planes shared with a neighbour: both of them touch [x - slack, x + 2 + slack) around the interface x
    m->ov_lo[0] = std::max(0, x0 - slack); m->ov_hi[0] = std::min(n, x0 + 2 + slack);
    m->ov_lo[1] = std::max(0, x1 - slack); m->ov_hi[1] = std::min(n, x1 + 2 + slack);
    m->x_begin = left_xbuf ? m->ov_lo[0] : 0;
    m->x_end = right_xbuf ? m->ov_hi[1] : n;
    for (int sd = 0; sd < 2; ++sd) {
        cudaFree(m->ov_total[sd]); m->ov_total[sd] = nullptr;
        if (!m->peer_xbuf[sd]) continue;
        const size_t bytes = (size_t)(m->ov_hi[sd] - m->ov_lo[sd]) * n * n * sizeof(float4);
        if (cudaMalloc(&m->ov_total[sd], bytes) != cudaSuccess) { m->error = "cudaMalloc failed (overlap totals)"; return 1; }
        cudaMemset(m->ov_total[sd], 0, bytes);
    }
    cudaMemset(m->xbuf, 0, sizeof(SlabFlags));
    m->graph_valid = false;
    m->g2p_pending = false;
    return 0;
}
// One phase of a substep (single-process drivers sequence the phases of all slabs; a multi-process rank calls mpm_step).
int mpm_slab_phase(Mpm* m, int phase, double dt_d, cudaStream_t st) {
    if (!m->slab) { m->error = "not in slab mode"; return 1; }
    if (check_bound(m)) return 1;
    const float dt = (float)dt_d;
    if (phase == 0) {
        if (!m->internal_valid) { if (fused_gather_from_user(m, st)) return 1; m->g2p_pending = false; }
        else if (m->steps_since_sort >= kResortEvery && fused_resort(m, st)) return 1;
        fused_launch(m, m->g2p_pending, true, true, dt, st);
        m->g2p_pending = false;
    } else if (phase == 1) {
        halo_launch(m, st);
    } else if (phase == 2) {
        gridbox_launch(m, dt, dt_d, st);
        m->g2p_pending = true; m->slab_dt = dt;
        ++m->steps_since_sort;
        m->user_stale = true;
    } else { m->error = "bad phase"; return 1; }
    return cudaGetLastError() != cudaSuccess;
}
int mpm_slab_error(Mpm* m, int* flag) {
    SlabFlags f{};
    if (cudaMemcpy(&f, m->xbuf, sizeof(f), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    *flag = f.error;
    return 0;
}

int mpm_grid_ptrs(Mpm* m, float** mv4, float** v4) {
    *mv4 = reinterpret_cast<float*>(m->grid_mv);
    *v4 = reinterpret_cast<float*>(m->grid_v);
    return 0;
}
long long mpm_launch_count(Mpm* m) { return m->launches; }
const std::string& mpm_error(Mpm* m) { return m->error; }

}  // namespace pixie

// MPM default path: cell-sorted struct-of-arrays particle state and TWO launches per substep.
//
//   mpm_fused_kernel : g2p(i) -> x, F_trial -> particle BCs(i+1) -> return map + stress(i+1) -> p2g(i+1)
//                      one thread per particle; v, C, F_trial, stress stay in registers between the gather and the
//                      scatter (they are written out only by the launches that end a step() call);
//   mpm_gridbox_kernel : normalise + gravity + damping + grid BCs -> grid_v, clear-on-read of {mv, m}, restricted to the
//                      box of nodes the particles can touch; advances the clock / moving cuboids into the other
//                      parity slot.
//
// What the round-1 profile of the four-kernel path asked for: the substep was instruction-bound
// (~4.5 k thread-instructions per particle) on uncoalesced AoS loads.  Here
//   * the particle state is a PRIVATE cell-sorted SoA copy (component-major [comp][particle]): every load/store of a
//     warp is one or two full 128-byte lines, and consecutive lanes share stencil nodes;
//   * the 27-node loops are evaluated as separable sums (tensor-product B-spline): the gather reduces over z, then y,
//     then x (441 FMAs instead of ~800), the scatter builds node values from per-axis factors (~8 FMAs per node
//     instead of ~35).  Same arithmetic, different association: results agree with the reference order to fp32
//     rounding (tests/test_mpm_golden.py holds both to the reference-generated vectors);
//   * the scatter is warp-aggregated like before (runs of equal base cell, segmented shuffle, one red.global.add.v4.f32
//     per run and node), with the run length a template parameter.
// Included by mpm.cu inside its anonymous namespace (DevBC, M3/V3, the constitutive functions).
// Reference statements restated: mpm_utils.py:338-463 (p2g, g2p), 467-526 (stress), 583-588 (damping);
// mpm_solver_warp.py:528-547 (particle BCs), :785-974 (grid BCs), :899-905 + :637 (moving cuboid, clock).
#pragma once

constexpr int kFusedThreads = 32;

// component rows of the SoA buffer (floats)
enum : int { FS_X = 0, FS_V = 3, FS_C = 6, FS_F = 15, FS_FT = 24, FS_TAU = 33, FS_MASS = 42, FS_VOL = 43, FS_MU = 44, FS_LAM = 45,
             FS_BULK = 46, FS_YS = 47, FS_COV = 48, FS_NFLOAT = 54 };

constexpr int kInlinePBC = 4;       // particle BCs (impulses / velocity modifiers) carried in the kernel parameters

struct ParticleBC {                 // what apply_force / modify_particle_v_before_p2g read (mpm_solver_warp.py:1004-1179)
    int kind;
    float start_time, end_time;
    float velocity[3];              // force for impulses
    float point[3], normal[3], h1[3], h2[3];
    float rotation_scale, translation_scale;
    const int* mask;
};

struct FusedState {
    float* f;                       // [FS_NFLOAT][cap]
    int *material, *selection;      // [cap]
    const int* perm;                // [cap] index in the caller's arrays (BC masks are in the caller's order)
    int cap, n;
    const float4* grid_v;           // velocities of the step whose g2p runs here
    float4* grid_mv;                // scatter target {mv.xyz, m}
    int* box;                       // [6] node box lo.xyz, hi.xyz (exclusive) that the grid kernel sweeps; grown here if needed
    const double* time;             // clock of the substep whose stress / p2g run in this launch
    const DevBC* bcs;
    int n_bc, n_particle_bc;
    int n_pbc_inline;               // >= 0: the particle BCs are pbc[0..n); -1: more than kInlinePBC, walk the device table
    ParticleBC pbc[kInlinePBC];
    int n_grid;
    float dx, inv_dx;
    float rpic_damping, alpha, hardening, xi, plastic_viscosity, softening;
    int update_cov_with_F;
    int do_g2p, do_p2g, write_all;
    // slab-decomposed runs (one scene over several GPUs): substep counter that the exchange kernels key their flags on, and
    // the plane range a particle's stencil base may lie in (owned planes +- slack); outside it the scatter would reach planes
    // that are neither exchanged nor swept, so the kernel raises the error flag instead
    int* slab_step;                 // nullptr outside slab mode
    int* slab_err;
    int base_lo, base_hi;
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct AxisW { float w0, w1, w2, d0, d1, d2, fx; int b; };   // weights, derivative weights (without inv_dx), offset, base

__device__ __forceinline__ AxisW axis_weights(float g) {
    AxisW a;
    a.b = (int)(g - 0.5f);                               // wp.int truncates toward zero (mpm_utils.py:344-346)
    const float fx = g - (float)a.b;
    a.fx = fx;
    const float wa = 1.5f - fx, wb = fx - 1.0f, wc = fx - 0.5f;
    a.w0 = wa * wa * 0.5f;
    a.w1 = 0.f - wb * wb + 0.75f;
    a.w2 = wc * wc * 0.5f;
    a.d0 = fx - 1.5f;
    a.d1 = -2.0f * (fx - 1.0f);
    a.d2 = fx - 0.5f;
    return a;
}

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r = {x, y, z}; return r; }
__device__ __forceinline__ F3 fma3(float s, F3 a, F3 b) { return f3(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)); }
__device__ __forceinline__ F3 mul3(float s, F3 a) { return f3(s * a.x, s * a.y, s * a.z); }

// ---- g2p of one particle as separable sums.  Returns v = sum W v_g, B[c] = sum W (node_c - fx_c) v_g (grid units),
//      G[c] = sum dW_c v_g (without inv_dx).  CHECK = per-node bounds test (the reference indexes out of bounds there).
template <bool CHECK>
__device__ __forceinline__ void gather27(const float4* __restrict__ gv, int n, const AxisW& ax, const AxisW& ay, const AxisW& az,
                                         F3& v, F3& Bx, F3& By, F3& Bz, F3& Gx, F3& Gy, F3& Gz) {
    const float wx[3] = {ax.w0, ax.w1, ax.w2}, wy[3] = {ay.w0, ay.w1, ay.w2}, wz[3] = {az.w0, az.w1, az.w2};
    const float ex[3] = {ax.d0, ax.d1, ax.d2}, ey[3] = {ay.d0, ay.d1, ay.d2}, ez[3] = {az.d0, az.d1, az.d2};
    // w * (node - fx) per axis
    const float mx[3] = {wx[0] * (0.f - ax.fx), wx[1] * (1.f - ax.fx), wx[2] * (2.f - ax.fx)};
    const float my[3] = {wy[0] * (0.f - ay.fx), wy[1] * (1.f - ay.fx), wy[2] * (2.f - ay.fx)};
    const float mz[3] = {wz[0] * (0.f - az.fx), wz[1] * (1.f - az.fx), wz[2] * (2.f - az.fx)};
    const F3 Z = f3(0.f, 0.f, 0.f);
    v = Z; Bx = Z; By = Z; Bz = Z; Gx = Z; Gy = Z; Gz = Z;
    const long long base = ((long long)ax.b * n + ay.b) * n + az.b;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        F3 aa = Z, ab = Z, ac = Z, ba = Z, ca = Z;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            F3 a = Z, b = Z, c = Z;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                bool ok = true;
                if (CHECK) ok = (unsigned)ax.b + (unsigned)i < (unsigned)n && (unsigned)ay.b + (unsigned)j < (unsigned)n && (unsigned)az.b + (unsigned)k < (unsigned)n;
                if (ok) g = __ldg(gv + (base + ((long long)i * n + j) * n + k));
                const F3 gg = f3(g.x, g.y, g.z);
                a = fma3(wz[k], gg, a);
                b = fma3(mz[k], gg, b);
                c = fma3(ez[k], gg, c);
            }
            aa = fma3(wy[j], a, aa);
            ab = fma3(my[j], a, ab);
            ac = fma3(ey[j], a, ac);
            ba = fma3(wy[j], b, ba);
            ca = fma3(wy[j], c, ca);
        }
        v = fma3(wx[i], aa, v);
        Bx = fma3(mx[i], aa, Bx);
        Gx = fma3(ex[i], aa, Gx);
        By = fma3(wx[i], ab, By);
        Gy = fma3(wx[i], ac, Gy);
        Bz = fma3(wx[i], ba, Bz);
        Gz = fma3(wx[i], ca, Gz);
    }
}

// ---- compute_stress_from_F_trial (mpm_utils.py:467-526) for one particle; may update yield / mu / lam like the reference.
__device__ __noinline__ void plastic_return_and_stress(int material, const M3& Ft, float& mu, float& lam, float& ys, float bulk,
                                                        const FusedState& s, float dt, M3& F, M3& tau) {
    F = Ft;
    if (material == 1) F = return_von_mises(Ft, mu, lam, ys, s.hardening, s.xi, false, 0.f, mu, lam);
    else if (material == 2) F = return_sand(Ft, mu, lam, s.alpha);
    else if (material == 3) F = return_viscoplastic(Ft, mu, ys, s.plastic_viscosity, dt);
    else if (material == 5) F = return_von_mises(Ft, mu, lam, ys, s.hardening, s.xi, true, s.softening, mu, lam);
    const float J = m3_det(F);
    tau = m3_zero();
    if (material == 6) tau = stress_water(J, bulk);
    else if (material == 5) {
        M3 R;
        if (polar_rotation(F, R)) tau = stress_fcr_R(F, R, J, mu, lam);
        else { M3 U, V; V3 sig; svd3(F, U, sig, V); tau = stress_fcr(F, U, V, J, mu, lam); }
    } else if (material >= 1 && material <= 3) {
        M3 U, V; V3 sig;
        svd3(F, U, sig, V);
        tau = (material == 2) ? stress_drucker_prager(F, U, V, sig, mu, lam) : stress_stvk(F, U, V, sig, mu, lam);
    }
}

__device__ __noinline__ void fcr_svd_fallback(const M3& F, float J, float mu, float lam, M3& tau) {
    M3 U, V; V3 sig;
    svd3(F, U, sig, V);
    tau = stress_fcr(F, U, V, J, mu, lam);
}

// ---- pre-p2g particle operations: all impulses first, then all velocity modifiers (mpm_solver_warp.py:528-547)
template <class BC>
__device__ __forceinline__ void apply_impulse(const BC& bc, int orig, float time, float dt, float mass, float& vx, float& vy, float& vz, bool& dirty) {
    if (time >= bc.start_time && time < bc.end_time && bc.mask[orig] == 1) {
        vx = vx + (bc.velocity[0] / mass) * dt;          // apply_force :1015-1027 (force stored in velocity[])
        vy = vy + (bc.velocity[1] / mass) * dt;
        vz = vz + (bc.velocity[2] / mass) * dt;
        dirty = true;
    }
}
template <class BC>
__device__ __forceinline__ void apply_modifier(const BC& bc, int orig, float time, float px, float py, float pz, float& vx, float& vy, float& vz,
                                               bool& dirty) {
    if (!(time >= bc.start_time && time < bc.end_time) || bc.mask[orig] != 1) return;
    if (bc.kind == PIXIE_BC_VELOCITY_TRANSLATION) {
        vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2];
    } else {                                                                             // rotation :1137-1179
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
    }
    dirty = true;
}

__device__ __noinline__ bool particle_bcs(const FusedState& s, int orig, float time, float dt, float mass, float px, float py, float pz,
                                          float& vx, float& vy, float& vz) {
    bool dirty = false;
    if (s.n_pbc_inline >= 0) {                       // descriptors in the constant bank: no table loads
        for (int k = 0; k < s.n_pbc_inline; ++k)
            if (s.pbc[k].kind == PIXIE_BC_IMPULSE) apply_impulse(s.pbc[k], orig, time, dt, mass, vx, vy, vz, dirty);
        for (int k = 0; k < s.n_pbc_inline; ++k)
            if (s.pbc[k].kind != PIXIE_BC_IMPULSE) apply_modifier(s.pbc[k], orig, time, px, py, pz, vx, vy, vz, dirty);
        return dirty;
    }
    for (int k = 0; k < s.n_bc; ++k) {
        const DevBC& bc = s.bcs[k];
        if (bc.kind == PIXIE_BC_IMPULSE) apply_impulse(bc, orig, time, dt, mass, vx, vy, vz, dirty);
    }
    for (int k = 0; k < s.n_bc; ++k) {
        const DevBC& bc = s.bcs[k];
        if (bc.kind == PIXIE_BC_VELOCITY_TRANSLATION || bc.kind == PIXIE_BC_VELOCITY_ROTATION) apply_modifier(bc, orig, time, px, py, pz, vx, vy, vz, dirty);
    }
    return dirty;
}

// true when some particle BC's time window contains `time` (uniform over the grid: one test per thread, no mask loads)
__device__ __forceinline__ bool any_particle_bc_active(const FusedState& s, float time) {
    if (s.n_pbc_inline < 0) return true;
    bool any = false;
    for (int k = 0; k < s.n_pbc_inline; ++k) any = any || (time >= s.pbc[k].start_time && time < s.pbc[k].end_time);
    return any;
}

// AGG = log2 of the longest run of equal-cell lanes that is summed before one red is issued (0: no aggregation)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// HOIST: issue the late-needed per-particle loads at the top (costs registers across the gather); otherwise only prefetch
// their lines into L1 there and load at the point of use
template <int AGG, bool HOIST>
__global__ void __maxnreg__(88)
mpm_fused_kernel(const __grid_constant__ FusedState s, const float dt) {
    // programmatic dependent launch: let the next kernel of the chain get scheduled while this one runs, and wait here
    // for the previous one (its grid velocities / the particle state it wrote) — no-ops in a plain launch
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // slab mode: one scattering launch per substep advances the substep counter (nothing in THIS launch reads it; the grid
    // and grid kernels behind it in the stream do)
    if (s.slab_step && s.do_p2g && blockIdx.x == 0 && threadIdx.x == 0) *s.slab_step = *s.slab_step + 1;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = tid < s.n;
    const int p = live ? tid : max(s.n - 1, 0);               // whole warps stay converged for the shuffles
    const size_t cap = (size_t)s.cap;
    float* __restrict__ f = s.f;
    const bool act = live && s.selection[p] == 0;
    const int n = s.n_grid;

    // every load whose address is known up front is issued here, back to back: with ~5 warps per scheduler each dependent
    // trip to L2 that is taken alone costs the warp ~300 idle cycles (r02 ncu: 40 % of the stall samples were long-scoreboard)
    float px = f[(FS_X + 0) * cap + p], py = f[(FS_X + 1) * cap + p], pz = f[(FS_X + 2) * cap + p];
    float mass, vol, mu, lam, time = 0.f;
    int material, orig = 0;
    int box[6];
    auto late_loads = [&]() {
        mass = f[FS_MASS * cap + p]; vol = f[FS_VOL * cap + p];
        mu = f[FS_MU * cap + p]; lam = f[FS_LAM * cap + p];
        material = s.material[p];
        if (s.n_particle_bc > 0) { orig = s.perm[p]; time = (float)(*s.time); }
#pragma unroll
        for (int k = 0; k < 6; ++k) box[k] = s.box[k];
    };
    if (HOIST) late_loads();
    else if ((threadIdx.x & 31) == 0) {           // one lane per warp pulls the warp's lines (128 B = 32 particles) into L1
        prefetch_l1(f + FS_MASS * cap + p); prefetch_l1(f + FS_VOL * cap + p); prefetch_l1(f + FS_MU * cap + p);
        prefetch_l1(f + FS_LAM * cap + p); prefetch_l1(s.material + p); prefetch_l1(s.box);
        if (s.n_particle_bc > 0) prefetch_l1(s.perm + p);
    }
    float vx, vy, vz;
    M3 C, Ft;

    if (s.do_g2p) {
        // ------------------------------------------------------------------ g2p (mpm_utils.py:412-463)
        M3 F;
#pragma unroll
        for (int k = 0; k < 9; ++k) F.m[k] = f[(FS_F + k) * cap + p];
        const AxisW ax = axis_weights(px * s.inv_dx), ay = axis_weights(py * s.inv_dx), az = axis_weights(pz * s.inv_dx);
        F3 v, Bx, By, Bz, Gx, Gy, Gz;
        const bool inside = ax.b >= 0 && ay.b >= 0 && az.b >= 0 && ax.b < n - 2 && ay.b < n - 2 && az.b < n - 2;   // (no b + 2: a blown-up
                                                                                                                  // position converts to INT_MAX)
        if (inside) gather27<false>(s.grid_v, n, ax, ay, az, v, Bx, By, Bz, Gx, Gy, Gz);
        else gather27<true>(s.grid_v, n, ax, ay, az, v, Bx, By, Bz, Gx, Gy, Gz);
        vx = v.x; vy = v.y; vz = v.z;
        px = px + dt * vx; py = py + dt * vy; pz = pz + dt * vz;
        const float c4 = s.inv_dx * 4.0f;
        C.m[0] = Bx.x * c4; C.m[1] = By.x * c4; C.m[2] = Bz.x * c4;
        C.m[3] = Bx.y * c4; C.m[4] = By.y * c4; C.m[5] = Bz.y * c4;
        C.m[6] = Bx.z * c4; C.m[7] = By.z * c4; C.m[8] = Bz.z * c4;
        M3 G;     // grad v = sum v_g (x) grad W
        G.m[0] = Gx.x * s.inv_dx; G.m[1] = Gy.x * s.inv_dx; G.m[2] = Gz.x * s.inv_dx;
        G.m[3] = Gx.y * s.inv_dx; G.m[4] = Gy.y * s.inv_dx; G.m[5] = Gz.y * s.inv_dx;
        G.m[6] = Gx.z * s.inv_dx; G.m[7] = Gy.z * s.inv_dx; G.m[8] = Gz.z * s.inv_dx;
        M3 A = m3_ident();
#pragma unroll
        for (int k = 0; k < 9; ++k) A.m[k] += G.m[k] * dt;
        Ft = m3_mul(A, F);
        if (act) {
            f[(FS_X + 0) * cap + p] = px; f[(FS_X + 1) * cap + p] = py; f[(FS_X + 2) * cap + p] = pz;
            if (s.write_all || !s.do_p2g) {
                f[(FS_V + 0) * cap + p] = vx; f[(FS_V + 1) * cap + p] = vy; f[(FS_V + 2) * cap + p] = vz;
#pragma unroll
                for (int k = 0; k < 9; ++k) { f[(FS_C + k) * cap + p] = C.m[k]; f[(FS_FT + k) * cap + p] = Ft.m[k]; }
            }
            if (s.update_cov_with_F) {                         // update_cov :315-335
                float cv[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) cv[k] = f[(FS_COV + k) * cap + p];
                M3 cn;
                cn.m[0] = cv[0]; cn.m[1] = cv[1]; cn.m[2] = cv[2]; cn.m[3] = cv[1]; cn.m[4] = cv[3]; cn.m[5] = cv[4];
                cn.m[6] = cv[2]; cn.m[7] = cv[4]; cn.m[8] = cv[5];
                const M3 a = m3_mul(G, cn), b = m3_mul_t(cn, G);
                float c1[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) c1[k] = cn.m[k] + dt * (a.m[k] + b.m[k]);
                f[(FS_COV + 0) * cap + p] = c1[0]; f[(FS_COV + 1) * cap + p] = c1[1]; f[(FS_COV + 2) * cap + p] = c1[2];
                f[(FS_COV + 3) * cap + p] = c1[4]; f[(FS_COV + 4) * cap + p] = c1[5]; f[(FS_COV + 5) * cap + p] = c1[8];
            }
        }
        if (!s.do_p2g) return;
        if (!act) {          // particles outside the selection keep their stored v / C / F_trial
            vx = f[(FS_V + 0) * cap + p]; vy = f[(FS_V + 1) * cap + p]; vz = f[(FS_V + 2) * cap + p];
            px = f[(FS_X + 0) * cap + p]; py = f[(FS_X + 1) * cap + p]; pz = f[(FS_X + 2) * cap + p];
        }
    } else {
        vx = f[(FS_V + 0) * cap + p]; vy = f[(FS_V + 1) * cap + p]; vz = f[(FS_V + 2) * cap + p];
#pragma unroll
        for (int k = 0; k < 9; ++k) { C.m[k] = f[(FS_C + k) * cap + p]; Ft.m[k] = f[(FS_FT + k) * cap + p]; }
    }

    // ---------------------------------------------------------------------- particle BCs, stress (substep i+1)
    if (!HOIST) late_loads();
    if (s.n_particle_bc > 0 && any_particle_bc_active(s, time)) {
        const bool dirty = particle_bcs(s, orig, time, dt, mass, px, py, pz, vx, vy, vz);
        // the reference stores the modified v; only particles outside the selection keep it (g2p overwrites the rest)
        if (dirty && live && !act) { f[(FS_V + 0) * cap + p] = vx; f[(FS_V + 1) * cap + p] = vy; f[(FS_V + 2) * cap + p] = vz; }
    }
    M3 tau;
    if (act) {
        M3 F;
        if (material == 0) {
            // fixed-corotated stress needs only R = U V^T: Newton polar iteration, SVD only if it does not converge
            F = Ft;
            const float J = m3_det(F);
            M3 R;
            if (polar_rotation(F, R)) tau = stress_fcr_R(F, R, J, mu, lam);
            else fcr_svd_fallback(F, J, mu, lam, tau);
        } else if (material == 4 || material > 6 || material < 0) {
            F = Ft;
            tau = m3_zero();
        } else {
            float ys = f[FS_YS * cap + p];
            const float ys0 = ys, mu0 = mu;
            plastic_return_and_stress(material, Ft, mu, lam, ys, f[FS_BULK * cap + p], s, dt, F, tau);
            if (ys != ys0) f[FS_YS * cap + p] = ys;
            if (mu != mu0) { f[FS_MU * cap + p] = mu; f[FS_LAM * cap + p] = lam; }
        }
        // enforce symmetry (:524)
        {
            const float t01 = (tau.m[1] + tau.m[3]) / 2.0f, t02 = (tau.m[2] + tau.m[6]) / 2.0f, t12 = (tau.m[5] + tau.m[7]) / 2.0f;
            tau.m[0] = (tau.m[0] + tau.m[0]) / 2.0f; tau.m[4] = (tau.m[4] + tau.m[4]) / 2.0f; tau.m[8] = (tau.m[8] + tau.m[8]) / 2.0f;
            tau.m[1] = tau.m[3] = t01; tau.m[2] = tau.m[6] = t02; tau.m[5] = tau.m[7] = t12;
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) f[(FS_F + k) * cap + p] = F.m[k];
        if (s.write_all) {
#pragma unroll
            for (int k = 0; k < 9; ++k) f[(FS_TAU + k) * cap + p] = tau.m[k];
        }
    } else {
        tau = m3_zero();
    }

    // ---------------------------------------------------------------------- p2g (mpm_utils.py:338-394), substep i+1
    const AxisW ax = axis_weights(px * s.inv_dx), ay = axis_weights(py * s.inv_dx), az = axis_weights(pz * s.inv_dx);
    const bool inside = ax.b >= 0 && ay.b >= 0 && az.b >= 0 && ax.b < n - 2 && ay.b < n - 2 && az.b < n - 2;   // (no b + 2: a blown-up
                                                                                                                  // position converts to INT_MAX)
    {   // RPIC damping of C (:374-379)
        const float r = s.rpic_damping;
        if (r < -0.001f) C = m3_zero();
        else if (r != 0.f) {
            const M3 Ct = m3_t(C);
            M3 Cn;
#pragma unroll
            for (int k = 0; k < 9; ++k) Cn.m[k] = (1.0f - r) * C.m[k] + r / 2.0f * (C.m[k] - Ct.m[k]);
            C = Cn;
        }
    }
    if (s.slab_step && act && (ax.b < s.base_lo || ax.b >= s.base_hi)) atomicExch(s.slab_err, 2);   // drifted beyond the slack planes
    // keep the grid kernel's node box ahead of the particles (rare: the box has a margin and is rebuilt at every sort)
    if (act && inside) {
        if (ax.b < box[0]) atomicMin(s.box + 0, ax.b);
        if (ay.b < box[1]) atomicMin(s.box + 1, ay.b);
        if (az.b < box[2]) atomicMin(s.box + 2, az.b);
        if (ax.b + 3 > box[3]) atomicMax(s.box + 3, ax.b + 3);
        if (ay.b + 3 > box[4]) atomicMax(s.box + 4, ay.b + 3);
        if (az.b + 3 > box[5]) atomicMax(s.box + 5, az.b + 3);
    } else if (act) {
        atomicMin(s.box + 0, max(ax.b, 0)); atomicMin(s.box + 1, max(ay.b, 0)); atomicMin(s.box + 2, max(az.b, 0));
        atomicMax(s.box + 3, min(max(ax.b, 0), n - 3) + 3); atomicMax(s.box + 4, min(max(ay.b, 0), n - 3) + 3); atomicMax(s.box + 5, min(max(az.b, 0), n - 3) + 3);
    }

    // runs of equal base cell among consecutive lanes, chopped at 2^AGG lanes
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool contrib = act;
    const int key = (contrib && inside) ? (ax.b * n + ay.b) * n + az.b : -1 - lane;      // odd particles never share a run
    bool head = true;
    bool c1 = false, c2 = false, c4 = false;
    if (AGG > 0) {
        const int kprev = __shfl_up_sync(full, key, 1);
        head = (lane == 0) || (key != kprev);
        unsigned H = __ballot_sync(full, head);
        const int hl = 31 - __clz(H & (0xffffffffu >> (31 - lane)));      // head lane of my run
        head = head || (((lane - hl) & ((1 << AGG) - 1)) == 0);
        H = __ballot_sync(full, head);
        const unsigned above = H & ~((2u << lane) - 1u);                   // heads strictly above this lane
        const int seg_end = above ? (__ffs(above) - 2) : 31;               // last lane of my segment
        c1 = lane + 1 <= seg_end; c2 = lane + 2 <= seg_end; c4 = lane + 4 <= seg_end;
    }
    auto segsum = [&](float v) {
        if (AGG >= 1) { const float t = __shfl_down_sync(full, v, 1); if (c1) v += t; }
        if (AGG >= 2) { const float t = __shfl_down_sync(full, v, 2); if (c2) v += t; }
        if (AGG >= 3) { const float t = __shfl_down_sync(full, v, 4); if (c4) v += t; }
        return v;
    };

    // per-axis factors: node(i,j,k) = P_i wy_j wz_k + wx_i Q_j wz_k + wx_i wy_j R_k, mass = m wx_i wy_j wz_k
    const float m = contrib ? mass : 0.f;
    const float cf = contrib ? dt * vol * s.inv_dx : 0.f;                  // dt * vol * (inv_dx of grad W)
    const float wxa[3] = {ax.w0, ax.w1, ax.w2}, wya[3] = {ay.w0, ay.w1, ay.w2}, wza[3] = {az.w0, az.w1, az.w2};
    const float exa[3] = {ax.d0, ax.d1, ax.d2}, eya[3] = {ay.d0, ay.d1, ay.d2}, eza[3] = {az.d0, az.d1, az.d2};
    const F3 mv = f3(m * vx, m * vy, m * vz);
    const F3 mC0 = f3(m * C.m[0], m * C.m[3], m * C.m[6]), mC1 = f3(m * C.m[1], m * C.m[4], m * C.m[7]), mC2 = f3(m * C.m[2], m * C.m[5], m * C.m[8]);
    const F3 t0 = f3(cf * tau.m[0], cf * tau.m[3], cf * tau.m[6]), t1 = f3(cf * tau.m[1], cf * tau.m[4], cf * tau.m[7]),
             t2 = f3(cf * tau.m[2], cf * tau.m[5], cf * tau.m[8]);
    F3 P[3], Q[3], R[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float di = ((float)i - ax.fx) * s.dx, dj = ((float)i - ay.fx) * s.dx, dk = ((float)i - az.fx) * s.dx;
        P[i] = fma3(-exa[i], t0, mul3(wxa[i], fma3(di, mC0, mv)));
        Q[i] = fma3(-eya[i], t1, mul3(wya[i] * dj, mC1));
        R[i] = fma3(-eza[i], t2, mul3(wza[i] * dk, mC2));
    }
    float* const gbase = reinterpret_cast<float*>(s.grid_mv);
    const long long nbase = ((long long)ax.b * n + ay.b) * n + az.b;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const F3 S = fma3(wxa[i], Q[j], mul3(wya[j], P[i]));
            const float T = wxa[i] * wya[j];
            const float Tm = T * m;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const F3 val = fma3(T, R[k], mul3(wza[k], S));
                float a0 = val.x, a1 = val.y, a2 = val.z, a3 = Tm * wza[k];
                a0 = segsum(a0); a1 = segsum(a1); a2 = segsum(a2); a3 = segsum(a3);
                bool ok = head && contrib;
                if (!inside) ok = ok && (unsigned)ax.b + (unsigned)i < (unsigned)n && (unsigned)ay.b + (unsigned)j < (unsigned)n && (unsigned)az.b + (unsigned)k < (unsigned)n;
                ptx::red_add_v4_if(ok, gbase + 4 * (nbase + ((long long)i * n + j) * n + k), a0, a1, a2, a3);
            }
        }
}

// ------------------------------------------------------------------------------------------ grid update over the node box
// ---- slab exchange block: lives at the start of the handle's exchange buffer, in front of grid_mv, so that ONE
//      cudaIpc handle (or one pointer in single-process tests) gives a neighbour both the flags and the partial sums
struct SlabFlags {
    int scatter_done;     // substep whose scatter into this rank's {mv, m} grid (parity k & 1) is complete and visible
    int error;            // 1: a neighbour did not show up in time, 2: a particle drifted beyond the slack planes
    int step;             // this rank's substep counter
    int pad[61];
};
static_assert(sizeof(SlabFlags) == 256, "SlabFlags layout");

// Phase API only (a single-process driver sequences the phases of SEVERAL slabs on one stream and must raise every slab's
// scatter_done before the first grid sweep waits): one thread raises the flag after the particle kernel in front of it.
__global__ void mpm_publish_kernel(SlabFlags* mine) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        st_release_sys(&mine->scatter_done, mine->step);
    }
}

struct GridBoxArgs {
    float4* grid_mv;             // the {mv, m} grid the particle kernel of this substep scattered into
    float4* grid_v;
    // slab mode (else mine == nullptr). Each rank keeps TWO {mv, m} grids and alternates between them by substep parity:
    // on the planes [ov_lo, ov_hi) shared with neighbour s the sweep adds the neighbour's partial sums of the same parity,
    // read straight from its memory (NVLink) after ONE flag handshake, and leaves its own partial sums there untouched (the
    // neighbour reads them at the same time); they are cleared one substep later (grid_other), when the neighbour's
    // scatter_done of that substep proves its sweep of this one has finished.
    SlabFlags* mine;
    const SlabFlags* peer[2];
    const float4* peer_mv[2];    // the neighbours' grids of this substep's parity
    float4* grid_other;          // this rank's grid of the other parity
    int ov_lo[2], ov_hi[2];
    int publish_scatter;         // 1: raise scatter_done here (chained runs); 0: a publish launch did (phase API)
    const int* box;              // lo.xyz, hi.xyz
    const double* time_in; double* time_out;
    const float* pts_in; float* pts_out;       // [n_bc][3] collider points, by parity (the cuboid ones move)
    const DevBC* bcs; int n_bc;
    int n_grid, x_begin, x_end;
    float dx, gx, gy, gz, grid_v_damping_scale;
};

__global__ void __launch_bounds__(256)
mpm_gridbox_kernel(const GridBoxArgs s, const float dt, const double dt_d) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (s.mine) {
        // the particle kernel in front of this launch has completed (griddepcontrol.wait / stream order): tell the neighbours
        // that this rank's partial sums of substep k are in place BEFORE waiting for theirs (release at system scope: they
        // read them over NVLink). The particle kernel stays free of fences: its reds remain fire-and-forget.
        const int k = s.mine->step;
        if (s.publish_scatter && blockIdx.x == 0 && threadIdx.x == 0) {
            __threadfence_system();
            st_release_sys(&s.mine->scatter_done, k);
        }
        // both neighbours are awaited at the same time (two polling threads per block, bounded)
        __shared__ int ok_s[2];
        if (threadIdx.x < 64 && (threadIdx.x & 31) == 0) {
            const int side = threadIdx.x >> 5;
            int ok = 1;
            if (s.peer[side]) {
                ok = 0;
                for (long long it = 0; it < (1ll << 22); ++it) {
                    if (ld_acquire_sys(&s.peer[side]->scatter_done) >= k) { ok = 1; break; }
                    __nanosleep(100);
                }
                if (!ok) atomicExch(&s.mine->error, 1);
            }
            ok_s[side] = ok;
        }
        __syncthreads();
        if (!(ok_s[0] && ok_s[1])) return;
    }
    const int n = s.n_grid;
    const int lx = max(s.box[0], s.x_begin), ly = s.box[1], lz = s.box[2];
    const int hx = min(s.box[3], s.x_end), hy = s.box[4], hz = s.box[5];
    const int ex = hx - lx, ey = hy - ly, ez = hz - lz;
    const float time = (float)(*s.time_in);
    // ---- substep epilogue: clock and moving cuboids into the other parity slot (nothing in this launch reads it)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double t = *s.time_in;
        for (int k = 0; k < s.n_bc; ++k) {
            const DevBC& bc = s.bcs[k];
            float q0 = s.pts_in[3 * k], q1 = s.pts_in[3 * k + 1], q2 = s.pts_in[3 * k + 2];
            if (bc.kind == PIXIE_BC_CUBOID && t >= (double)bc.start_time && t < (double)bc.end_time) {
                // modify(): Python-float arithmetic, stored back as fp32 (mpm_solver_warp.py:899-905)
                q0 = (float)((double)q0 + dt_d * (double)bc.velocity[0]);
                q1 = (float)((double)q1 + dt_d * (double)bc.velocity[1]);
                q2 = (float)((double)q2 + dt_d * (double)bc.velocity[2]);
            }
            s.pts_out[3 * k] = q0; s.pts_out[3 * k + 1] = q1; s.pts_out[3 * k + 2] = q2;
        }
        *s.time_out = t + dt_d;
    }
    if (ex <= 0 || ey <= 0 || ez <= 0) return;
    const long long total = (long long)ex * ey * ez;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(t % ez), iy = (int)((t / ez) % ey), ix = (int)(t / ((long long)ez * ey));
        const int gx = lx + ix, gy = ly + iy, gz = lz + iz;
        const size_t idx = ((size_t)gx * n + gy) * n + gz;
        const float4 own = s.grid_mv[idx];
        float4 mv = own;
        bool shared_plane = false;
        if (s.mine) {
#pragma unroll
            for (int side = 0; side < 2; ++side)
                if (s.peer[side] && gx >= s.ov_lo[side] && gx < s.ov_hi[side]) {
                    const float4 oth = s.peer_mv[side][idx];       // neighbour's memory; own + oth == oth + own on its side
                    mv = make_float4(own.x + oth.x, own.y + oth.y, own.z + oth.z, own.w + oth.w);
                    shared_plane = true;
                }
        }
        float vx = 0.f, vy = 0.f, vz = 0.f;
        if (mv.w > 1e-15f) {                                   // grid_normalization_and_gravity :398-409
            const float inv = 1.0f / mv.w;
            vx = mv.x * inv + dt * s.gx; vy = mv.y * inv + dt * s.gy; vz = mv.z * inv + dt * s.gz;
        }
        if (s.grid_v_damping_scale < 1.0f) {                   // add_damping_via_grid :583-588 (only if < 1)
            vx *= s.grid_v_damping_scale; vy *= s.grid_v_damping_scale; vz *= s.grid_v_damping_scale;
        }
        for (int k = 0; k < s.n_bc; ++k) {
            const DevBC& bc = s.bcs[k];
            if (bc.kind > PIXIE_BC_BOUNDING_BOX) continue;
            const bool active = time >= bc.start_time && time < bc.end_time;
            if (bc.kind == PIXIE_BC_SURFACE_COLLIDER) {        // :785-840
                if (active) {
                    const float ox = (float)gx * s.dx - s.pts_in[3 * k], oy = (float)gy * s.dx - s.pts_in[3 * k + 1], oz = (float)gz * s.dx - s.pts_in[3 * k + 2];
                    if (ox * bc.normal[0] + oy * bc.normal[1] + oz * bc.normal[2] < 0.0f) {
                        if (bc.surface_type == 11) {
                            const float zz = (float)gz * s.dx;
                            if (zz < 0.4f || zz > 0.53f) { vx = 0.f; vy = 0.f; vz = 0.f; }
                            else { vx = vx * 0.3f; vy = 0.0f * 0.3f; vz = vz * 0.3f; }
                        } else {
                            // sticky -> 0; slip / separate: the reference computes the projected velocity and then
                            // overwrites the node with zero (:838-840)
                            vx = 0.f; vy = 0.f; vz = 0.f;
                        }
                    }
                }
            } else if (bc.kind == PIXIE_BC_CUBOID) {           // :874-897
                if (active) {
                    const float ox = (float)gx * s.dx - s.pts_in[3 * k], oy = (float)gy * s.dx - s.pts_in[3 * k + 1], oz = (float)gz * s.dx - s.pts_in[3 * k + 2];
                    if (fabsf(ox) < bc.size[0] && fabsf(oy) < bc.size[1] && fabsf(oz) < bc.size[2]) {
                        vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2];
                    }
                } else if (bc.reset == 1) {
                    if (time < bc.end_time + 15.0f * dt) { vx = 0.f; vy = 0.f; vz = 0.f; }
                }
            } else {                                           // bounding box :917-974
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
        if (shared_plane) s.grid_other[idx] = make_float4(0.f, 0.f, 0.f, 0.f);      // last substep's sums: the neighbour is done with them
        else if (own.x != 0.f || own.y != 0.f || own.z != 0.f || own.w != 0.f) s.grid_mv[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------ sorted-state maintenance
// base-cell key of a position (clamped), for the radix sort
__global__ void fs_key_kernel(const float* __restrict__ x, long long stride_comp, long long stride_part, int n, float inv_dx, int n_grid,
                              int* __restrict__ keys, int* __restrict__ idx) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const AxisW ax = axis_weights(x[0 * stride_comp + p * stride_part] * inv_dx), ay = axis_weights(x[1 * stride_comp + p * stride_part] * inv_dx),
                az = axis_weights(x[2 * stride_comp + p * stride_part] * inv_dx);
    const int bx = min(max(ax.b, 0), n_grid - 1), by = min(max(ay.b, 0), n_grid - 1), bz = min(max(az.b, 0), n_grid - 1);
    keys[p] = (bx * n_grid + by) * n_grid + bz;
    idx[p] = p;
}

// node box of all particles (+ margin), from positions in either layout
__global__ void fs_box_kernel(const float* __restrict__ x, long long stride_comp, long long stride_part, int n, float inv_dx, int n_grid,
                              int margin, int* __restrict__ box, int finalize) {
    if (finalize) {      // second launch: apply the margin and clamp
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            for (int a = 0; a < 3; ++a) {
                box[a] = max(box[a] - margin, 0);
                box[3 + a] = min(box[3 + a] + margin, n_grid);
            }
        }
        return;
    }
    int lo[3] = {n_grid, n_grid, n_grid}, hi[3] = {0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const AxisW w = axis_weights(x[a * stride_comp + p * stride_part] * inv_dx);
            lo[a] = min(lo[a], max(w.b, 0));
            hi[a] = max(hi[a], min(max(w.b, 0), n_grid - 3) + 3);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(box + a, lo[a]); atomicMax(box + 3 + a, hi[a]); }
    }
}

// slab runs: how many planes the farthest particle's stencil base lies outside [lo, hi) (0 = all inside); `out` is max-ed into
__global__ void fs_excursion_kernel(const float* __restrict__ x, long long stride_part, int n, float inv_dx, int lo, int hi,
                                    int* __restrict__ out) {
    int e = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int b = axis_weights(x[p * stride_part] * inv_dx).b;
        e = max(e, max(lo - b, b - (hi - 1)));
    }
    for (int o = 16; o > 0; o >>= 1) e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
    if ((threadIdx.x & 31) == 0 && e > 0) atomicMax(out, e);
}

struct FsUser {      // the caller's arrays (array-of-structs, original order)
    float *x, *v, *C, *F, *Ft, *stress, *mass, *vol, *mu, *lam, *bulk, *ys, *cov;
    int *material, *selection;
};

// caller's arrays -> sorted SoA (new slot q takes particle order[q])
__global__ void fs_gather_kernel(const FsUser u, const int* __restrict__ order, int n, int cap, float* __restrict__ f,
                                 int* __restrict__ material, int* __restrict__ selection, int* __restrict__ perm, int with_cov) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int o = order[q];
    const size_t c = (size_t)cap;
#pragma unroll
    for (int k = 0; k < 3; ++k) { f[(FS_X + k) * c + q] = u.x[3 * (size_t)o + k]; f[(FS_V + k) * c + q] = u.v[3 * (size_t)o + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        f[(FS_C + k) * c + q] = u.C[9 * (size_t)o + k]; f[(FS_F + k) * c + q] = u.F[9 * (size_t)o + k];
        f[(FS_FT + k) * c + q] = u.Ft[9 * (size_t)o + k]; f[(FS_TAU + k) * c + q] = u.stress[9 * (size_t)o + k];
    }
    f[FS_MASS * c + q] = u.mass[o]; f[FS_VOL * c + q] = u.vol[o]; f[FS_MU * c + q] = u.mu[o]; f[FS_LAM * c + q] = u.lam[o];
    f[FS_BULK * c + q] = u.bulk[o]; f[FS_YS * c + q] = u.ys[o];
    if (with_cov) {
#pragma unroll
        for (int k = 0; k < 6; ++k) f[(FS_COV + k) * c + q] = u.cov[6 * (size_t)o + k];
    }
    material[q] = u.material[o]; selection[q] = u.selection[o]; perm[q] = o;
}

// sorted SoA -> caller's arrays (everything a substep writes)
__global__ void fs_unsort_kernel(const FsUser u, const int* __restrict__ perm, int n, int cap, const float* __restrict__ f, int with_cov) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int o = perm[q];
    const size_t c = (size_t)cap;
#pragma unroll
    for (int k = 0; k < 3; ++k) { u.x[3 * (size_t)o + k] = f[(FS_X + k) * c + q]; u.v[3 * (size_t)o + k] = f[(FS_V + k) * c + q]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        u.C[9 * (size_t)o + k] = f[(FS_C + k) * c + q]; u.F[9 * (size_t)o + k] = f[(FS_F + k) * c + q];
        u.Ft[9 * (size_t)o + k] = f[(FS_FT + k) * c + q]; u.stress[9 * (size_t)o + k] = f[(FS_TAU + k) * c + q];
    }
    u.mu[o] = f[FS_MU * c + q]; u.lam[o] = f[FS_LAM * c + q]; u.ys[o] = f[FS_YS * c + q];
    if (with_cov) {
#pragma unroll
        for (int k = 0; k < 6; ++k) u.cov[6 * (size_t)o + k] = f[(FS_COV + k) * c + q];
    }
}

// re-sort of the live sorted state: slot q of the destination takes slot order[q] of the source
__global__ void fs_permute_kernel(const float* __restrict__ fs, const int* __restrict__ ms, const int* __restrict__ ss, const int* __restrict__ ps,
                                  const int* __restrict__ order, int n, int cap, float* __restrict__ fd, int* __restrict__ md, int* __restrict__ sd,
                                  int* __restrict__ pd) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int o = order[q];
    const size_t c = (size_t)cap;
#pragma unroll 6
    for (int k = 0; k < FS_NFLOAT; ++k) fd[k * c + q] = fs[k * c + o];
    md[q] = ms[o]; sd[q] = ss[o]; pd[q] = ps[o];
}

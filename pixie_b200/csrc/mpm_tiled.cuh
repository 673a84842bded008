// MPM "tiled" path: cell-sorted particles, one CTA per occupied tile of 4x4x4 cells with its 6x6x6
// node halo in shared memory, and g2p(i) -> clock -> particle BCs(i+1) -> stress(i+1) -> p2g(i+1)
// fused into ONE launch per substep.
//
// Why: the three-kernel path is bound by L2 atomics (27 vector reds per particle = 10.8 M fp32
// atomics per substep at 100k particles, ncu: p2g 39 us, g2p 19 us). Here a tile's particles
// accumulate into the shared-memory halo (lane-rotated node order keeps same-cell lanes on different
// nodes) and one `red.global.add.v4.f32` per touched halo node flushes the tile: ~50x fewer global
// atomics, the 27 gathers of g2p hit shared memory, and x / F are the only per-particle arrays that
// cross HBM/L2 in the steady state (v, C, F_trial live in registers between g2p and p2g).
//
// Included by mpm.cu inside its anonymous namespace (re-uses DevBC, M3/V3, bspline, the constitutive
// functions).  Reference statements restated here: mpm_utils.py:338-463, 467-526, 583-588;
// mpm_solver_warp.py:528-547 (particle BCs), :785-974 (grid BCs), :899-905 + :637 (clock, moving cuboid).
#pragma once

constexpr int kTile = 4;                 // cells per tile edge
constexpr int kHalo = kTile + 2;         // nodes per halo edge (quadratic B-spline: base .. base+2)
constexpr int kHaloNodes = kHalo * kHalo * kHalo;
constexpr int kTiledThreads = 128;

struct TiledState {
    // sorted particle arrays (internal order); perm[p] = index in the caller's arrays
    float *x, *v, *C, *F, *Ft, *stress;
    float *mass, *vol, *mu, *lam, *bulk, *yield_stress;
    int *material, *selection;
    const int* perm;
    // tiles
    const int* occ;            // [n_occ] tile ids
    const int* tile_off;       // [ntiles + 1] particle offsets per tile
    int nt;                    // tiles per axis
    // grid: three rotating {mv.xyz, m} buffers
    const float4* mv_read;     // complete scatter of the step whose g2p runs now
    float4* mv_write;          // scatter target of the next step (zero on entry)
    float4* mv_clear;          // buffer the previous launch read; cleared here for the launch after next
    // clock + moving BC points, double buffered by substep parity
    const double* time_in; double* time_out;
    const float* pts_in; float* pts_out;      // [n_bc][3]
    const DevBC* bcs; int n_bc;
    int n, n_grid;
    float dx, inv_dx, gx, gy, gz;
    float rpic_damping, grid_v_damping_scale, alpha, hardening, xi, plastic_viscosity, softening;
    int do_g2p, do_p2g, write_all;            // write_all: also store v, C, F_trial, stress (last launches of a step())
};

// grid_normalization_and_gravity + add_damping_via_grid + every grid BC, for ONE node.
__device__ __forceinline__ float4 node_velocity(const TiledState& s, float4 mv, int gx, int gy, int gz, float time, float dt) {
    const int n = s.n_grid;
    float vx = 0.f, vy = 0.f, vz = 0.f;
    if (mv.w > 1e-15f) {
        const float inv = 1.0f / mv.w;
        vx = mv.x * inv + dt * s.gx; vy = mv.y * inv + dt * s.gy; vz = mv.z * inv + dt * s.gz;
    }
    if (s.grid_v_damping_scale < 1.0f) { vx *= s.grid_v_damping_scale; vy *= s.grid_v_damping_scale; vz *= s.grid_v_damping_scale; }
    for (int k = 0; k < s.n_bc; ++k) {
        const DevBC& bc = s.bcs[k];
        if (bc.kind > PIXIE_BC_BOUNDING_BOX) continue;
        const bool active = time >= bc.start_time && time < bc.end_time;
        const float px = s.pts_in[3 * k], py = s.pts_in[3 * k + 1], pz = s.pts_in[3 * k + 2];
        if (bc.kind == PIXIE_BC_SURFACE_COLLIDER) {
            if (active) {
                const float ox = (float)gx * s.dx - px, oy = (float)gy * s.dx - py, oz = (float)gz * s.dx - pz;
                if (ox * bc.normal[0] + oy * bc.normal[1] + oz * bc.normal[2] < 0.0f) {
                    if (bc.surface_type == 11) {
                        const float zz = (float)gz * s.dx;
                        if (zz < 0.4f || zz > 0.53f) { vx = 0.f; vy = 0.f; vz = 0.f; }
                        else { vx = vx * 0.3f; vy = 0.0f * 0.3f; vz = vz * 0.3f; }
                    } else { vx = 0.f; vy = 0.f; vz = 0.f; }
                }
            }
        } else if (bc.kind == PIXIE_BC_CUBOID) {
            if (active) {
                const float ox = (float)gx * s.dx - px, oy = (float)gy * s.dx - py, oz = (float)gz * s.dx - pz;
                if (fabsf(ox) < bc.size[0] && fabsf(oy) < bc.size[1] && fabsf(oz) < bc.size[2]) {
                    vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2];
                }
            } else if (bc.reset == 1) {
                if (time < bc.end_time + 15.0f * dt) { vx = 0.f; vy = 0.f; vz = 0.f; }
            }
        } else {   // bounding box
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
    return make_float4(vx, vy, vz, 0.f);
}

// shared-memory float2 accumulate (ATOMS has no fp32 add on sm_100: one 64-bit CAS per two components)
__device__ __forceinline__ void smem_add2(float* addr, float a, float b) {
    unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *p, assumed;
    do {
        assumed = old;
        float2 f = *reinterpret_cast<float2*>(&assumed);
        f.x += a; f.y += b;
        old = atomicCAS(p, assumed, *reinterpret_cast<unsigned long long*>(&f));
    } while (old != assumed);
}

__device__ __forceinline__ float sel3(int i, float a, float b, float c) { return i == 0 ? a : (i == 1 ? b : c); }

__global__ void __launch_bounds__(kTiledThreads, 4)
mpm_tiled_kernel(const TiledState s, const float dt, const double dt_d) {
    __shared__ __align__(16) float4 sv[kHaloNodes];      // node velocities of this tile's halo (g2p source)
    // {mv.xyz, m} accumulators (p2g target): one private copy per warp, so no shared-memory atomics are needed
    // (fp32 smem atomics are ATOMS.CAS loops on sm_100: measured 7x slower than the global-red path they replaced);
    // lanes of a warp that hit the same node in the same step are serialised with match.any
    __shared__ __align__(16) float4 wacc[kTiledThreads / 32][kHaloNodes];
    const int tid = threadIdx.x;
    const int n = s.n_grid;
    const size_t nodes = (size_t)n * n * n;

    // ---- clear the buffer the launch after next will scatter into
    for (size_t i = (size_t)blockIdx.x * blockDim.x + tid; i < nodes; i += (size_t)gridDim.x * blockDim.x)
        s.mv_clear[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int tile = s.occ[blockIdx.x];
    const int tz = tile % s.nt, ty = (tile / s.nt) % s.nt, tx = tile / (s.nt * s.nt);
    const int bx0 = tx * kTile, by0 = ty * kTile, bz0 = tz * kTile;     // first base cell = first halo node
    const double t_d = *s.time_in;
    const float time = (float)t_d;
    const float time1 = s.do_g2p ? (float)(t_d + dt_d) : time;          // clock seen by the fused p2g of the NEXT step

    for (int l = tid; l < kHaloNodes; l += blockDim.x) {
        const int i = l / (kHalo * kHalo), j = (l / kHalo) % kHalo, k = l % kHalo;
        const int gx = bx0 + i, gy = by0 + j, gz = bz0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s.do_g2p && gx < n && gy < n && gz < n) v = node_velocity(s, s.mv_read[((size_t)gx * n + gy) * n + gz], gx, gy, gz, time, dt);
        sv[l] = v;
    }
    for (int l = tid; l < (kTiledThreads / 32) * kHaloNodes; l += blockDim.x) (&wacc[0][0])[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    const int p_begin = s.tile_off[tile], p_end = s.tile_off[tile + 1];
    float4* my_acc = wacc[tid >> 5];
    const unsigned lane = tid & 31, lt_mask = (1u << lane) - 1u;
    for (int pb = p_begin; pb < p_end; pb += blockDim.x) {       // warp-uniform trip count (the scatter below is warp-collective)
      const int p = pb + tid;
      const bool active = p < p_end;
      bool scatter = false;
      float vx = 0.f, vy = 0.f, vz = 0.f, px = 0.f, py = 0.f, pz = 0.f, mass = 0.f, vol = 0.f;
      M3 C = m3_zero(), tau = m3_zero();
      if (active) do {
        px = s.x[3 * p]; py = s.x[3 * p + 1]; pz = s.x[3 * p + 2];
        const bool simulated = s.selection[p] == 0;
        M3 Ft;
        if (s.do_g2p && simulated) {
            // ---- g2p (mpm_utils.py:412-463) from the shared halo; particles that drifted out of their tile
            //      since the last sort read the global grid instead
            const Weights W = bspline_t(s.inv_dx, px, py, pz);
            const int lx = W.bx - bx0, ly = W.by - by0, lz = W.bz - bz0;
            const bool in_tile = (unsigned)lx < (unsigned)kTile && (unsigned)ly < (unsigned)kTile && (unsigned)lz < (unsigned)kTile;
            float nvx = 0.f, nvy = 0.f, nvz = 0.f;
            M3 nC = m3_zero(), nF = m3_zero();
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float4 gv;
                        if (in_tile) gv = sv[((lx + i) * kHalo + (ly + j)) * kHalo + (lz + k)];
                        else {
                            const int ix = W.bx + i, iy = W.by + j, iz = W.bz + k;
                            gv = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((unsigned)ix < (unsigned)n && (unsigned)iy < (unsigned)n && (unsigned)iz < (unsigned)n)
                                gv = node_velocity(s, s.mv_read[((size_t)ix * n + iy) * n + iz], ix, iy, iz, time, dt);
                        }
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
            vx = nvx; vy = nvy; vz = nvz;
            px = px + dt * nvx; py = py + dt * nvy; pz = pz + dt * nvz;
            C = nC;
            M3 A = m3_ident();
#pragma unroll
            for (int i = 0; i < 9; ++i) A.m[i] += nF.m[i] * dt;
            Ft = m3_mul(A, load_m3(s.F, p));
            s.x[3 * p] = px; s.x[3 * p + 1] = py; s.x[3 * p + 2] = pz;
            if (!s.do_p2g || s.write_all) {
                s.v[3 * p] = vx; s.v[3 * p + 1] = vy; s.v[3 * p + 2] = vz;
                store_m3(s.C, p, C);
                store_m3(s.Ft, p, Ft);
            }
        } else {
            vx = s.v[3 * p]; vy = s.v[3 * p + 1]; vz = s.v[3 * p + 2];
            C = load_m3(s.C, p);
            Ft = load_m3(s.Ft, p);
        }
        if (!s.do_p2g) break;

        // ---- pre-p2g particle operations at the clock of the step being scattered (mpm_solver_warp.py:528-547)
        mass = s.mass[p];
        bool v_dirty = false;
        for (int k = 0; k < s.n_bc; ++k) {
            const DevBC& bc = s.bcs[k];
            if (bc.kind != PIXIE_BC_IMPULSE) continue;
            if (time1 >= bc.start_time && time1 < bc.end_time && bc.mask[s.perm[p]] == 1) {
                vx = vx + (bc.velocity[0] / mass) * dt; vy = vy + (bc.velocity[1] / mass) * dt; vz = vz + (bc.velocity[2] / mass) * dt;
                v_dirty = true;
            }
        }
        for (int k = 0; k < s.n_bc; ++k) {
            const DevBC& bc = s.bcs[k];
            if (bc.kind == PIXIE_BC_VELOCITY_TRANSLATION) {
                if (time1 >= bc.start_time && time1 < bc.end_time && bc.mask[s.perm[p]] == 1) {
                    vx = bc.velocity[0]; vy = bc.velocity[1]; vz = bc.velocity[2]; v_dirty = true;
                }
            } else if (bc.kind == PIXIE_BC_VELOCITY_ROTATION) {
                if (time1 >= bc.start_time && time1 < bc.end_time && bc.mask[s.perm[p]] == 1) {
                    const float ox = px - bc.point[0], oy = py - bc.point[1], oz = pz - bc.point[2];
                    const float on = ox * bc.normal[0] + oy * bc.normal[1] + oz * bc.normal[2];
                    const float hx = ox - on * bc.normal[0], hy = oy - on * bc.normal[1], hz = oz - on * bc.normal[2];
                    const float hd = sqrtf(hx * hx + hy * hy + hz * hz);
                    float theta = acosf((ox * bc.h1[0] + oy * bc.h1[1] + oz * bc.h1[2]) / hd);
                    if (!(ox * bc.h2[0] + oy * bc.h2[1] + oz * bc.h2[2] > 0.f)) theta = -theta;
                    const float a1 = -hd * sinf(theta) * bc.rotation_scale, a2 = hd * cosf(theta) * bc.rotation_scale;
                    const float av = bc.translation_scale;
                    vx = a1 * bc.h1[0] + a2 * bc.h2[0] + av * bc.normal[0];
                    vy = a1 * bc.h1[1] + a2 * bc.h2[1] + av * bc.normal[1];
                    vz = a1 * bc.h1[2] + a2 * bc.h2[2] + av * bc.normal[2];
                    v_dirty = true;
                }
            }
        }
        if (v_dirty && (!s.do_g2p || s.write_all)) { s.v[3 * p] = vx; s.v[3 * p + 1] = vy; s.v[3 * p + 2] = vz; }
        if (!simulated) break;

        // ---- compute_stress_from_F_trial (mpm_utils.py:467-526)
        const int material = s.material[p];
        float mu = s.mu[p], lam = s.lam[p];
        M3 F = Ft;
        if (material == 1) {
            float ys = s.yield_stress[p]; const float ys0 = ys;
            F = return_von_mises(Ft, mu, lam, ys, s.hardening, s.xi, false, 0.f, mu, lam);
            if (ys != ys0) s.yield_stress[p] = ys;
        } else if (material == 2) {
            F = return_sand(Ft, mu, lam, s.alpha);
        } else if (material == 3) {
            F = return_viscoplastic(Ft, mu, s.yield_stress[p], s.plastic_viscosity, dt);
        } else if (material == 5) {
            float ys = s.yield_stress[p]; const float ys0 = ys, mu0 = mu;
            F = return_von_mises(Ft, mu, lam, ys, s.hardening, s.xi, true, s.softening, mu, lam);
            if (ys != ys0) s.yield_stress[p] = ys;
            if (mu != mu0) { s.mu[p] = mu; s.lam[p] = lam; }
        }
        store_m3(s.F, p, F);
        const float J = m3_det(F);
        if (material == 6) tau = stress_water(J, s.bulk[p]);
        else if (material == 0 || material == 5) {
            M3 R;
            if (polar_rotation(F, R)) tau = stress_fcr_R(F, R, J, mu, lam);
            else { M3 U, V; V3 sig; svd3(F, U, sig, V); tau = stress_fcr(F, U, V, J, mu, lam); }
        } else if (material >= 1 && material <= 3) {
            M3 U, V; V3 sig;
            svd3(F, U, sig, V);
            if (material == 2) tau = stress_drucker_prager(F, U, V, sig, mu, lam);
            else tau = stress_stvk(F, U, V, sig, mu, lam);
        }
        {
            const M3 tt = m3_t(tau);
#pragma unroll
            for (int i = 0; i < 9; ++i) tau.m[i] = (tau.m[i] + tt.m[i]) / 2.0f;
        }
        if (s.write_all) store_m3(s.stress, p, tau);

        {
            const float r = s.rpic_damping;
            const M3 Ct = m3_t(C);
            M3 Cn;
#pragma unroll
            for (int i = 0; i < 9; ++i) Cn.m[i] = (1.0f - r) * C.m[i] + r / 2.0f * (C.m[i] - Ct.m[i]);
            C = (r < -0.001f) ? m3_zero() : Cn;
        }
        vol = s.vol[p];
        scatter = true;
      } while (false);

      // ---- p2g_apic_with_stress (mpm_utils.py:338-394) into the warp's private halo copy. Warp-collective, skipped by
      //      warps without work. All lanes walk their 27 nodes in the same (i,j,k) order, so two lanes touch the same
      //      node in the same step only if they share the base cell: one match.any per particle gives each lane its
      //      turn, and plain read-modify-writes are race free.
      if (s.do_p2g && __any_sync(0xffffffffu, scatter)) {
        const Weights W = bspline_t(s.inv_dx, px, py, pz);
        const int lx = W.bx - bx0, ly = W.by - by0, lz = W.bz - bz0;
        const bool in_tile = scatter && (unsigned)lx < (unsigned)kTile && (unsigned)ly < (unsigned)kTile && (unsigned)lz < (unsigned)kTile;
        const int key = in_tile ? (lx * kTile + ly) * kTile + lz : (int)(kTile * kTile * kTile + lane);   // dummy keys are unique
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(peers & lt_mask);
        const int rounds = __reduce_max_sync(0xffffffffu, __popc(peers));
        float4* nbase = my_acc + (in_tile ? (lx * kHalo + ly) * kHalo + lz : 0);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const V3 dpos = {((float)i - W.fx[0]) * s.dx, ((float)j - W.fx[1]) * s.dx, ((float)k - W.fx[2]) * s.dx};
                    const float weight = W.w[0][i] * W.w[1][j] * W.w[2][k];
                    const V3 dweight = {W.dw[0][i] * W.w[1][j] * W.w[2][k] * s.inv_dx,
                                        W.w[0][i] * W.dw[1][j] * W.w[2][k] * s.inv_dx,
                                        W.w[0][i] * W.w[1][j] * W.dw[2][k] * s.inv_dx};
                    const V3 sd = m3_mulv(tau, dweight);
                    const V3 cd = m3_mulv(C, dpos);
                    const float wm = weight * mass;
                    const float ax = wm * (vx + cd.x) + dt * (-vol * sd.x);
                    const float ay = wm * (vy + cd.y) + dt * (-vol * sd.y);
                    const float az = wm * (vz + cd.z) + dt * (-vol * sd.z);
                    float4* node = nbase + (i * kHalo + j) * kHalo + k;
                    for (int rr = 0; rr < rounds; ++rr) {
                        if (in_tile && rank == rr) {
                            float4 a = *node;
                            a.x += ax; a.y += ay; a.z += az; a.w += wm;
                            *node = a;
                        }
                        __syncwarp();
                    }
                    if (scatter && !in_tile) {
                        const int ix = W.bx + i, iy = W.by + j, iz = W.bz + k;
                        if ((unsigned)ix < (unsigned)n && (unsigned)iy < (unsigned)n && (unsigned)iz < (unsigned)n)
                            ptx::red_add_v4(reinterpret_cast<float*>(s.mv_write + ((size_t)ix * n + iy) * n + iz), ax, ay, az, wm);
                    }
                }
      }
    }

    // ---- flush the tile's halo: one vector red per touched node
    if (s.do_p2g) {
        __syncthreads();
        for (int l = tid; l < kHaloNodes; l += blockDim.x) {
            float4 a = wacc[0][l];
#pragma unroll
            for (int w = 1; w < kTiledThreads / 32; ++w) {
                const float4 b = wacc[w][l];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (a.w != 0.f || a.x != 0.f || a.y != 0.f || a.z != 0.f) {
                const int i = l / (kHalo * kHalo), j = (l / kHalo) % kHalo, k = l % kHalo;
                const int gx = bx0 + i, gy = by0 + j, gz = bz0 + k;
                if (gx < n && gy < n && gz < n)
                    ptx::red_add_v4(reinterpret_cast<float*>(s.mv_write + ((size_t)gx * n + gy) * n + gz), a.x, a.y, a.z, a.w);
            }
        }
    }

    // ---- substep epilogue (only launches that ran a g2p advance the clock): next parity slot
    if (s.do_g2p && blockIdx.x == 0 && tid == 0) {
        for (int k = 0; k < s.n_bc; ++k) {
            const DevBC& bc = s.bcs[k];
            float qx = s.pts_in[3 * k], qy = s.pts_in[3 * k + 1], qz = s.pts_in[3 * k + 2];
            if (bc.kind == PIXIE_BC_CUBOID && t_d >= (double)bc.start_time && t_d < (double)bc.end_time) {
                qx = (float)((double)qx + dt_d * (double)bc.velocity[0]);
                qy = (float)((double)qy + dt_d * (double)bc.velocity[1]);
                qz = (float)((double)qz + dt_d * (double)bc.velocity[2]);
            }
            s.pts_out[3 * k] = qx; s.pts_out[3 * k + 1] = qy; s.pts_out[3 * k + 2] = qz;
        }
        *s.time_out = t_d + dt_d;
    }
}

// ------------------------------------------------------------------------------------ sort / permute
__global__ void tiled_count_kernel(const float* __restrict__ x, int n, float inv_dx, int n_grid, int nt, int* __restrict__ keys, int* __restrict__ counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int t[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int b = (int)(x[3 * p + a] * inv_dx - 0.5f);
        b = max(0, min(n_grid - 1, b));
        t[a] = b / kTile;
    }
    const int key = (t[0] * nt + t[1]) * nt + t[2];
    keys[p] = key;
    atomicAdd(counts + key, 1);
}
// single block: exclusive scan of counts -> off[0..ntiles], compaction of occupied tiles
__global__ void __launch_bounds__(1024) tiled_scan_kernel(const int* __restrict__ counts, int ntiles, int* __restrict__ off, int* __restrict__ cursor,
                                                          int* __restrict__ occ, int* __restrict__ n_occ) {
    __shared__ int part[1024];
    __shared__ int occ_part[1024];
    const int tid = threadIdx.x;
    const int per = (ntiles + 1023) / 1024;
    const int b = tid * per, e = min(ntiles, b + per);
    int s = 0, oc = 0;
    for (int i = b; i < e; ++i) { s += counts[i]; oc += counts[i] > 0; }
    part[tid] = s; occ_part[tid] = oc;
    __syncthreads();
    if (tid == 0) {
        int acc = 0, oacc = 0;
        for (int i = 0; i < 1024; ++i) { const int t = part[i]; part[i] = acc; acc += t; const int o = occ_part[i]; occ_part[i] = oacc; oacc += o; }
        off[ntiles] = acc;
        *n_occ = oacc;
    }
    __syncthreads();
    int acc = part[tid], oacc = occ_part[tid];
    for (int i = b; i < e; ++i) {
        off[i] = acc; cursor[i] = acc;
        if (counts[i] > 0) occ[oacc++] = i;
        acc += counts[i];
    }
}
__global__ void tiled_place_kernel(const int* __restrict__ keys, int n, int* __restrict__ cursor, int* __restrict__ order) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    order[atomicAdd(cursor + keys[p], 1)] = p;        // order[new position] = old position
}

struct PermuteArgs {
    const int* order;          // destination index -> source index
    int n;
    const float *x, *v, *C, *F, *Ft, *stress, *mass, *vol, *mu, *lam, *bulk, *ys;
    const int *material, *selection, *perm_src;       // perm_src == nullptr: source is the caller's order (perm = order)
    float *ox, *ov, *oC, *oF, *oFt, *ostress, *omass, *ovol, *omu, *olam, *obulk, *oys;
    int *omaterial, *oselection, *operm;
};
__global__ void tiled_permute_kernel(const PermuteArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n) return;
    const int q = a.order[p];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a.ox[3 * p + k] = a.x[3 * q + k]; a.ov[3 * p + k] = a.v[3 * q + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        a.oC[9 * p + k] = a.C[9 * q + k]; a.oF[9 * p + k] = a.F[9 * q + k]; a.oFt[9 * p + k] = a.Ft[9 * q + k];
        a.ostress[9 * p + k] = a.stress[9 * q + k];
    }
    a.omass[p] = a.mass[q]; a.ovol[p] = a.vol[q]; a.omu[p] = a.mu[q]; a.olam[p] = a.lam[q]; a.obulk[p] = a.bulk[q]; a.oys[p] = a.ys[q];
    a.omaterial[p] = a.material[q]; a.oselection[p] = a.selection[q];
    a.operm[p] = a.perm_src ? a.perm_src[q] : q;
}
// sorted -> caller's arrays (everything a substep can modify)
struct UnsortArgs {
    const int* perm; int n;
    const float *x, *v, *C, *F, *Ft, *stress, *mu, *lam, *ys;
    float *ox, *ov, *oC, *oF, *oFt, *ostress, *omu, *olam, *oys;
};
__global__ void tiled_unsort_kernel(const UnsortArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n) return;
    const int q = a.perm[p];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a.ox[3 * q + k] = a.x[3 * p + k]; a.ov[3 * q + k] = a.v[3 * p + k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        a.oC[9 * q + k] = a.C[9 * p + k]; a.oF[9 * q + k] = a.F[9 * p + k]; a.oFt[9 * q + k] = a.Ft[9 * p + k];
        a.ostress[9 * q + k] = a.stress[9 * p + k];
    }
    a.omu[q] = a.mu[p]; a.olam[q] = a.lam[p]; a.oys[q] = a.ys[p];
}

// Bring-up probe (not part of the library): can a K-major SWIZZLE_128B tcgen05 operand start at an ARBITRARY 128-byte row of a
// TMA-style slab?  A 3x3x3 convolution whose input plane is stored flattened needs tap (kh, kw) = the same slab read from row
// kh*(W+2) + kw; rows that are not multiples of the 8-row swizzle atom need the descriptor's base-offset field (bits 49-51).
// The slab is written with the SW128 pattern (16-byte chunk index XOR row%8, as TMA does), B is an identity tile, so
// D[m][n] must equal A[m + off][n].  Prints, per row offset, whether base_offset = 0 / (off % 8) reproduce it.
#include "ptx.cuh"
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

using namespace pixie::ptx;

constexpr int kRows = 160;   // slab rows (128 + offsets up to 24)

__global__ void __launch_bounds__(128, 1) rowoff_kernel(int off, int use_base_offset, float* out /* [128][64] */) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // [kRows][128 B]
    uint8_t* sB = smem + 32 * 1024;           // [64][128 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kRows * 64; i += blockDim.x) {
        const int r = i / 64, c = i % 64;
        const float v = (float)(r % 32) + (float)(c % 32) / 32.0f;
        const int chunk = (c / 8) ^ (r % 8);
        *reinterpret_cast<__half*>(sA + r * 128 + chunk * 16 + (c % 8) * 2) = __float2half(v);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        const int chunk = (k / 8) ^ (n % 8);
        *reinterpret_cast<__half*>(sB + n * 128 + chunk * 16 + (k % 8) * 2) = __float2half(n == k ? 1.0f : 0.0f);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_f16(128, 64);
            const uint32_t a_addr = smem_u32(sA) + (uint32_t)off * 128u;
            uint64_t da = make_sw128_desc(a_addr, 1024);
            if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
            const uint64_t db = make_sw128_desc(smem_u32(sB), 1024);
            for (int k4 = 0; k4 < 4; ++k4) umma_f16(tmem_base, da + (uint64_t)(2 * k4), db + (uint64_t)(2 * k4), idesc, k4 ? 1u : 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    bool done = false;
    for (int it = 0; it < (1 << 22) && !done; ++it) done = mbar_try_wait(&bar, 0);      // bounded: a faulting MMA must not hang the box
    if (!done) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
    tc_fence_after();
    // every warp reads its 32 TMEM lanes
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

int main() {
    float* d; cudaMalloc(&d, 128 * 64 * 4);
    static float h[128 * 64];
    cudaFuncSetAttribute(rowoff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int all_ok_bo = 1;
    for (int off : {0, 1, 2, 3, 5, 7, 8, 9, 18, 19, 20, 25}) {
        int ok[2] = {0, 0};
        for (int ubo = 0; ubo < 2; ++ubo) {
            cudaMemset(d, 0xFF, 128 * 64 * 4);
            rowoff_kernel<<<1, 128, 64 * 1024>>>(off, ubo, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("off=%d base_offset=%d: CUDA error %s\n", off, ubo, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    const float want = (float)((m + off) % 32) + (float)(n % 32) / 32.0f;
                    if (h[m * 64 + n] != want) ++bad;
                }
            ok[ubo] = bad == 0;
            if (bad && off < 3) printf("   off=%d ubo=%d first row got %g %g %g %g want %g ...\n", off, ubo, h[0], h[1], h[64], h[65], (float)(off % 32));
        }
        printf("row offset %2d: base_offset=0 %s | base_offset=(addr>>7)&7 %s\n", off, ok[0] ? "OK  " : "FAIL", ok[1] ? "OK  " : "FAIL");
        all_ok_bo &= ok[1];
    }
    printf("SUMMARY arbitrary-row starts with base_offset: %s\n", all_ok_bo ? "WORK" : "DO NOT WORK");
    return 0;
}

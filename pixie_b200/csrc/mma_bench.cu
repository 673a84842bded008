// Microbenchmark of tcgen05.mma issue patterns (bring-up tool, not part of the library).
// One CTA per SM, operands are whatever is in shared memory (values irrelevant), SW128 K-major.
// Reports cycles per 128xNx16 MMA for: dependent chains on one accumulator, round-robin over several
// accumulators, and A-operand reuse, at N = 64 / 128 / 192 / 256.
#include "ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>

using namespace pixie::ptx;

struct Result { long long cycles; };

// mode 0: all MMAs into accumulator 0            (dependent chain)
// mode 1: round-robin over `nacc` accumulators   (independent chains, switching every instruction)
// mode 2: runs of `run` MMAs per accumulator, then switch
// b_tiles > 1: the B operand cycles through `b_tiles` tiles `b_stride16` (16 B units) apart, changing every 4 MMAs (one 64-wide K block
// per tile, as in the conv kernel); d_shift: the accumulator base moves by d_shift columns every `run` MMAs (mode 3).
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int N, int iters, int mode, int nacc, int run, int a_stride16, int b_tiles, int b_stride16, int d_shift, Result* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 1) {
        const uint32_t idesc = make_idesc_f16(128, (uint32_t)N);
        const uint64_t dfix = make_sw128_desc(0, 1024);
        const uint32_t a16 = smem_u32(smem) >> 4, b16 = (smem_u32(smem) + 48 * 1024) >> 4;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            // warm-up
            for (int i = 0; i < 8; ++i) umma_f16(tmem_base, dfix | a16, dfix | b16, idesc, i ? 1u : 0u);
            umma_commit(&bar);
        }
        __syncwarp();
        while (!mbar_try_wait(&bar, 0)) {}
        t0 = clock64();
        if (elect_one()) {
            int acc = 0, inrun = 0;
            if (mode == 5) {
                // latency: `run` MMAs, commit, wait for the mbarrier, repeat (issue -> retire -> visible to the issuing thread)
                const uint32_t lo = (uint32_t)dfix, hi = (uint32_t)(dfix >> 32);
                uint32_t par = 1;
                for (int i = 0; i < iters; i += run) {
                    for (int k = 0; k < run; ++k)
                        umma_f16_lohi<true>(tmem_base, (lo | a16) + (uint32_t)((k & 3) * 2), (lo | b16) + (uint32_t)((k & 3) * 2), hi, idesc);
                    umma_commit(&bar);
                    while (!mbar_try_wait(&bar, par)) {}
                    par ^= 1;
                }
            } else if (mode == 4) {
                // fully unrolled: 12 MMAs per iteration, every operand offset an immediate (the conv kernel's issue pattern);
                // the rolled loop below spends 50-70 cycles per iteration on its own index arithmetic and hides the pipe's rate
                const uint32_t lo = (uint32_t)dfix, hi = (uint32_t)(dfix >> 32);
                for (int i = 0; i < iters; i += 12) {
                    const uint32_t d = tmem_base + (uint32_t)(acc * d_shift);
                    const uint32_t a0 = lo | a16, b0 = lo | b16;
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        umma_f16_lohi<true>(d, a0 + (uint32_t)((k >> 2) * 128 + (k & 3) * 2), b0 + (uint32_t)((k >> 2) * 1536 + (k & 3) * 2), hi, idesc);
                    acc = (acc + 1 == nacc) ? 0 : acc + 1;
                }
            } else
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tmem_base + (uint32_t)(mode == 3 ? acc * d_shift : acc * N);
                const uint64_t da = dfix | (uint64_t)((a16 + (uint32_t)((i & 3) * 2) + (uint32_t)(((i >> 2) % 6) * a_stride16)) & 0x3FFF);
                const uint64_t db = dfix | (uint64_t)((b16 + (uint32_t)((i & 3) * 2) + (uint32_t)(((i >> 2) % b_tiles) * b_stride16)) & 0x3FFF);
                umma_f16(d, da, db, idesc, 1u);
                if (mode == 1) { acc = (acc + 1 == nacc) ? 0 : acc + 1; }
                else if (mode >= 2) { if (++inrun == run) { inrun = 0; acc = (acc + 1 == nacc) ? 0 : acc + 1; } }
            }
            if (mode != 5) umma_commit(&bar);
        }
        __syncwarp();
        if (mode != 5) while (!mbar_try_wait(&bar, 1)) {}
        t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) out->cycles = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int main() {
    Result* d; cudaMalloc(&d, sizeof(Result));
    cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4092;   // multiple of 12
    struct Cfg { int N, mode, nacc, run, astr; const char* name; int bt = 1, bstr = 0, dshift = 0; };
    const Cfg cfgs[] = {
        {64, 0, 1, 0, 0, "N=64  chain, same A"},      {64, 0, 1, 0, 128, "N=64  chain, A moves 2KB every 4"},
        {128, 0, 1, 0, 128, "N=128 chain"},           {192, 0, 1, 0, 128, "N=192 chain"},          {256, 0, 1, 0, 128, "N=256 chain"},
        {64, 1, 2, 0, 128, "N=64  round-robin 2 acc"}, {64, 1, 4, 0, 128, "N=64  round-robin 4 acc"},
        {64, 2, 3, 12, 128, "N=64  runs of 12 over 3 acc"}, {64, 2, 4, 36, 128, "N=64  runs of 36 over 4 acc"},
        {128, 1, 2, 0, 128, "N=128 round-robin 2 acc"}, {32, 0, 1, 0, 128, "N=32  chain"}, {16, 0, 1, 0, 128, "N=16  chain"},
        {64, 0, 1, 0, 128, "N=64  chain, B cycles 9 tiles", 9, 512, 0},   {128, 0, 1, 0, 128, "N=128 chain, B cycles 4 tiles", 4, 1024, 0},
        {192, 0, 1, 0, 128, "N=192 chain, B cycles 3 tiles", 3, 1536, 0},  {192, 3, 2, 12, 128, "N=192 B cycles, D shifts 64 cols /12", 3, 1536, 64},
        {128, 3, 2, 12, 128, "N=128 B cycles, D shifts 64 cols /12", 4, 1024, 64}, {192, 3, 4, 12, 128, "N=192 B cycles, D 4 x 64 /12", 3, 1536, 64},
        {64, 0, 1, 0, 0, "N=64  chain, same A, B cycles 9", 9, 512, 0},   {256, 0, 1, 0, 128, "N=256 chain, B cycles 2", 2, 2048, 0},
        {16, 4, 1, 0, 0, "unrolled N=16  one acc", 1, 0, 0},   {32, 4, 1, 0, 0, "unrolled N=32  one acc", 1, 0, 0},
        {64, 4, 1, 0, 0, "unrolled N=64  one acc", 1, 0, 0},   {64, 4, 4, 0, 0, "unrolled N=64  D +64 /12, 4 acc", 1, 0, 64},
        {128, 4, 1, 0, 0, "unrolled N=128 one acc", 1, 0, 0},  {128, 4, 3, 0, 0, "unrolled N=128 D +64 /12, 3 pos", 1, 0, 64},
        {192, 4, 1, 0, 0, "unrolled N=192 one acc", 1, 0, 0},  {192, 4, 2, 0, 0, "unrolled N=192 D +64 /12, 2 pos", 1, 0, 64},
        {64, 5, 1, 1, 0, "latency: 1 x N=64 + commit + wait", 1, 0, 0},   {64, 5, 1, 12, 0, "latency: 12 x N=64 + commit + wait", 1, 0, 0},
        {128, 5, 1, 1, 0, "latency: 1 x N=128 + commit + wait", 1, 0, 0}, {128, 5, 1, 12, 0, "latency: 12 x N=128 + commit + wait", 1, 0, 0},
        {256, 4, 1, 0, 0, "unrolled N=256 one acc", 1, 0, 0},  {256, 4, 2, 0, 0, "unrolled N=256 D +128 /12, 2 pos", 1, 0, 128},
    };
    for (const Cfg& c : cfgs) {
        for (int grid : {1, 148}) {
            mma_bench_kernel<<<grid, 128, 200 * 1024>>>(c.N, iters, c.mode, c.nacc, c.run, c.astr, c.bt, c.bstr, c.dshift, d);
            cudaError_t e = cudaDeviceSynchronize();
            Result r{}; cudaMemcpy(&r, d, sizeof(r), cudaMemcpyDeviceToHost);
            printf("%-34s grid=%3d  %7.1f cycles/MMA  (ideal %d)  %s\n", c.name, grid, (double)r.cycles / iters, c.N / 2,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}

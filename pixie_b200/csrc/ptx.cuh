// Thin inline-PTX wrappers for the sm_100a features the hot kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
// Nothing here is portable: this header only compiles for sm_100a.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pixie {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a wrong descriptor must produce a failed test, not a hung GPU box.
// `*abort_flag` (shared) is raised on timeout so every role in the CTA bails out.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        if (mbar_try_wait(bar, parity)) return true;
        if ((it & 1023u) == 1023u && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
          "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        :
        : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0),
          "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc];  kind::f16 (fp16/bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f8f6f4 (8-bit float operands, K = 32 per instruction, fp32 accumulate): twice the MAC rate of kind::f16
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        :
        : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <bool kAccumulate>
__device__ __forceinline__ void umma_f8_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
    if (kAccumulate) {
        asm volatile(
            "{\n\t"
            ".reg .b64 da, db;\n\t"
            ".reg .pred p;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.eq.b32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n\t"
            "}\n"
            :
            : "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .b64 da, db;\n\t"
            ".reg .pred p;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n\t"
            "}\n"
            :
            : "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
            : "memory");
    }
}
// Same, with the two descriptors given as (low word, shared high word): only the 14-bit start address in the low
// word differs between operands / K steps, so the issuing thread needs one 32-bit add per operand per MMA.
template <bool kAccumulate>
__device__ __forceinline__ void umma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
    if (kAccumulate) {
        asm volatile(
            "{\n\t"
            ".reg .b64 da, db;\n\t"
            ".reg .pred p;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.eq.b32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
            "}\n"
            :
            : "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .b64 da, db;\n\t"
            ".reg .pred p;\n\t"
            "mov.b64 da, {%1, %3};\n\t"
            "mov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, 0, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
            "}\n"
            :
            : "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc)
            : "memory");
    }
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i = lane i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (r points at 32 consecutive array elements).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows at 128 B pitch, 8-row atoms
// of 1024 B). `sbo_bytes` = distance between consecutive 8-row atoms.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);         // start address   [0,14)
    d |= static_cast<uint64_t>(1) << 16;                             // LBO (unused for SW128 K-major)
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;     // SBO             [32,46)
    d |= static_cast<uint64_t>(1) << 46;                             // descriptor version (sm_100)
    d |= static_cast<uint64_t>(2) << 61;                             // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: A,B = fp16, D = fp32, both operands K-major, M x N tile.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)            // D format: F32
           | (0u << 7)          // A format: F16
           | (0u << 10)         // B format: F16
           | ((N >> 3) << 17)   // N / 8
           | ((M >> 4) << 24);  // M / 16
}

// kind::f8f6f4 instruction descriptor with A = B = E5M2, D = fp32, both operands K-major.
__host__ __device__ __forceinline__ uint32_t make_idesc_e5m2(uint32_t M, uint32_t N) {
    return (1u << 4)            // D format: F32
           | (1u << 7)          // A format: E5M2 (E4M3 = 0)
           | (1u << 10)         // B format: E5M2
           | ((N >> 3) << 17)   // N / 8
           | ((M >> 4) << 24);  // M / 16
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-byte sector per thread per instruction
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
// ---------------------------------------------------------------- global red / ld helpers
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}

// Predicated form: no branch around the instruction (27 of these per particle sit in a fully unrolled loop; as
// `if (ok) red` each cost a divergent branch — r02 ncu: branch_resolving was the top stall on that line)
__device__ __forceinline__ void red_add_v4_if(bool ok, float* addr, float a, float b, float c, float d) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t"
        "}\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)ok)
        : "memory");
}

}  // namespace ptx
}  // namespace pixie

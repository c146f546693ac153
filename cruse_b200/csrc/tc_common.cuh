// tc_common.cuh -- sm_100a tensor-core plumbing shared by the GEMM-shaped kernels:
// mbarrier, TMA (cp.async.bulk.tensor), TMEM allocation, tcgen05.mma / commit / ld wrappers and
// the shared-memory / instruction descriptors (bit layouts as in the PTX ISA "tcgen05" chapter).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cruse {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive without release semantics: the signal does not wait for the thread's earlier global stores to be performed
// (use when the barrier only hands back on-chip state, e.g. "I have finished reading this TMEM accumulator")
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// the same for waits that are expected to be long (a producer waiting for the consumer to free a ring slot): back off
// with nanosleep between polls so that the spinning warps do not take issue slots from the warps doing the work
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns = 256) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}

// one lane of a fully converged warp (the compiler can then keep single-thread tcgen05 issue on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n"
        : "+r"(pred));
    return pred != 0;
}

// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tile load: c0 = innermost (contiguous) coordinate, c1 = row coordinate, both in elements
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---------------------------------------------------------------- TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// K-major operand tile in the canonical SWIZZLE_128B layout: rows of 128 bytes, 8-row groups of 1024 B
// (SBO), exactly what a TMA box {32 x fp32 | 64 x bf16, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes.
// The tile base must be 1024-byte aligned; stepping K inside the 128-byte row = adding bytes to the start.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset = 1024 B  [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell) [46,48)
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B [61,64)
    return d;
}
// MN-major 32-bit operand tile: stored [K rows][32 fp32 of M|N = 128 bytes].  For 4-byte elements tcgen05 knows ONE
// MN-major layout, "128-byte swizzle with 32-byte atoms" (layout type 1; a plain SWIZZLE_128B descriptor reads zeros):
// within a 128-byte row the four 32-byte chunks are XOR-ed with (row & 3), the pattern repeating every 4 rows = 512 B --
// what a TMA box {32 x fp32, rows} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.  Consecutive 32-element blocks of M|N
// lie `lbo_bytes` apart (leading byte offset), consecutive 4-row K groups 512 B apart (stride byte offset).  One tf32 MMA
// (UMMA_K = 8) reads two K groups of every block: stepping K = adding 1024 B to the start.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128_32b(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset [16,30)
    d |= (uint64_t)(512 >> 4) << 32;                // stride byte offset = 512 B  [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell) [46,48)
    d |= (uint64_t)1 << 61;                         // layout type SWIZZLE_128B_BASE32B [61,64)
    return d;
}
// instruction descriptor: D fp32, A/B of `fmt` (0 f16, 1 bf16, 2 tf32), shape M x N; a_mn / b_mn = 1: that operand's
// shared-memory tile is MN-major (bits 15 / 16), else K-major
__host__ __device__ constexpr uint32_t instr_desc(int fmt, int M, int N, bool a_mn = false, bool b_mn = false) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, one CTA.  Issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane i <- TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc

// ---------------------------------------------------------------- host: tensor maps (driver entry point, no libcuda link)
// 2-D fp32 tensor [rows, cols] with row pitch `pitch_bytes`, box = {32 cols (128 B), box_rows}, SWIZZLE_128B,
// out-of-bounds elements read as zero.  `as_tf32` makes the TMA unit round fp32 -> tf32 on the way in.
// `atom32` selects the 128-byte swizzle with 32-byte atoms (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), the only layout tcgen05
// accepts for an MN-major 32-bit operand (smem_desc_mn_sw128_32b).
int make_tmap_2d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows,
                 bool as_tf32, bool atom32 = false);

}  // namespace cruse

// gru_seq_tc.cu -- the GRU recurrence (the T sequential steps of nn.GRU, model/cruse_net.py:23-31,43-50
// of the reference; cuDNN's persistent RNN kernel on the reference path) on the 5th-generation tensor
// cores, one thread-block cluster per (group, PAIR of 16-utterance slices).
//
// Per step the recurrent half of the cell is   pre[3H] = W_hh[g] . h_{t-1}   for every utterance: a
// [3H x H] . [H x 16] product per slice.  The cluster splits the H hidden units 32 per CTA (H = 256 -> 8
// CTAs); CTA c keeps the 96 gate rows {r,z,n} x {its 32 units} of W_hh resident in TENSOR MEMORY for the
// whole sequence as the A operand of ONE tcgen05.mma shape (M = 128 [96 used], N = 16 utterances, K = H,
// kind::tf32, fp32 accumulation in TMEM; lane = gate row, column = k, 256 of the 512 TMEM columns).
// (A first version kept the slice in shared memory: every step then re-streams 128 KB through the
// 128 B/clk shared-memory port, measured 47 clk per K=8 MMA; from TMEM the same MMA is not port-bound.)
// The B operand is h_{t-1} itself in shared memory ([16 utterances][H], K-major, 128-byte swizzle), double buffered.
//
// The step is a latency chain (h arrival -> 32 MMAs -> commit -> tcgen05.ld -> gate math -> scatter), so one
// cluster runs TWO independent slices of 16 utterances software-pipelined against each other: warp 8 issues the
// MMAs of slice 0 then slice 1 every step (both read the same W_hh in TMEM, separate accumulators), warps 0-3
// and 4-7 are the two slices' warpgroups.  While one slice waits for its h exchange and tensor-core pass the
// other does its gate math, which roughly halves the time per (utterance, step) and halves the clusters a
// layer needs -- so both GRU layers of the bottleneck fit on the GPU side by side (cruse_net.GGRU._wavefront).
//
// Per slice and step: the accumulator (lane = gate row, column = utterance) is pulled out with tcgen05.ld,
// transposed through a small shared-memory pad so that every thread owns (utterance, 4 adjacent units), the
// gates are evaluated in fp32 (the z*h_{t-1} carry uses the thread's own full-precision register copy, only
// the matmul operand is rounded to tf32), and the new h slice -- exactly k-block `c` of everybody's next B
// operand -- is scattered to all CTAs of the cluster with st.async (16-byte DSMEM stores that credit the
// destination CTA's mbarrier).  No cluster barrier and no fence sits inside the time loop: step t+1's MMA is
// released by the byte count of h_t arriving; the slice warpgroups only meet on their own named barrier.
//
// x-projections (+ folded biases) come from the tcgen05 input GEMM (gru_ih_tc.cu) and are prefetched
// two steps ahead.  Optional `gates` output saves r, z, n and W_hn.h+b_hn for the backward pass.  Rows of
// xproj / y are addressed through (utterance, frame) strides so the same kernel runs whole sequences in
// frame order and time chunks of time-major buffers.
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <cstring>

// Operand format of the recurrent product.  1 (default): both operands travel and multiply as IEEE half (tcgen05 kind::f16, fp32
// accumulation) -- h_t lies in (-1, 1) and W_hh in (-1/sqrt(H), 1/sqrt(H)) at initialisation, where half has the SAME 10-bit
// mantissa as tf32 (below 2^-14 it is absolutely, not relatively, accurate: error <= 3e-8), so the numerics are those of the tf32
// path, while every step moves HALF the bytes through the cluster's distributed shared memory (measured: the step is bound by that
// exchange, ~17 B/clk per SM), issues 16 instead of 32 MMAs (K = 16 per instruction) and W_hh takes 128 instead of 256 TMEM columns.
// 0: tf32 operands (the round-1 path).
#ifndef CRUSE_SEQ_F16
#define CRUSE_SEQ_F16 1
#endif

namespace cg = cooperative_groups;

namespace cruse {
namespace {

constexpr int SQ_NB = 16;                    // utterances per slice = MMA N
constexpr int SQ_U = 32;                     // hidden units per CTA (= one 128-byte k-block of the B operand)
constexpr int SQ_WG = 128;                   // threads per slice warpgroup: thread -> (utterance wt/8, 4 adjacent units)
constexpr int SQ_THREADS = 2 * SQ_WG + 32;   // two slice warpgroups + the MMA-issuing warp
constexpr int SQ_TMEM_COLS = 512;            // A: up to 256 columns (K) + 2 x D: 16 columns -> whole TMEM (1 CTA/SM)
constexpr int SQ_D_COL = 256;                // accumulator column offset (slice s at + 16*s)
constexpr int SQ_H_KB = SQ_NB * 128;         // bytes of one B k-block tile: 16 rows x 128 bytes (32 tf32 / 64 half)
#if CRUSE_SEQ_F16
constexpr int SQ_KBE = 64;                   // K elements per 128-byte row of the B operand
constexpr int SQ_EB = 2;                     // bytes per operand element
#else
constexpr int SQ_KBE = 32;
constexpr int SQ_EB = 4;
#endif
constexpr int sq_nkb(int nc) { return (nc * SQ_U + SQ_KBE - 1) / SQ_KBE; }   // k-block tiles of one h buffer
constexpr int SQ_PRE_LD = SQ_NB + 1;         // padded leading dim of the gate-row x utterance pad

#ifdef CRUSE_SEQ_TIMING
// developer instrumentation (never built into the shipped library): per-step clock64 stamps of cluster 0 / CTA 0
__device__ long long g_seq_timing[8 * 2048];
#define SEQ_STAMP(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t < 2048 && (threadIdx.x & 31) == 0) g_seq_timing[t * 8 + (slot)] = clock64(); } while (0)
#else
#define SEQ_STAMP(slot) do {} while (0)
#endif

// Optional chunk-level synchronisation with OTHER kernels running beside this one (the two-layer wavefront without
// relaunches, cruse_net.GGRU._wavefront): before frame bounds[k] is prefetched the slice waits until wait[k] >= wait_target
// (set by a tiny kernel queued behind the producer of that chunk's x-projections), and after frame bounds[k+1]-1 is stored it
// adds 1 to done[k] (release) so that a consumer queued behind cruse_flag_wait may read y up to there.  nchunks == 0: off.
struct SeqSync {
    const unsigned* wait;
    unsigned* done;
    int* err;
    unsigned wait_target;
    int nchunks;
    int bounds[18];
    unsigned long long* trace;      // developer progress trace (cruse_debug_seq_trace), NULL = off: [128] step stamps + [2*16] chunk-wait stamps
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin (one thread) until *flag >= target; gives up after 2 s and records it, so a broken dependency never hangs the GPU
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned target, int* err) {
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_gpu(flag) < target) {
        __nanosleep(200);
        if (globaltimer_ns() - t0 > 2000000000ull) {
            if (err) atomicExch(err, 1);
            break;
        }
    }
}

struct SeqPtrs {
    const float* w_hh[CRUSE_MAX_GROUPS];
    const float* b_hh[CRUSE_MAX_GROUPS];
};

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B operand whose k-block tiles are `kb_bytes` apart
__device__ __forceinline__ uint32_t sw128_off(int row, int k, int kb_bytes) {
    const int kb = k / SQ_KBE, byte = (k % SQ_KBE) * SQ_EB;            // byte position inside the 128-byte row
    return (uint32_t)(kb * kb_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((((byte >> 4) ^ (row & 7))) << 4) + (byte & 15));
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {   // this warp's 32 lanes x 16 columns
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);       // .x (low 16 bits) = lo
    return *reinterpret_cast<const uint32_t*>(&h);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, half operands (two consecutive k per 32-bit TMEM cell of A), K = 16 per instruction
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, tf32: the A operand (lane = row, column = k) is read from tensor memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// fast gate nonlinearities (MUFU ex2 / rcp): abs error ~1e-6, far below the tf32 operand rounding of this kernel
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }
// one MUFU per gate (tanh.approx; sigmoid(x) = 0.5 tanh(x/2) + 0.5) instead of ex2 + rcp: two MUFU latencies less on the per-step
// critical path (measured r2: 1.14 -> 1.04 us per step, cfg-2 parity unchanged: mask 1.99e-4 vs 1.9e-4 of full scale, loss 2.7e-6).
// The training forward (gates saved for the BPTT) uses them too since the end of round 2 -- same gate arithmetic as inference;
// measured at cfg-3: step 4.19 -> 4.16 ms, worst gradient rel-L2 against the oracle 3.25e-2 -> 3.28e-2 (conv4.weight, tf32 mode),
// every gradient test unchanged (-DCRUSE_SEQ_TANH_TRAIN=0 restores the ex2/rcp gates there).
#ifndef CRUSE_SEQ_TANH
#define CRUSE_SEQ_TANH 1
#endif
#ifndef CRUSE_SEQ_TANH_TRAIN          // 1: the training forward (gates saved) uses the one-MUFU gates too
#define CRUSE_SEQ_TANH_TRAIN 1
#endif
__device__ __forceinline__ float tanh_mufu(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_mufu(float x) { return fmaf(0.5f, tanh_mufu(0.5f * x), 0.5f); }

template <int NC>
__global__ void __launch_bounds__(SQ_THREADS, 1)
gru_seq_tc_kernel(const float* __restrict__ xproj, const SeqPtrs ptrs, const float* __restrict__ h0, float* __restrict__ y,
                  float* __restrict__ hT, float* __restrict__ gates, int B, int T, int G, int H, int y_fs, int y_gs,
                  long long x_bs, long long x_ts, long long y_bs, long long y_ts, const SeqSync sync) {
    // row (b, t) of xproj is b*x_bs + t*x_ts, of y b*y_bs + t*y_ts ([B,T] frame order: (T,1); time-major [T,B]: (1,B))
    constexpr int NKB = sq_nkb(NC);                       // k-block tiles (16 rows x 128 B) of one h buffer
    constexpr int SLICE_BYTES = 2 * NKB * SQ_H_KB;       // two h buffers of one slice
    extern __shared__ uint8_t smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int g = blockIdx.y, bpair = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sH = smem_raw + (base - tc::smem_u32(smem_raw));          // [slice][buffer][NC k-blocks][16 x 128 B]
    float* sPre = reinterpret_cast<float*>(sH + 2 * SLICE_BYTES);      // [slice][96][PRE_LD]
    uint64_t* hbar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sPre + 2 * 96 * SQ_PRE_LD) + 7) & ~(uintptr_t)7);   // [slice][buffer]
    uint64_t* acc_full = hbar + 4;                                     // [slice]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

    // ---- warps 0-3 / 4-7 = slice 0 / 1: thread -> (utterance b, units u0 .. u0+3)
    const int sl = (warp >> 2) & 1;
    const int wt = tid & (SQ_WG - 1);
    const int b = wt >> 3, jq = wt & 7;
    const int u0 = rank * SQ_U + 4 * jq;
    const bool uvalid = u0 < H;                       // H % 4 == 0, so u0+3 < H as well
    const int bg = (bpair * 2 + sl) * SQ_NB + b;
    const bool valid = uvalid && bg < B && warp < 8;
    const bool act0 = (bpair * 2) * SQ_NB < B, act1 = (bpair * 2 + 1) * SQ_NB < B;   // slice has any utterance (CTA-uniform)
    const bool my_act = sl ? act1 : act0;

    for (int i = tid; i < 2 * SLICE_BYTES / 16; i += SQ_THREADS) reinterpret_cast<float4*>(sH)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) tc::mbar_init(&hbar[i], 1);
        tc::mbar_init(&acc_full[0], 1);
        tc::mbar_init(&acc_full[1], 1);
        tc::fence_barrier_init();
    }
    if (warp == 8) tc::tmem_alloc<SQ_TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // ---- W_hh slice -> tensor memory (A operand): lane q*32+j = gate q of hidden unit rank*32+j, column = k, tf32.
    //      A warp can only touch its own lane quadrant (warp % 4); warps 0-3 / 4-7 split the k-blocks.
    if (warp < 8) {
        const float* W = ptrs.w_hh[g];
        const int q = warp & 3, unit = rank * SQ_U + lane;
        const bool rowvalid = q < 3 && unit < H;
        const float* wrow = W + ((size_t)(rowvalid ? q : 0) * H + (rowvalid ? unit : 0)) * H;
        // two k-blocks (16 independent 16-byte loads) in flight per round: this prologue is paid once per time chunk
        for (int kb0 = (warp >> 2); kb0 < NC; kb0 += 4) {
            float v[2][32];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = (kb0 + 2 * u) * 32 + i * 4;
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rowvalid && kb0 + 2 * u < NC && k < H) f = __ldg(reinterpret_cast<const float4*>(wrow + k));
#if CRUSE_SEQ_F16
                    v[u][i * 4 + 0] = f.x; v[u][i * 4 + 1] = f.y; v[u][i * 4 + 2] = f.z; v[u][i * 4 + 3] = f.w;
#else
                    v[u][i * 4 + 0] = to_tf32(f.x); v[u][i * 4 + 1] = to_tf32(f.y); v[u][i * 4 + 2] = to_tf32(f.z); v[u][i * 4 + 3] = to_tf32(f.w);
#endif
                }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (kb0 + 2 * u < NC) {
#if CRUSE_SEQ_F16
                    uint32_t pk[16];                     // two consecutive k per TMEM cell: 32 k -> 16 columns
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = pack_half2(v[u][2 * i], v[u][2 * i + 1]);
                    tmem_st_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((kb0 + 2 * u) * 16), pk);
#else
                    tmem_st_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((kb0 + 2 * u) * 32), v[u]);
#endif
                }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // ---- h_0 -> buffer 0 of my slice (tf32 operand copy) and the thread's own full-precision copy
    float4 hold = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h0 && warp < 8 && my_act) {
        if (valid) hold = __ldg(reinterpret_cast<const float4*>(h0 + ((size_t)g * B + bg) * H + u0));
        uint8_t* dstb = sH + sl * SLICE_BYTES;
        // every CTA needs the whole h_0 of its utterances: thread -> (utterance wt/8, 16-byte chunk wt%8) of every k-block,
        // all NC loads in flight at once
        const int bb = wt >> 3, c4 = wt & 7;
        const int bgg = (bpair * 2 + sl) * SQ_NB + bb;
        float4 hv[NC];
#pragma unroll
        for (int kb = 0; kb < NC; ++kb) {
            const int k = kb * 32 + c4 * 4;
            hv[kb] = (bgg < B && k < H) ? __ldg(reinterpret_cast<const float4*>(h0 + ((size_t)g * B + bgg) * H + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kb = 0; kb < NC; ++kb)
#if CRUSE_SEQ_F16
            *reinterpret_cast<uint2*>(dstb + sw128_off(bb, kb * 32 + c4 * 4, SQ_H_KB)) =
                make_uint2(pack_half2(hv[kb].x, hv[kb].y), pack_half2(hv[kb].z, hv[kb].w));
#else
            *reinterpret_cast<float4*>(dstb + sw128_off(bb, kb * 32 + c4 * 4, SQ_H_KB)) =
                make_float4(to_tf32(hv[kb].x), to_tf32(hv[kb].y), to_tf32(hv[kb].z), to_tf32(hv[kb].w));
#endif
    }
    tc::fence_proxy_async_smem();      // generic-proxy smem writes (h_0) -> visible to the tensor core's async proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    cluster.sync();                    // every CTA's buffers + barriers exist before any remote store lands

    constexpr uint32_t STEP_BYTES = (uint32_t)NC * SQ_WG * 4 * SQ_EB;     // all of h_t of one slice: 16 utterances x 32*NC elements
    constexpr uint32_t idesc = tc::instr_desc(CRUSE_SEQ_F16 ? 0 /*f16*/ : 2 /*tf32*/, 128, SQ_NB);
    const size_t N3 = (size_t)3 * H;

    if (warp == 8) {
        // ================= MMA issuer: pre = W_slice . h_t for slice 0, then slice 1, every step =================
        // h_t (t >= 1) is arrival number (t-1)/2 on buffer t&1 of its slice -> wait parity ((t-1)>>1)&1
        int p = 0;
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (!(s ? act1 : act0)) continue;
                if (lane == 0) tc::mbar_expect_tx(&hbar[s * 2 + (p ^ 1)], STEP_BYTES);   // arm the barrier h_{t+1} will complete
                if (s == 0) SEQ_STAMP(0);
                if (t > 0) {
                    tc::mbar_wait(&hbar[s * 2 + p], (uint32_t)(((t - 1) >> 1) & 1));
                    tc::fence_proxy_async_smem();          // h_t was written by (remote) generic-proxy stores
                }
                tc::tc_fence_after();
                if (s == 0) SEQ_STAMP(1);
                const uint64_t bdesc0 = tc::smem_desc_sw128(base + (uint32_t)s * SLICE_BYTES + (uint32_t)p * (NKB * SQ_H_KB));
                if (tc::elect_one()) {
#if CRUSE_SEQ_F16
                    // K = 16 half per instruction = 32 bytes along the B row = 8 TMEM cells of A; 2 * NC instructions cover K = 32 * NC
#pragma unroll
                    for (int k16 = 0; k16 < 2 * NC; ++k16)
                        umma_f16_ts(tmem_base + SQ_D_COL + 16 * s, tmem_base + (uint32_t)(k16 * 8),
                                    bdesc0 + (uint64_t)(((k16 >> 2) * SQ_H_KB + (k16 & 3) * 32) >> 4), idesc, k16 ? 1u : 0u);
#else
#pragma unroll
                    for (int kb = 0; kb < NC; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_tf32_ts(tmem_base + SQ_D_COL + 16 * s, tmem_base + (uint32_t)(kb * 32 + ks * 8),
                                         bdesc0 + (uint64_t)((kb * SQ_H_KB + ks * 32) >> 4), idesc, (kb | ks) ? 1u : 0u);
#endif
                    tc::umma_commit(&acc_full[s]);
                }
                __syncwarp();
                if (s == 0) SEQ_STAMP(2);
            }
            p ^= 1;
        }
        // drain: the last step's scatter is still in flight towards my buffers -- wait for it before anybody exits
        if (T > 0) {
            if (act0) tc::mbar_wait(&hbar[0 + p], (uint32_t)(((T - 1) >> 1) & 1));
            if (act1) tc::mbar_wait(&hbar[2 + p], (uint32_t)(((T - 1) >> 1) & 1));
        }
    } else if (my_act) {
        // ================= slice warpgroup: accumulator -> gates -> h_{t+1} scatter =================
        float* myPre = sPre + sl * (96 * SQ_PRE_LD);
        const uint32_t tmem_d = tmem_base + SQ_D_COL + 16 * sl;
        // remote (shared::cluster) addresses of my 16-byte h slot and of the slice's barriers in every CTA of the cluster
        uint32_t rem_h[NC], rem_bar[NC];
        {
            const uint32_t lh = tc::smem_u32(sH) + (uint32_t)sl * SLICE_BYTES + sw128_off(b, rank * SQ_U + 4 * jq, SQ_H_KB);   // my units' slot
            const uint32_t lb = tc::smem_u32(&hbar[sl * 2]);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_h[c]) : "r"(lh), "r"(c));
                asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_bar[c]) : "r"(lb), "r"(c));
            }
        }
        const size_t xstep = (size_t)x_ts * G * N3;
        const size_t ystep = (size_t)y_ts * G * H;
        const size_t bsel = (size_t)(valid ? bg : 0);
        const float* xp = xproj + (bsel * (size_t)x_bs * G + g) * N3 + (uvalid ? u0 : 0);
        float* yp = y + bsel * (size_t)y_bs * G * H + (size_t)(uvalid ? u0 : 0) * y_fs + (size_t)g * y_gs;
        float* gp = gates ? gates + ((bsel * T) * G + g) * (4 * (size_t)H) + (uvalid ? u0 : 0) : nullptr;   // gates: always [B,T]
        float4 b_hn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (uvalid && ptrs.b_hh[g]) b_hn = __ldg(reinterpret_cast<const float4*>(ptrs.b_hh[g] + 2 * H + u0));

        // x-projections are streamed from HBM PF steps ahead of their use (a step is shorter than a DRAM round trip, and
        // the skip convs / input projections running beside the recurrence load the memory system)
#ifndef CRUSE_SEQ_PF
#define CRUSE_SEQ_PF 4
#endif
        constexpr int PF = CRUSE_SEQ_PF;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 xq_r[PF], xq_z[PF], xq_n[PF];                          // queue: [0] = current step, [PF-1] = newest prefetch
        int kw = 0, kd = 0;                                           // next chunk to wait for / to report
        int next_wait = sync.nchunks > 0 ? sync.bounds[0] : -1;       // first frame of chunk kw (-1: nothing to wait for)
        int next_done = sync.nchunks > 0 ? sync.bounds[1] : -1;       // one past the last frame of chunk kd
        auto chunk_gate = [&](int tp) {                               // called (WG-uniformly) before frame tp is prefetched
            if (tp == next_wait) {
                if (sync.wait) {
                    const bool tr = sync.trace && wt == 0 && sl == 0 && rank == 0 && g == 0 && bpair == 0 && kw < 16;
                    if (tr) sync.trace[128 + 2 * kw] = globaltimer_ns();
                    if (wt == 0) spin_until(sync.wait + kw, sync.wait_target, sync.err);
                    if (tr) sync.trace[128 + 2 * kw + 1] = globaltimer_ns();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + sl) : "memory");
                }
                ++kw;
                next_wait = kw < sync.nchunks ? sync.bounds[kw] : -1;
            }
        };
#pragma unroll
        for (int d = 0; d < PF; ++d) {
            xq_r[d] = xq_z[d] = xq_n[d] = zero4;
            if (d < T) chunk_gate(d);
            if (valid && d < T) {
                const float* xq = xp + (size_t)d * xstep;
                xq_r[d] = __ldcg(reinterpret_cast<const float4*>(xq));
                xq_z[d] = __ldcg(reinterpret_cast<const float4*>(xq + H));
                xq_n[d] = __ldcg(reinterpret_cast<const float4*>(xq + 2 * H));
            }
        }
        const int q = warp & 3;
        int p = 0;
        for (int t = 0; t < T; ++t) {
            // prefetch the x-projections of step t+PF while the tensor core works
            float4 nxr = zero4, nxz = zero4, nxn = zero4;
            if (t + PF < T) chunk_gate(t + PF);
            if (valid && t + PF < T) {
                const float* xq = xp + (size_t)(t + PF) * xstep;
                nxr = __ldcg(reinterpret_cast<const float4*>(xq));
                nxz = __ldcg(reinterpret_cast<const float4*>(xq + H));
                nxn = __ldcg(reinterpret_cast<const float4*>(xq + 2 * H));
            }
            const float4 xr = xq_r[0], xz = xq_z[0], xn = xq_n[0];
            tc::mbar_wait(&acc_full[sl], (uint32_t)(t & 1));
            tc::tc_fence_after();
            if (sl == 0 && wt == 0) SEQ_STAMP(3);
            if (q < 3) {
                // accumulator lanes 32*q.. = gate q of my CTA's 32 units; columns = the slice's utterances
                float v[16];
                tmem_ld_32x16(tmem_d + ((uint32_t)(q * 32) << 16), v);
                tc::tmem_ld_wait();
                float* dst = myPre + (q * 32 + lane) * SQ_PRE_LD;
#pragma unroll
                for (int c = 0; c < 16; ++c) dst[c] = v[c];
            }
            tc::tc_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + sl) : "memory");     // this slice's four warps only
            if (sl == 0 && wt == 0) SEQ_STAMP(4);
            float hn[4], gr[4], gz[4], gn[4], ghn[4];
            const float xrv[4] = {xr.x, xr.y, xr.z, xr.w}, xzv[4] = {xz.x, xz.y, xz.z, xz.w}, xnv[4] = {xn.x, xn.y, xn.z, xn.w};
            const float bhv[4] = {b_hn.x, b_hn.y, b_hn.z, b_hn.w}, hov[4] = {hold.x, hold.y, hold.z, hold.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float pr = myPre[(0 * 32 + 4 * jq + e) * SQ_PRE_LD + b];
                const float pz = myPre[(1 * 32 + 4 * jq + e) * SQ_PRE_LD + b];
                ghn[e] = myPre[(2 * 32 + 4 * jq + e) * SQ_PRE_LD + b] + bhv[e];
                if (CRUSE_SEQ_TANH && (CRUSE_SEQ_TANH_TRAIN || gp == nullptr)) {
                    gr[e] = sigmoid_mufu(xrv[e] + pr);
                    gz[e] = sigmoid_mufu(xzv[e] + pz);
                    gn[e] = tanh_mufu(xnv[e] + gr[e] * ghn[e]);
                } else {
                    gr[e] = sigmoid_fast(xrv[e] + pr);
                    gz[e] = sigmoid_fast(xzv[e] + pz);
                    gn[e] = tanh_fast(xnv[e] + gr[e] * ghn[e]);
                }
                hn[e] = valid ? ((1.f - gz[e]) * gn[e] + gz[e] * hov[e]) : 0.f;
            }
            hold = make_float4(hn[0], hn[1], hn[2], hn[3]);
            if (sl == 0 && wt == 0) SEQ_STAMP(7);
            // scatter my 16 bytes of h_{t+1} (tf32-rounded operand copy) into the other buffer of every CTA
            {
                const uint32_t poff = (uint32_t)(p ^ 1) * (NKB * SQ_H_KB), boff = (uint32_t)(p ^ 1) * 8;
#if CRUSE_SEQ_F16
                const uint32_t w0 = pack_half2(hn[0], hn[1]), w1 = pack_half2(hn[2], hn[3]);
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(
                                     rem_h[c] + poff),
                                 "r"(w0), "r"(w1), "r"(rem_bar[c] + boff)
                                 : "memory");
#else
                const uint32_t w0 = __float_as_uint(to_tf32(hn[0])), w1 = __float_as_uint(to_tf32(hn[1]));
                const uint32_t w2 = __float_as_uint(to_tf32(hn[2])), w3 = __float_as_uint(to_tf32(hn[3]));
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                                     rem_h[c] + poff),
                                 "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(rem_bar[c] + boff)
                                 : "memory");
#endif
            }
            if (sl == 0 && wt == 0) SEQ_STAMP(5);
            if (valid) {
                float* yo = yp + (size_t)t * ystep;
                if (y_fs == 1) {
                    *reinterpret_cast<float4*>(yo) = hold;
                } else {
                    yo[0] = hn[0]; yo[y_fs] = hn[1]; yo[2 * y_fs] = hn[2]; yo[3 * y_fs] = hn[3];
                }
                if (gp) {
                    float* go = gp + (size_t)t * G * 4 * H;
                    *reinterpret_cast<float4*>(go) = make_float4(gr[0], gr[1], gr[2], gr[3]);
                    *reinterpret_cast<float4*>(go + H) = make_float4(gz[0], gz[1], gz[2], gz[3]);
                    *reinterpret_cast<float4*>(go + 2 * H) = make_float4(gn[0], gn[1], gn[2], gn[3]);
                    *reinterpret_cast<float4*>(go + 3 * H) = make_float4(ghn[0], ghn[1], ghn[2], ghn[3]);
                }
            }
            if (sl == 0 && wt == 0) SEQ_STAMP(6);
            if (sync.trace && (t & 7) == 0 && wt == 0 && sl == 0 && rank == 0 && g == 0 && bpair == 0 && (t >> 3) < 128) sync.trace[t >> 3] = globaltimer_ns();
            if (t + 1 == next_done) {                                  // chunk kd is stored: publish it
                if (sync.done) {
                    __threadfence();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + sl) : "memory");
                    if (wt == 0) atomicAdd(sync.done + kd, 1u);
                }
                ++kd;
                next_done = kd < sync.nchunks ? sync.bounds[kd + 1] : -1;
            }
            p ^= 1;
#pragma unroll
            for (int d = 0; d + 1 < PF; ++d) { xq_r[d] = xq_r[d + 1]; xq_z[d] = xq_z[d + 1]; xq_n[d] = xq_n[d + 1]; }
            xq_r[PF - 1] = nxr; xq_z[PF - 1] = nxz; xq_n[PF - 1] = nxn;
        }
        if (hT && valid) *reinterpret_cast<float4*>(hT + ((size_t)g * B + bg) * H + u0) = hold;
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster.sync();
    if (warp == 8) tc::tmem_dealloc<SQ_TMEM_COLS>(tmem_base);
}

// The request is padded to 136 KB although the kernel uses ~78 KB: the recurrence holds all 512 TMEM columns of its SM, so a
// tensor-core GEMM CTA (gru_ih_tc.cu, 99.5 KB shared memory, 256 TMEM columns) co-scheduled on the same SM would sit in
// tcgen05.alloc until the recurrence CTA exits.  With 136 KB taken no such CTA fits beside it, and the input projections
// that run beside the wavefront go to the SMs the recurrence does not use.
constexpr size_t seq_smem_bytes(int NC) {
    const size_t used = 1024 + 2 * 2 * (size_t)sq_nkb(NC) * SQ_H_KB + (2 * 96 * SQ_PRE_LD + 4) * 4 + 128;
    return used > 136 * 1024 ? used : 136 * 1024;
}

template <int NC>
void seq_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int G, int npairs, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(NC, G, npairs);
    cfg.blockDim = dim3(SQ_THREADS);
    cfg.dynamicSmemBytes = seq_smem_bytes(NC);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

template <int NC>
int max_clusters_tc() {
    if (cudaFuncSetAttribute(gru_seq_tc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes(NC)) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    seq_cfg<NC>(cfg, attr, 1, 1024, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gru_seq_tc_kernel<NC>, &cfg) != cudaSuccess) return -2;
    return n;
}

template <int NC>
int launch_seq_tc_nc(const float* xproj, const SeqPtrs& ptrs, const float* h0, float* y, float* hT, float* gates, int B, int T,
                     int G, int H, int y_fs, int y_gs, long long x_bs, long long x_ts, long long y_bs, long long y_ts, const SeqSync& sync,
                     cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(gru_seq_tc_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes(NC)));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    seq_cfg<NC>(cfg, attr, G, (B + 2 * SQ_NB - 1) / (2 * SQ_NB), st);      // one cluster per (group, pair of 16-utterance slices)
    CRUSE_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_seq_tc_kernel<NC>, xproj, ptrs, h0, y, hT, gates, B, T, G, H, y_fs, y_gs, x_bs, x_ts, y_bs, y_ts, sync));
    return 0;
}

}  // namespace
}  // namespace cruse

using namespace cruse;

#ifdef CRUSE_SEQ_TIMING
extern "C" int cruse_debug_seq_timing(long long* out_host, int n) {
    return (int)cudaMemcpyFromSymbol(out_host, g_seq_timing, sizeof(long long) * n);
}
#endif

#define SEQ_TC_DISPATCH(nc, CALL)                 \
    switch (nc) {                                 \
        case 1: return CALL(1);                   \
        case 2: return CALL(2);                   \
        case 3: return CALL(3);                   \
        case 4: return CALL(4);                   \
        case 5: return CALL(5);                   \
        case 6: return CALL(6);                   \
        case 7: return CALL(7);                   \
        case 8: return CALL(8);                   \
        default: break;                           \
    }

extern "C" int cruse_gru_seq_tc_max_clusters(int H) {
#define CALL(N) max_clusters_tc<N>()
    SEQ_TC_DISPATCH((H + SQ_U - 1) / SQ_U, CALL)
#undef CALL
    set_error("gru_seq_tc_max_clusters: unsupported H=%d", H);
    return -1;
}

// developer instrumentation: the next flagged recurrence launches (one per layer) write their progress trace (globaltimer stamps of
// every 8th step and of the chunk waits of cluster 0) to buf[k * 160 ...] for k = 0, 1, ...; NULL switches it off
static unsigned long long* g_seq_trace = nullptr;
static int g_seq_trace_next = 0;
extern "C" int cruse_debug_seq_trace(void* device_buf) {
    g_seq_trace = static_cast<unsigned long long*>(device_buf);
    g_seq_trace_next = 0;
    return 0;
}

static int gru_seq_tc_impl(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h0, float* y,
                           float* hT, float* gates, int B, int T, int G, int H, int y_fs, int y_gs, long long x_bs, long long x_ts,
                           long long y_bs, long long y_ts, void* stream, const SeqSync* syncp = nullptr) {
    CRUSE_CHECK_ARG(xproj && y && w_hh, "gru_seq_fwd_tc: null pointer");
    SeqSync sync;
    memset(&sync, 0, sizeof(sync));
    if (syncp) sync = *syncp;
    if (syncp && g_seq_trace) sync.trace = g_seq_trace + 160 * ((g_seq_trace_next++) & 1);      // layer 1, layer 2, layer 1, ...
    CRUSE_CHECK_ARG(B > 0 && T >= 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && (H % 4) == 0 && H <= 256,
                    "gru_seq_fwd_tc: bad sizes B=%d T=%d G=%d H=%d (H%%4==0, H<=256, G<=%d)", B, T, G, H, CRUSE_MAX_GROUPS);
    SeqPtrs ptrs;
    for (int i = 0; i < CRUSE_MAX_GROUPS; ++i) { ptrs.w_hh[i] = nullptr; ptrs.b_hh[i] = nullptr; }
    for (int i = 0; i < G; ++i) {
        CRUSE_CHECK_ARG(w_hh[i], "gru_seq_fwd_tc: null weight pointer for group %d", i);
        ptrs.w_hh[i] = w_hh[i];
        ptrs.b_hh[i] = b_hh ? b_hh[i] : nullptr;
    }
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) launch_seq_tc_nc<N>(xproj, ptrs, h0, y, hT, gates, B, T, G, H, y_fs, y_gs, x_bs, x_ts, y_bs, y_ts, sync, st)
    SEQ_TC_DISPATCH((H + SQ_U - 1) / SQ_U, CALL)
#undef CALL
    set_error("gru_seq_fwd_tc: unsupported H=%d", H);
    return -1;
}

extern "C" int cruse_gru_seq_fwd_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h0,
                                    float* y, float* hT, float* gates, int B, int T, int G, int H, int y_fs, int y_gs,
                                    void* stream) {
    return gru_seq_tc_impl(xproj, w_hh, b_hh, h0, y, hT, gates, B, T, G, H, y_fs, y_gs, T, 1, T, 1, stream);
}

extern "C" int cruse_gru_seq_chunk_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h0,
                                      float* y, float* hT, int B, int Tc, int G, int H, int y_fs, int y_gs, long long x_bs,
                                      long long x_ts, long long y_bs, long long y_ts, void* stream) {
    CRUSE_CHECK_ARG(x_bs > 0 && x_ts > 0 && y_bs > 0 && y_ts > 0, "gru_seq_chunk_tc: row strides must be positive");
    return gru_seq_tc_impl(xproj, w_hh, b_hh, h0, y, hT, nullptr, B, Tc, G, H, y_fs, y_gs, x_bs, x_ts, y_bs, y_ts, stream);
}

// ---- the wavefront without relaunches: one launch per layer, chunk-level flags between the kernels ------------------
namespace cruse {
namespace {
__global__ void flag_wait_kernel(const unsigned* flag, unsigned target, int* err) {
    if (threadIdx.x == 0) spin_until(flag, target, err);
}
__global__ void flag_set_kernel(unsigned* flag, unsigned value) {
    __threadfence();
    atomicExch(flag, value);
}
// a bounded spin that timed out must not leave plausible numbers behind: every output of the step becomes NaN
struct PoisonBufs {
    float* p[4];
    long long n[4];
};
__global__ void __launch_bounds__(256) poison_on_error_kernel(const int* err, PoisonBufs b) {
    if (*reinterpret_cast<const volatile int*>(err) == 0) return;
    const float nan = __int_as_float(0x7fc00000);
    for (int k = 0; k < 4; ++k)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; b.p[k] && i < b.n[k]; i += (long long)gridDim.x * blockDim.x)
            b.p[k][i] = nan;
}
}  // namespace
}  // namespace cruse

extern "C" int cruse_flag_wait(const unsigned* flag, unsigned target, int* err, void* stream) {
    CRUSE_CHECK_ARG(flag, "flag_wait: null pointer");
    cruse::flag_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag, target, err);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_poison_on_error(const int* err, float* const* bufs, const long long* counts, int nbufs, void* stream) {
    CRUSE_CHECK_ARG(err && bufs && counts && nbufs >= 1 && nbufs <= 4, "poison_on_error: 1..4 buffers and an error flag needed");
    cruse::PoisonBufs b;
    for (int k = 0; k < 4; ++k) {
        b.p[k] = k < nbufs ? bufs[k] : nullptr;
        b.n[k] = k < nbufs ? counts[k] : 0;
    }
    cruse::poison_on_error_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(err, b);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_flag_set(unsigned* flag, unsigned value, void* stream) {
    CRUSE_CHECK_ARG(flag, "flag_set: null pointer");
    cruse::flag_set_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_gru_seq_flagged_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh, float* y, int B,
                                        int T, int G, int H, int y_fs, int y_gs, long long x_bs, long long x_ts, long long y_bs,
                                        long long y_ts, const int* bounds, int nchunks, const unsigned* wait_flags,
                                        unsigned wait_target, unsigned* done_flags, int* err, void* stream) {
    CRUSE_CHECK_ARG(bounds && nchunks >= 1 && nchunks <= 16, "gru_seq_flagged_tc: 1 <= nchunks <= 16 chunk bounds needed");
    CRUSE_CHECK_ARG(bounds[0] == 0 && bounds[nchunks] == T, "gru_seq_flagged_tc: bounds must run from 0 to T");
    for (int k = 0; k < nchunks; ++k)
        CRUSE_CHECK_ARG(bounds[k + 1] - bounds[k] >= 8, "gru_seq_flagged_tc: chunk %d has %d frames, at least 8 are needed", k, bounds[k + 1] - bounds[k]);
    SeqSync sync;
    memset(&sync, 0, sizeof(sync));
    sync.wait = wait_flags;
    sync.done = done_flags;
    sync.err = err;
    sync.wait_target = wait_target;
    sync.nchunks = nchunks;
    for (int k = 0; k <= nchunks; ++k) sync.bounds[k] = bounds[k];
    return gru_seq_tc_impl(xproj, w_hh, b_hh, nullptr, y, nullptr, nullptr, B, T, G, H, y_fs, y_gs, x_bs, x_ts, y_bs, y_ts, stream, &sync);
}

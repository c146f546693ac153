// gru_seq_tc.cu -- the GRU recurrence (the T sequential steps of nn.GRU, model/cruse_net.py:23-31,43-50
// of the reference; cuDNN's persistent RNN kernel on the reference path) on the 5th-generation tensor
// cores, one thread-block cluster per (group, 16-utterance slice).
//
// Per step the recurrent half of the cell is   pre[3H] = W_hh[g] . h_{t-1}   for every utterance: a
// [3H x H] . [H x 16] product.  The cluster splits the H hidden units 32 per CTA (H = 256 -> 8 CTAs);
// CTA c keeps the 96 gate rows {r,z,n} x {its 32 units} of W_hh resident in TENSOR MEMORY for the whole
// sequence as the A operand of ONE tcgen05.mma shape (M = 128 [96 used], N = 16 utterances, K = H,
// kind::tf32, fp32 accumulation in TMEM; lane = gate row, column = k, 256 of the 512 TMEM columns).
// (A first version kept the slice in shared memory: every step then re-streams 128 KB through the
// 128 B/clk shared-memory port, measured 47 clk per K=8 MMA; from TMEM the same MMA is not port-bound.)
// The B operand is h_{t-1} itself in shared memory ([16 utterances][H], K-major, 128-byte swizzle), double buffered.
// After the MMA the accumulator (lane = gate row, column = utterance) is pulled out with tcgen05.ld,
// transposed through a small shared-memory pad so that every thread owns (utterance, 2 adjacent
// units), the gates are evaluated in fp32 (the z*h_{t-1} carry uses the thread's own full-precision
// register copy, only the matmul operand is rounded to tf32), and the new 2 KB h slice -- which is
// exactly k-block `c` of everybody's next B operand -- is scattered to all CTAs of the cluster with
// st.async (8-byte DSMEM stores that credit the destination CTA's mbarrier).  No cluster barrier and
// no fence sits inside the time loop: step t+1's MMA is released by the byte count of h_t arriving.
//
// x-projections (+ folded biases) come from the tcgen05 input GEMM (gru_ih_tc.cu) and are prefetched
// one step ahead.  Optional `gates` output saves r, z, n and W_hn.h+b_hn for the backward pass.
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace cruse {
namespace {

constexpr int SQ_NB16 = 16;                  // utterances per cluster = MMA N = 16 * NI  (NI = 1 or 2)
constexpr int SQ_U = 32;                     // hidden units per CTA (= one 128-byte k-block of the B operand)
constexpr int SQ_THREADS = 256;              // thread -> (utterance tid/16, unit pair tid%16)
constexpr int SQ_TMEM_COLS = 512;            // A: up to 256 columns (K) + D: 16 columns -> whole TMEM (1 CTA/SM)
constexpr int SQ_D_COL = 256;                // accumulator column offset

#ifdef CRUSE_SEQ_TIMING
// developer instrumentation (never built into the shipped library): per-step clock64 stamps of cluster 0 / CTA 0
__device__ long long g_seq_timing[8 * 2048];
#define SEQ_STAMP(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && t < 2048 && (threadIdx.x & 31) == 0) g_seq_timing[t * 8 + (slot)] = clock64(); } while (0)
#else
#define SEQ_STAMP(slot) do {} while (0)
#endif

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
    const int kb = k >> 5, kk = k & 31;
    return (uint32_t)(kb * kb_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7))) << 4) + ((kk & 3) << 2));
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

template <int NC, int NI>
__global__ void __launch_bounds__(SQ_THREADS, 1)
gru_seq_tc_kernel(const float* __restrict__ xproj, const SeqPtrs ptrs, const float* __restrict__ h0, float* __restrict__ y,
                  float* __restrict__ hT, float* __restrict__ gates, int B, int T, int G, int H, int y_fs, int y_gs) {
    constexpr int NB = SQ_NB16 * NI;              // utterances per cluster = MMA N
    constexpr int H_KB = NB * 128;               // bytes of one B k-block tile: NB rows x 32 tf32
    constexpr int PRE_LD = NB + 1;               // padded leading dim of the gate-row x utterance pad
    extern __shared__ uint8_t smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int g = blockIdx.y, bslice = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sH = smem_raw + (base - tc::smem_u32(smem_raw));          // 2 buffers x NC k-blocks x H_KB
    float* sPre = reinterpret_cast<float*>(sH + 2 * NC * H_KB);        // [96][PRE_LD]
    uint64_t* hbar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sPre + 96 * PRE_LD) + 7) & ~(uintptr_t)7);
    uint64_t* acc_full = hbar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    // ---- thread -> (utterances b + 16*i, units u0, u0+1)
    const int b = tid >> 4, jp = tid & 15;
    const int u0 = rank * SQ_U + 2 * jp;
    const bool uvalid = u0 < H;                     // H is even, so u0+1 < H as well
    int bg[NI];
    bool valid[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        bg[i] = bslice * NB + b + 16 * i;
        valid[i] = uvalid && bg[i] < B;
    }

    for (int i = tid; i < 2 * NC * H_KB / 16; i += SQ_THREADS) reinterpret_cast<float4*>(sH)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        tc::mbar_init(&hbar[0], 1);
        tc::mbar_init(&hbar[1], 1);
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 3) tc::tmem_alloc<SQ_TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // ---- W_hh slice -> tensor memory (A operand): lane q*32+j = gate q of hidden unit rank*32+j, column = k, tf32.
    //      A warp can only touch its own lane quadrant (warp % 4); warps 0-3 / 4-7 split the k-blocks.
    {
        const float* W = ptrs.w_hh[g];
        const int q = warp & 3, unit = rank * SQ_U + lane;
        const bool rowvalid = q < 3 && unit < H;
        const float* wrow = W + ((size_t)(rowvalid ? q : 0) * H + (rowvalid ? unit : 0)) * H;
        for (int kb = (warp >> 2); kb < NC; kb += 2) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = kb * 32 + i * 4;
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rowvalid && k < H) f = __ldg(reinterpret_cast<const float4*>(wrow + k));
                v[i * 4 + 0] = to_tf32(f.x); v[i * 4 + 1] = to_tf32(f.y); v[i * 4 + 2] = to_tf32(f.z); v[i * 4 + 3] = to_tf32(f.w);
            }
            tmem_st_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kb * 32), v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // ---- h_0 -> hbuf[0] (tf32 operand copy) and the thread's own full-precision copy
    float2 hold[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) hold[i] = make_float2(0.f, 0.f);
    if (h0) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
            if (valid[i]) hold[i] = __ldg(reinterpret_cast<const float2*>(h0 + ((size_t)g * B + bg[i]) * H + u0));
        for (int i = tid; i < NB * (H / 2); i += SQ_THREADS) {      // every CTA needs the whole h_0 of its utterances
            const int bb = i / (H / 2), k = (i - bb * (H / 2)) * 2;
            const int bgg = bslice * NB + bb;
            if (bgg < B) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(h0 + ((size_t)g * B + bgg) * H + k));
                *reinterpret_cast<float2*>(sH + sw128_off(bb, k, H_KB)) = make_float2(to_tf32(v.x), to_tf32(v.y));
            }
        }
    }
    tc::fence_proxy_async_smem();      // generic-proxy smem writes (h_0) -> visible to the tensor core's async proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_base + SQ_D_COL;
    cluster.sync();                    // every CTA's buffers + barriers exist before any remote store lands

    // ---- remote (shared::cluster) addresses of my 8-byte h slot and of the barriers in every CTA of the cluster
    uint32_t rem_h[NC], rem_bar[NC];
    {
        const uint32_t lh = tc::smem_u32(sH) + (uint32_t)rank * H_KB + sw128_off(b, 2 * jp, H_KB);
        const uint32_t lb = tc::smem_u32(&hbar[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_h[c]) : "r"(lh), "r"(c));
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_bar[c]) : "r"(lb), "r"(c));
        }
    }
    constexpr uint32_t STEP_BYTES = (uint32_t)NC * SQ_THREADS * 8 * NI;   // all of h_t: NB x 32*NC floats

    const size_t N3 = (size_t)3 * H;
    const size_t xstep = (size_t)G * N3;
    const size_t ystep = (size_t)G * H;
    const float* xp[NI];
    float* yp[NI];
    float* gp[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const size_t bt = (size_t)(valid[i] ? bg[i] : 0) * T;
        xp[i] = xproj + (bt * G + g) * N3 + (uvalid ? u0 : 0);
        yp[i] = y + bt * ystep + (size_t)(uvalid ? u0 : 0) * y_fs + (size_t)g * y_gs;
        gp[i] = gates ? gates + (bt * G + g) * (4 * (size_t)H) + (uvalid ? u0 : 0) : nullptr;
    }
    float2 b_hn = make_float2(0.f, 0.f);
    if (uvalid && ptrs.b_hh[g]) b_hn = __ldg(reinterpret_cast<const float2*>(ptrs.b_hh[g] + 2 * H + u0));

    // x-projections are streamed from HBM two steps ahead of their use (a step is shorter than a DRAM round trip)
    float2 xr[NI], xz[NI], xn[NI], x1r[NI], x1z[NI], x1n[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        xr[i] = xz[i] = xn[i] = x1r[i] = x1z[i] = x1n[i] = make_float2(0.f, 0.f);
        if (valid[i] && T > 0) {
            xr[i] = __ldg(reinterpret_cast<const float2*>(xp[i]));
            xz[i] = __ldg(reinterpret_cast<const float2*>(xp[i] + H));
            xn[i] = __ldg(reinterpret_cast<const float2*>(xp[i] + 2 * H));
        }
        if (valid[i] && T > 1) {
            x1r[i] = __ldg(reinterpret_cast<const float2*>(xp[i] + xstep));
            x1z[i] = __ldg(reinterpret_cast<const float2*>(xp[i] + xstep + H));
            x1n[i] = __ldg(reinterpret_cast<const float2*>(xp[i] + xstep + 2 * H));
        }
    }
    constexpr uint32_t idesc = tc::instr_desc(2 /*tf32*/, 128, NB);

    int p = 0;
    uint32_t ph0 = 0, ph1 = 0;
    for (int t = 0; t < T; ++t) {
        if (tid == 0) tc::mbar_expect_tx(&hbar[p ^ 1], STEP_BYTES);   // arm the barrier h_{t+1} will complete
        if (warp == 3) {
            // ===== MMA issue: pre = W_slice . h_t.  The whole warp waits (converged), one elected lane issues =====
            SEQ_STAMP(0);
            if (t > 0) {
                if (p) { tc::mbar_wait(&hbar[1], ph1); ph1 ^= 1; } else { tc::mbar_wait(&hbar[0], ph0); ph0 ^= 1; }
                tc::fence_proxy_async_smem();          // h_t was written by (remote) generic-proxy stores
            }
            tc::tc_fence_after();
            SEQ_STAMP(1);
            const uint64_t bdesc0 = tc::smem_desc_sw128(base + (uint32_t)p * (NC * H_KB));
            if (tc::elect_one()) {
#pragma unroll
                for (int kb = 0; kb < NC; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32_ts(tmem_d, tmem_base + (uint32_t)(kb * 32 + ks * 8), bdesc0 + (uint64_t)((kb * H_KB + ks * 32) >> 4),
                                     idesc, (kb | ks) ? 1u : 0u);
                tc::umma_commit(acc_full);
            }
            __syncwarp();
            SEQ_STAMP(2);
        }
        // prefetch the x-projections of step t+2 while the tensor core works
        float2 nxr[NI], nxz[NI], nxn[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            nxr[i] = nxz[i] = nxn[i] = make_float2(0.f, 0.f);
            if (valid[i] && t + 2 < T) {
                const float* q = xp[i] + (size_t)(t + 2) * xstep;
                nxr[i] = __ldg(reinterpret_cast<const float2*>(q));
                nxz[i] = __ldg(reinterpret_cast<const float2*>(q + H));
                nxn[i] = __ldg(reinterpret_cast<const float2*>(q + 2 * H));
            }
        }
        tc::mbar_wait(acc_full, (uint32_t)(t & 1));
        tc::tc_fence_after();
        if (tid == 0) SEQ_STAMP(3);
        if (warp < 3) {
            // accumulator lanes 32*warp.. = gate `warp` of my CTA's 32 units; columns = utterances
            float* dst = sPre + (warp * 32 + lane) * PRE_LD;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                float v[16];
                tmem_ld_32x16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(16 * i), v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) dst[16 * i + c] = v[c];
            }
        }
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) SEQ_STAMP(4);
        float2 hnew[NI], gr[NI], gz[NI], gn[NI], ghn[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int bb = b + 16 * i;
            const float pr0 = sPre[(0 * 32 + 2 * jp) * PRE_LD + bb], pr1 = sPre[(0 * 32 + 2 * jp + 1) * PRE_LD + bb];
            const float pz0 = sPre[(1 * 32 + 2 * jp) * PRE_LD + bb], pz1 = sPre[(1 * 32 + 2 * jp + 1) * PRE_LD + bb];
            ghn[i].x = sPre[(2 * 32 + 2 * jp) * PRE_LD + bb] + b_hn.x;
            ghn[i].y = sPre[(2 * 32 + 2 * jp + 1) * PRE_LD + bb] + b_hn.y;
            gr[i].x = sigmoid_fast(xr[i].x + pr0); gr[i].y = sigmoid_fast(xr[i].y + pr1);
            gz[i].x = sigmoid_fast(xz[i].x + pz0); gz[i].y = sigmoid_fast(xz[i].y + pz1);
            gn[i].x = tanh_fast(xn[i].x + gr[i].x * ghn[i].x); gn[i].y = tanh_fast(xn[i].y + gr[i].y * ghn[i].y);
            hnew[i].x = valid[i] ? ((1.f - gz[i].x) * gn[i].x + gz[i].x * hold[i].x) : 0.f;
            hnew[i].y = valid[i] ? ((1.f - gz[i].y) * gn[i].y + gz[i].y * hold[i].y) : 0.f;
            hold[i] = hnew[i];
        }
        if (tid == 0) SEQ_STAMP(7);
        // scatter my 8 bytes per utterance of h_{t+1} (tf32-rounded operand copy) into the other buffer of every CTA
        {
            const uint32_t poff = (uint32_t)(p ^ 1) * (NC * H_KB), boff = (uint32_t)(p ^ 1) * 8;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const uint32_t w0 = __float_as_uint(to_tf32(hnew[i].x)), w1 = __float_as_uint(to_tf32(hnew[i].y));
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(
                                     rem_h[c] + poff + (uint32_t)(i * 2048)),      // utterance b+16 = two 8-row groups further
                                 "r"(w0), "r"(w1), "r"(rem_bar[c] + boff)
                                 : "memory");
            }
        }
        if (tid == 0) SEQ_STAMP(5);
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            if (!valid[i]) continue;
            float* yo = yp[i] + (size_t)t * ystep;
            if (y_fs == 1) {
                *reinterpret_cast<float2*>(yo) = hnew[i];
            } else {
                yo[0] = hnew[i].x;
                yo[y_fs] = hnew[i].y;
            }
            if (gp[i]) {
                float* go = gp[i] + (size_t)t * G * 4 * H;
                *reinterpret_cast<float2*>(go) = gr[i];
                *reinterpret_cast<float2*>(go + H) = gz[i];
                *reinterpret_cast<float2*>(go + 2 * H) = gn[i];
                *reinterpret_cast<float2*>(go + 3 * H) = ghn[i];
            }
        }
        if (tid == 0) SEQ_STAMP(6);
        p ^= 1;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            xr[i] = x1r[i]; xz[i] = x1z[i]; xn[i] = x1n[i];
            x1r[i] = nxr[i]; x1z[i] = nxz[i]; x1n[i] = nxn[i];
        }
    }
    if (hT) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
            if (valid[i]) *reinterpret_cast<float2*>(hT + ((size_t)g * B + bg[i]) * H + u0) = hold[i];
    }
    // drain: the last step's scatter is still in flight towards my buffers -- wait for it before anybody exits
    if (T > 0 && warp == 3) {
        if (p) tc::mbar_wait(&hbar[1], ph1); else tc::mbar_wait(&hbar[0], ph0);
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster.sync();
    if (warp == 3) tc::tmem_dealloc<SQ_TMEM_COLS>(tmem_base);
}

constexpr size_t seq_smem_bytes(int NC, int NI) {
    return 1024 + 2 * (size_t)NC * (SQ_NB16 * NI * 128) + (96 * (SQ_NB16 * NI + 1) + 4) * 4 + 64;
}

template <int NC, int NI>
void seq_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int G, int nslices, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(NC, G, nslices);
    cfg.blockDim = dim3(SQ_THREADS);
    cfg.dynamicSmemBytes = seq_smem_bytes(NC, NI);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

template <int NC, int NI>
int max_clusters_tc() {
    if (cudaFuncSetAttribute(gru_seq_tc_kernel<NC, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes(NC, NI)) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    seq_cfg<NC, NI>(cfg, attr, 1, 1024, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gru_seq_tc_kernel<NC, NI>, &cfg) != cudaSuccess) return -2;
    return n;
}

template <int NC, int NI>
int launch_seq_tc(const float* xproj, const SeqPtrs& ptrs, const float* h0, float* y, float* hT, float* gates, int B, int T,
                  int G, int H, int y_fs, int y_gs, cudaStream_t st) {
    CRUSE_CUDA_OK(cudaFuncSetAttribute(gru_seq_tc_kernel<NC, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem_bytes(NC, NI)));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    seq_cfg<NC, NI>(cfg, attr, G, (B + SQ_NB16 * NI - 1) / (SQ_NB16 * NI), st);
    CRUSE_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_seq_tc_kernel<NC, NI>, xproj, ptrs, h0, y, hT, gates, B, T, G, H, y_fs, y_gs));
    return 0;
}

template <int NC>
int launch_seq_tc_nc(const float* xproj, const SeqPtrs& ptrs, const float* h0, float* y, float* hT, float* gates, int B, int T,
                     int G, int H, int y_fs, int y_gs, cudaStream_t st) {
    // 16 utterances per cluster while all clusters are co-resident (shortest step); 32 when that avoids a second wave
    static thread_local int cap16 = -1;
    if (cap16 < 0) cap16 = max_clusters_tc<NC, 1>();
    const int need16 = G * ((B + 15) / 16);
    if (cap16 > 0 && need16 > cap16) return launch_seq_tc<NC, 2>(xproj, ptrs, h0, y, hT, gates, B, T, G, H, y_fs, y_gs, st);
    return launch_seq_tc<NC, 1>(xproj, ptrs, h0, y, hT, gates, B, T, G, H, y_fs, y_gs, st);
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
#define CALL(N) max_clusters_tc<N, 1>()
    SEQ_TC_DISPATCH((H + SQ_U - 1) / SQ_U, CALL)
#undef CALL
    set_error("gru_seq_tc_max_clusters: unsupported H=%d", H);
    return -1;
}

extern "C" int cruse_gru_seq_fwd_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h0,
                                    float* y, float* hT, float* gates, int B, int T, int G, int H, int y_fs, int y_gs,
                                    void* stream) {
    CRUSE_CHECK_ARG(xproj && y && w_hh, "gru_seq_fwd_tc: null pointer");
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
#define CALL(N) launch_seq_tc_nc<N>(xproj, ptrs, h0, y, hT, gates, B, T, G, H, y_fs, y_gs, st)
    SEQ_TC_DISPATCH((H + SQ_U - 1) / SQ_U, CALL)
#undef CALL
    set_error("gru_seq_fwd_tc: unsupported H=%d", H);
    return -1;
}

// conv_wgrad_tc.cu -- weight (and bias) gradients of the encoder / skip convolutions as a split-K GEMM on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in tensor memory).
//
//   dW[co, (ci,kt,kf)] = sum over positions p = (b,t,fo) of  dz[p, co] * x[b, t-(KT-1)+kt, ci, SF*fo-1+kf]
//   dbias[co]          = sum over positions of dz[p, co]
//
// (on the reference path these are the cuDNN wgrad calls autograd makes for
// model/cruse_net.py:138-143,149-156.)  The reduction runs over the POSITIONS, and both operands are contiguous
// along positions in the frame-major layout ([b,t][c][f]: for a fixed channel the bins of a frame are adjacent),
// so both are K-major as they lie in HBM:
//   A[M = co][K = 32 positions]   = 32 consecutive (frame, bin) values of one dz channel,
//   B[N = (ci,kt,kf)][K = 32 pos] = the same 32 positions of one input channel, shifted by the tap
//                                   (+ one constant row of ones: its accumulator column is dbias).
// A CTA walks k-blocks of 32 positions (grid-stride): producer warps load the rows with coalesced loads (lane =
// position; the three frequency taps of a row are the lane's own values plus one shuffle from the neighbouring bin),
// round to tf32 and store them as 128-byte rows of the SWIZZLE_128B layout (one row per store instruction: 32 distinct
// banks); one warp issues 4 tcgen05.mma (M128 x N x K8) per k-block into ONE accumulator that lives in TMEM for the
// whole launch; at the end the CTA writes its partial [Cout x Cin*KT*3 | Cout] and cruse_conv_wgrad sums the
// partials in a fixed order (deterministic, no atomics).  Replaces the CUDA-core kernel of conv_bwd.cu (0.2-0.9 ms
// per stage) when the conv mode is tf32.
//
// MODE 1 is the transposed conv of the decoder, dW[ci, (co,k)] = sum_p x[p, ci] * dz[b,t,co, 2i+k]: the plain operand
// (A, lanes) is the input x, the tapped operand (B, columns) is dz read at bins 2i, 2i+1, 2i+2 (cropped), and the row of
// ones sits in A: its accumulator lane holds sum_p dz[.., 2i+k], i.e. dbias[co] = columns (co,0) + (co,1).
#include "common.cuh"
#include "tc_common.cuh"

namespace cruse {
int conv_max_ctas();     // conv_tc.cu: optional cap on persistent grids (0 = none)
namespace {

constexpr int WT_NPW = 8;                          // producer warps per set
constexpr int WT_SETS = 2;                         // sets alternate over k-blocks
constexpr int WT_PROD_WARPS = WT_NPW * WT_SETS;
constexpr int WT_MMA_WARP = WT_PROD_WARPS;
constexpr int WT_THREADS = (WT_PROD_WARPS + 1) * 32;   // 544
constexpr int WT_STAGES = 4;
constexpr int WT_TMEM_COLS = 256;

// MODE 0: conv (KT x 3, stride SF, pad 1): plain operand A = dz (CP = Cout channels), tapped operand B = x (CQ = Cin).
// MODE 1: convT (1 x 3, stride 2):          plain operand A = x (CP = Cin channels),  tapped operand B = dz (CQ = Cout).
// FO = bins per frame of the POSITIONS (= of the plain operand); the tapped operand has TSF * FO bins.
template <int MODE, int KT, int SF, int CP, int CQ, int FO>
struct WgradCfg {
    static constexpr int NTAP = KT * 3;
    static constexpr int NW = CQ * NTAP;                           // accumulator columns that are weights
    static constexpr int NWP = (NW + 15) / 16 * 16;
    static constexpr int N = MODE == 0 ? NWP + 16 : NWP;           // conv: + the bias block (row NWP of B = ones)
    static constexpr int TSF = MODE == 0 ? SF : 2;                 // bins of the tapped operand per position
    static constexpr int FQ = TSF * FO;
    static constexpr int A_BYTES = 128 * 128;                      // 128 rows (zero above CP [+1]) x 32 positions
    static constexpr int B_BYTES = N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM = 1024 + WT_STAGES * STAGE_BYTES + 256;
    static constexpr int NA = CP / WT_NPW;                         // plain rows per producer warp and k-block
    static constexpr int NC = CQ / WT_NPW;                         // tapped channels per producer warp and k-block
    static_assert(CP % WT_NPW == 0 && CQ % WT_NPW == 0, "channels must split over the producer warps");
    static_assert(N <= WT_TMEM_COLS && N % 16 == 0 && CP + 1 <= 128, "accumulator shape");
    static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
    static_assert(FO % 16 == 0 && (MODE == 0 || KT == 1), "geometry");
};

__device__ __forceinline__ uint32_t to_tf32_bits(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// byte offset of (row r, position k) inside a K-major SWIZZLE_128B tile of 128-byte rows
__device__ __forceinline__ uint32_t row_off(int r, int k) { return (uint32_t)(r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + ((k & 3) << 2)); }

template <int MODE, int KT, int SF, int CP, int CQ, int FO>
__global__ void __launch_bounds__(WT_THREADS, 1)
conv_wgrad_tc_kernel(const float* __restrict__ pl, const float* __restrict__ tp, float* __restrict__ ws, int B, int T, int pitch) {
    // pl = plain operand [B,T,CP,FO] (conv: dz, convT: x);  tp = tapped operand [B,T,CQ,FQ] (conv: x, convT: dz)
    using C = WgradCfg<MODE, KT, SF, CP, CQ, FO>;
    constexpr int TS = C::TSF;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* stages = smem_raw + (base - tc::smem_u32(smem_raw));
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + WT_STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + WT_STAGES;
    uint64_t* acc_full = empty + WT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long NP = (long long)B * T * FO;                    // positions
    const int nkb = (int)((NP + 31) / 32);
    const int my_nkb = (nkb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < WT_STAGES; ++i) { tc::mbar_init(&full[i], WT_NPW * 32); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == WT_MMA_WARP) tc::tmem_alloc<WT_TMEM_COLS>(tmem_slot);
    // constant parts of every stage: zero rows (dz rows >= COUT, weight-column padding), the row of ones behind the bias column
    for (int i = tid; i < WT_STAGES * C::STAGE_BYTES / 16; i += WT_THREADS) reinterpret_cast<float4*>(stages)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int i = tid; i < WT_STAGES * 32; i += WT_THREADS) {
        const int s = i >> 5, k = i & 31;
        if (MODE == 0) *reinterpret_cast<float*>(stages + s * C::STAGE_BYTES + C::A_BYTES + row_off(C::NWP, k)) = 1.0f;
        else *reinterpret_cast<float*>(stages + s * C::STAGE_BYTES + row_off(CP, k)) = 1.0f;
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp < WT_PROD_WARPS) {
        // ================= producers =================
        const int set = warp / WT_NPW, wq = warp % WT_NPW;
        float av[C::NA];                                    // plain rows c = wq + 8*i
        float xv[C::NC][KT][TS];                            // tapped channel c = wq + 8*i, time tap kt: bins TS*f (, TS*f+1)
        float xe[C::NC][KT][2];                             // neighbours that are not in the warp: [0] left of lane 0, [1] right of lane 31
        int fo_cur = 0;

        auto load_kb = [&](int j) {                         // j = local k-block index of this CTA
            const long long P = ((long long)blockIdx.x + (long long)j * gridDim.x) * 32 + lane;
            const bool pv = P < NP;
            const long long g = pv ? P / FO : 0;            // frame record b*T + t
            const int fo = pv ? (int)(P - g * FO) : 0;
            const int t = (int)(g % T);
            fo_cur = fo;
            const float* pp = pl + (size_t)g * (CP * FO) + fo;
#pragma unroll
            for (int i = 0; i < C::NA; ++i) av[i] = pv ? __ldg(pp + (size_t)(wq + WT_NPW * i) * FO) : 0.f;
#pragma unroll
            for (int i = 0; i < C::NC; ++i) {
                const int cq = wq + WT_NPW * i;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const bool fv = pv && (t - (KT - 1) + kt >= 0);          // the frame of this time tap exists
                    const float* xp = tp + ((size_t)(g - (KT - 1) + kt) * CQ + cq) * C::FQ + TS * fo;
                    if (TS == 2) {
                        const float2 v = fv ? __ldg(reinterpret_cast<const float2*>(xp)) : make_float2(0.f, 0.f);
                        xv[i][kt][0] = v.x; xv[i][kt][TS - 1] = v.y;
                        xe[i][kt][0] = (MODE == 0 && fv && lane == 0 && fo > 0) ? __ldg(xp - 1) : 0.f;
                        xe[i][kt][1] = (MODE == 1 && fv && lane == 31 && fo < FO - 1) ? __ldg(xp + 2) : 0.f;
                    } else {
                        xv[i][kt][0] = fv ? __ldg(xp) : 0.f;
                        xe[i][kt][0] = (fv && lane == 0 && fo > 0) ? __ldg(xp - 1) : 0.f;
                        xe[i][kt][1] = (fv && lane == 31 && fo < FO - 1) ? __ldg(xp + 1) : 0.f;
                    }
                }
            }
        };
        auto store_kb = [&](uint8_t* st) {
            uint8_t* sA = st;
            uint8_t* sB = st + C::A_BYTES;
            const int fo = fo_cur;
#pragma unroll
            for (int i = 0; i < C::NA; ++i)
                *reinterpret_cast<uint32_t*>(sA + row_off(wq + WT_NPW * i, lane)) = to_tf32_bits(av[i]);
#pragma unroll
            for (int i = 0; i < C::NC; ++i) {
                const int cq = wq + WT_NPW * i;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    float t0, t1, t2;                                         // taps 0, 1, 2
                    if (MODE == 1) {                                          // convT: dz bins 2i, 2i+1, 2i+2 (the last one cropped)
                        float r = __shfl_down_sync(0xffffffffu, xv[i][kt][0], 1);
                        if (lane == 31) r = xe[i][kt][1];
                        if (fo == FO - 1) r = 0.f;
                        t0 = xv[i][kt][0]; t1 = xv[i][kt][TS - 1]; t2 = r;
                    } else if (TS == 2) {                                     // conv stride 2: bins 2fo-1, 2fo, 2fo+1
                        float l = __shfl_up_sync(0xffffffffu, xv[i][kt][TS - 1], 1);
                        if (lane == 0) l = xe[i][kt][0];
                        if (fo == 0) l = 0.f;
                        t0 = l; t1 = xv[i][kt][0]; t2 = xv[i][kt][TS - 1];
                    } else {                                                  // conv stride 1: bins fo-1, fo, fo+1
                        float l = __shfl_up_sync(0xffffffffu, xv[i][kt][0], 1);
                        float r = __shfl_down_sync(0xffffffffu, xv[i][kt][0], 1);
                        if (lane == 0) l = xe[i][kt][0];
                        if (lane == 31) r = xe[i][kt][1];
                        if (fo == 0) l = 0.f;
                        if (fo == FO - 1) r = 0.f;
                        t0 = l; t1 = xv[i][kt][0]; t2 = r;
                    }
                    const int n0 = (cq * KT + kt) * 3;
                    *reinterpret_cast<uint32_t*>(sB + row_off(n0 + 0, lane)) = to_tf32_bits(t0);
                    *reinterpret_cast<uint32_t*>(sB + row_off(n0 + 1, lane)) = to_tf32_bits(t1);
                    *reinterpret_cast<uint32_t*>(sB + row_off(n0 + 2, lane)) = to_tf32_bits(t2);
                }
            }
        };

        if (set < my_nkb) load_kb(set);
#pragma unroll 1
        for (int j = set; j < my_nkb; j += WT_SETS) {
            const int s = j % WT_STAGES;
            tc::mbar_wait_backoff(&empty[s], ((j / WT_STAGES) & 1) ^ 1, 64);
            store_kb(stages + s * C::STAGE_BYTES);
            // (the shuffles above read the registers of THIS k-block: its loads were issued one iteration ago)
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&full[s]);
            if (j + WT_SETS < my_nkb) load_kb(j + WT_SETS);
        }
    } else {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = tc::instr_desc(2 /*tf32*/, 128, C::N);
        for (int j = 0; j < my_nkb; ++j) {
            const int s = j % WT_STAGES;
            tc::mbar_wait(&full[s], (j / WT_STAGES) & 1);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t sa = base + s * C::STAGE_BYTES, sb = sa + C::A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_tf32(tmem_d, tc::smem_desc_sw128(sa + k * 32), tc::smem_desc_sw128(sb + k * 32), idesc, (j | k) ? 1u : 0u);
                tc::umma_commit(&empty[s]);
            }
            __syncwarp();
        }
        if (tc::elect_one()) tc::umma_commit(acc_full);
        __syncwarp();
    }

    // ================= epilogue: the CTA's partial sums -> ws[cta][COUT x NW | COUT] =================
    if (warp < 4) {
        float* o = ws + (size_t)blockIdx.x * pitch;
        const int co = warp * 32 + lane;
        if (MODE == 1 && tid < CQ) o[(size_t)CP * C::NW + tid] = 0.f;     // bias partial is accumulated from two columns below
        if (MODE == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
        if (my_nkb > 0) {
            tc::mbar_wait(acc_full, 0);
            tc::tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < C::N; c0 += 16) {
            float v[16];
            if (my_nkb > 0) {
                uint32_t* r = reinterpret_cast<uint32_t*>(v);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0)
                    : "memory");
                tc::tmem_ld_wait();
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = 0.f;
            }
            if (MODE == 0) {
                if (co < CP) {                                   // lane = co; columns = (ci,kt,kf) | bias
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int n = c0 + q;
                        if (n < C::NW) o[(size_t)co * C::NW + n] = v[q];
                        else if (n == C::NWP) o[(size_t)CP * C::NW + co] = v[q];
                    }
                }
            } else {
                if (co < CP) {                                   // lane = ci; columns = (co,k)
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (c0 + q < C::NW) o[(size_t)co * C::NW + c0 + q] = v[q];
                } else if (co == CP) {                           // the lane of ones: sum_p dz[co, 2i+k]  ->  dbias = k=0 + k=1
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int n = c0 + q;
                        if (n < C::NW && (n % 3) != 2) atomicAdd(o + (size_t)CP * C::NW + n / 3, v[q]);
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == WT_MMA_WARP) tc::tmem_dealloc<WT_TMEM_COLS>(tmem_d);
}

template <int MODE, int KT, int SF, int CP, int CQ, int FO>
int launch_wgrad_tc(const float* pl, const float* tp, float* ws, int B, int T, int pitch, int max_grid, cudaStream_t st) {
    using C = WgradCfg<MODE, KT, SF, CP, CQ, FO>;
    auto kern = conv_wgrad_tc_kernel<MODE, KT, SF, CP, CQ, FO>;
    static bool attr_set = false;
    if (!attr_set) {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    const long long nkb = ((long long)B * T * FO + 31) / 32;
    int grid = sm_count();
    if (grid > max_grid) grid = max_grid;
    if (conv_max_ctas() > 0 && grid > conv_max_ctas()) grid = conv_max_ctas();     // running beside the BPTT on a side stream
    if ((long long)grid > nkb) grid = (int)nkb;
    kern<<<grid, WT_THREADS, C::SMEM, st>>>(pl, tp, ws, B, T, pitch);
    CRUSE_LAUNCH_OK();
    return grid;
}

}  // namespace

// Returns the number of partials written (> 0) when the stage ran on the tensor cores, 0 when no instantiation matches,
// < 0 on error.  Partial p lies at ws + p * pitch: [Cout x Cin*kt*3 weights | Cout bias sums] (the layout of conv_bwd.cu).
int conv_wgrad_tc_try(const float* x, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt,
                      int fstride, int pitch, int max_grid, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) & 7) != 0) return 0;
#define CRUSE_WT(KT_, SF_, CI_, CO_, FO_) \
    if (kt == KT_ && fstride == SF_ && Cin == CI_ && Cout == CO_ && Fout == FO_ && Fin == SF_ * FO_) \
        return launch_wgrad_tc<0, KT_, SF_, CO_, CI_, FO_>(dz, x, ws, B, T, pitch, max_grid, st);
    CRUSE_WT(2, 2, 8, 16, 64)
    CRUSE_WT(2, 2, 16, 32, 32)
    CRUSE_WT(2, 2, 32, 64, 16)
    CRUSE_WT(1, 1, 8, 8, 128)
    CRUSE_WT(1, 1, 16, 16, 64)
    CRUSE_WT(1, 1, 32, 32, 32)
    CRUSE_WT(1, 1, 64, 64, 16)
#undef CRUSE_WT
    return 0;
}

// the same for the transposed conv (partial layout [Cin x Cout*3 weights | Cout bias sums])
int convT_wgrad_tc_try(const float* x, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int pitch,
                       int max_grid, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(dz) & 7) != 0) return 0;
#define CRUSE_WTT(CI_, CO_, FI_) \
    if (Cin == CI_ && Cout == CO_ && Fin == FI_ && Fout == 2 * FI_) \
        return launch_wgrad_tc<1, 1, 2, CI_, CO_, FI_>(x, dz, ws, B, T, pitch, max_grid, st);
    CRUSE_WTT(64, 32, 16)
    CRUSE_WTT(32, 16, 32)
    CRUSE_WTT(16, 8, 64)
#undef CRUSE_WTT
    return 0;
}

}  // namespace cruse

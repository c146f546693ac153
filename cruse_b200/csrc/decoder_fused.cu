// decoder_fused.cu -- LayerNorm 2 (+ skip 4) and the whole 4-stage decoder of the 256-bin pyramid in ONE kernel (eval mode).
//
// replaces, per frame range: model/cruse_net.py:51 (ln2) + :160 (+ skip_connect_4) and :161-164 repaired
// (3 x [ConvTranspose2d(1,3)/s(1,2) + BatchNorm(eval) + act + skip], ConvTranspose2d(8->1) + sigmoid) -- five launches and
// three intermediate tensors (3 x 65.7 MB at 32 x 501 frames, written and read back) in the staged schedule.
//
// The decoder has NO time taps: every frame is an independent chain 1024 -> [32x32] -> [16x64] -> [8x128] -> 256.  One WARP owns one
// frame from the GRU output to the mask; the activations never leave the SM:
//   * prologue: LayerNorm over the 1024 features of the frame (the row lives in registers), + skip 4, written as the stage-4 operand
//     X4[pos][channel] into the warp's shared-memory buffer (tf32-rounded, cvt.rna);
//   * stages 4..2 as warp-level tensor-core GEMMs (mma.sync m16n8k8 tf32, fp32 accumulate): M = the frame's input bins, K = Cin,
//     N = Cout; out[co, 2i] = W[:,co,0].x[:,i] + W[:,co,2].x[:,i-1] and out[co, 2i+1] = W[:,co,1].x[:,i] -- the i-1 tap is the same
//     operand buffer read one row up (row 0 is a zero row), so no im2col and no cross-lane exchange.  Epilogue in registers: folded
//     BN + bias, ReLU / PReLU, + skip (float2 loads issued BEFORE the k loop), stored as the next stage's operand;
//   * stage 1 (8 -> 1 channel, sigmoid) on the FMA pipe (N = 1 is no GEMM), mask stored with 256-byte warp transactions.
// A frame is 16..128 rows -- below one tcgen05 tile -- and the chain is bound by the 20 KB per frame it reads (GRU output + four
// skip tensors), not by the 0.35 MFLOP it computes, so the warp-private mma.sync pipeline (no barriers between warps, 16 frames in
// flight per SM) is the right tool here; the tcgen05 kernels stay where M is large (conv_tc.cu, gru_*_tc.cu).
// Weights of all four stages + the folded epilogue constants (35 KB, tf32-rounded, K-contiguous per output channel) are laid out
// ONCE per forward pass by decoder_fused_prep_kernel as the exact shared-memory image; a CTA stages it with one bulk copy
// (cp.async.bulk global -> shared, completion on an mbarrier) that runs under the LayerNorm of the CTA's first frames.
#include "common.cuh"
#include "tc_common.cuh"

namespace cruse {
namespace {

constexpr int DF_WARPS = 16, DF_THREADS = DF_WARPS * 32;      // SKC = false
constexpr int DF_WARPS_SKC = 12;                              // SKC = true: 64 KB more constants, 12 frames in flight per SM
constexpr int C4 = 64, F4 = 16, C3 = 32, F3 = 32, C2 = 16, F2 = 64, C1 = 8, F1 = 128, F0 = 256;
constexpr int LD4 = C4 + 4, LD3 = C3 + 4, LD2 = C2 + 4;     // operand rows: [1 + pos][channel], stride = Cin + 4 floats (conflict-free fragments)
constexpr int LD1 = F1 + 4;                                  // stage-1 input is channel-major [8][2 + pos]
constexpr int W4_OFF = 0, W4_SZ = 3 * C3 * LD4;              // [tap][co][ci + 4]
constexpr int W3_OFF = W4_OFF + W4_SZ, W3_SZ = 3 * C2 * LD3;
constexpr int W2_OFF = W3_OFF + W3_SZ, W2_SZ = 3 * C1 * LD2;
constexpr int W1_OFF = W2_OFF + W2_SZ, W1_SZ = 32;           // [ci][tap] (24) + bias
constexpr int EP_OFF = W1_OFF + W1_SZ, EP_N = C3 + C2 + C1;  // per output channel of stages 4..2: scale, shift (bias folded in), alpha
constexpr int WTOT = EP_OFF + 3 * EP_N + 8;
constexpr int BUFA = 1312;                                   // X4 (17 x 68 = 1156) then X2 (65 x 20 = 1300)
constexpr int BUFB = 1232;                                   // [E4 (18 x 68 = 1224)] X3 (33 x 36 = 1188) then X1 (8 x 132 = 1056)
// SKC: skip convs 4 and 3 (Conv2d(1,3), padding 1, no bias; model/cruse_net.py:143,153-156) computed here from the encoder outputs
constexpr int WS4_OFF = WTOT, WS4_SZ = 3 * C4 * LD4;         // [tap][co][ci + 4]
constexpr int WS3_OFF = WS4_OFF + WS4_SZ, WS3_SZ = 3 * C3 * LD3;
constexpr int WTOT_SKC = WS3_OFF + WS3_SZ;
template <bool SKC> constexpr int image_floats() { return SKC ? WTOT_SKC : WTOT; }
template <bool SKC> constexpr int df_warps() { return SKC ? DF_WARPS_SKC : DF_WARPS; }
template <bool SKC> constexpr size_t df_smem() { return (size_t)(image_floats<SKC>() + df_warps<SKC>() * (BUFA + BUFB)) * sizeof(float); }
static_assert(df_smem<true>() <= 227 * 1024 - 1024 && df_smem<false>() <= 227 * 1024 - 1024, "shared memory budget");
static_assert((WTOT % 4) == 0 && (WTOT_SKC % 4) == 0 && (BUFA % 4) == 0 && (BUFB % 4) == 0, "16-byte aligned buffers");

struct DecFusedArgs {
    const float* y2;                 // [B,T,1024] output of GRU layer 2
    const float *ln_g, *ln_b;
    float eps;
    const float* skip[4];            // skip4 [B,T,64,16] (added to LN2), skip3 [B,T,32,32], skip2 [B,T,16,64], skip1 [B,T,8,128];
                                     // SKC: skip[0] = e4 TIME-MAJOR [T,B,64,16], skip[1] = e3 [B,T,32,32] (the skip convs' inputs)
    const float* image;              // image_floats<SKC>() floats written by decoder_fused_prep_kernel
    float* mask;                     // [B,T,256]
    int B, T, t0, t1;
    // optional: the frame's share of wo_male (loss_func/loss.py:121-148) on est = mask * X, formed right where the mask is produced
    const float* ref;                // clean spectrum S
    cruse_cplx_layout lr;
    const float* unp;                // noisy spectrum X
    cruse_cplx_layout lu;
    float* loss_rows;                // [B*T] one partial sum per frame (fixed order: the loss does not depend on how the frames are grouped)
};

struct DecPrepArgs {
    const float* w[4];               // conv4_t [64,32,1,3], conv3_t [32,16,1,3], conv2_t [16,8,1,3], conv1_t [8,1,1,3]
    const float* bias[4];
    const float* scale[3];           // folded eval BatchNorm of stages 4..2
    const float* shift[3];
    const float* alpha[3];           // PReLU slopes (null: ReLU)
    const float *wskip4, *wskip3;    // skip_connect_4.weight [64,64,1,3], skip_connect_3.weight [32,32,1,3] or null
    float* image;
};

__device__ __forceinline__ float tf32r(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

__device__ __forceinline__ float2 ld_cplx(const float* __restrict__ p, long long off, long long im_off) {
    if (im_off == 1 && ((off & 1) == 0)) return __ldg(reinterpret_cast<const float2*>(p + off));
    return make_float2(__ldg(p + off), __ldg(p + off + im_off));
}
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// one bin of wo_male with |mask * X| = mask * |X| (the arithmetic of wo_male_partial_kernel<true>, loss.cu)
__device__ __forceinline__ float wo_male_bin(float2 r, float2 u, float mk) {
    const float mr = sqrt_approx(r.x * r.x + r.y * r.y), mu = sqrt_approx(u.x * u.x + u.y * u.y);
    const float iam = mr * rcp_approx(mu);
    const float w = ex2_approx(2.f * 1.4426950408889634f * rcp_approx(1.f + iam));
    const float d = 0.30102999566398120f * lg2_approx((fabsf(mk) * mu + 1.f) * rcp_approx(mr + 1.f));
    return w * fabsf(d);
}

// the skip values a lane adds in the epilogue of a stage (its accumulator positions), loaded one stage AHEAD of their use so that the
// HBM latency hides under the previous stage's MMAs: [m-tile][n-tile][row g / g+8][channel 2t / 2t+1] x (even bin, odd bin)
template <int COUT, int FIN>
struct SkipFrag {
    float2 v[FIN / 16][COUT / 8][2][2];
    __device__ __forceinline__ void load(const float* __restrict__ skip, int lane) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int mt = 0; mt < FIN / 16; ++mt)
#pragma unroll
            for (int nt = 0; nt < COUT / 8; ++nt)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        v[mt][nt][r][c] = __ldg(reinterpret_cast<const float2*>(skip + (size_t)(nt * 8 + 2 * t + c) * (2 * FIN) + 2 * (mt * 16 + g + 8 * r)));
    }
};

// one transposed-conv stage of one frame, by one warp.  in: [1 + FIN][CIN + 4] (row 0 = zeros); w: [3][COUT][CIN + 4];
// out: [1 + 2 FIN][COUT + 4] with a zero row 0, or (OUT_CM) channel-major [COUT][2 FIN + 4] with two zero columns in front.
template <int CIN, int COUT, int FIN, bool OUT_CM, bool HAS_SKIP = true>
__device__ __forceinline__ void convT_stage(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                                            const float* __restrict__ ep, const SkipFrag<COUT, FIN>& sk, int lane) {
    constexpr int LDI = CIN + 4, NT = COUT / 8, KS = CIN / 8, MT = FIN / 16, FOUT = 2 * FIN;
    constexpr int LDO = OUT_CM ? (FOUT + 4) : (COUT + 4);
    const int g = lane >> 2, t = lane & 3;
    if (OUT_CM) {
        if (lane < 2 * COUT) out[(lane >> 1) * LDO + (lane & 1)] = 0.f;
    } else {
        for (int i = lane; i < LDO; i += 32) out[i] = 0.f;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        float ae[NT][4], ao[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) { ae[nt][e] = 0.f; ao[nt][e] = 0.f; }
        const float* xc = in + (1 + mt * 16 + g) * LDI + t;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const float* x = xc + ks * 8;
            const float cur[4] = {x[0], x[8 * LDI], x[4], x[8 * LDI + 4]};
            const float prv[4] = {x[-LDI], x[7 * LDI], x[4 - LDI], x[7 * LDI + 4]};
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float* wp = w + (nt * 8 + g) * LDI + ks * 8 + t;
                mma_tf32(ae[nt], cur, wp[0], wp[4]);                                        // tap 0 at bin i
                mma_tf32(ae[nt], prv, wp[2 * COUT * LDI], wp[2 * COUT * LDI + 4]);          // tap 2 at bin i - 1
                mma_tf32(ao[nt], cur, wp[COUT * LDI], wp[COUT * LDI + 4]);                  // tap 1 at bin i
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int co = nt * 8 + 2 * t + c, i = mt * 16 + g + 8 * r;
                    const float sc = ep[co], sh = ep[EP_N + co], al = ep[2 * EP_N + co];
                    float ve = fmaf(ae[nt][2 * r + c], sc, sh), vo = fmaf(ao[nt][2 * r + c], sc, sh);
                    ve = (ve > 0.f ? ve : al * ve) + (HAS_SKIP ? sk.v[mt][nt][r][c].x : 0.f);
                    vo = (vo > 0.f ? vo : al * vo) + (HAS_SKIP ? sk.v[mt][nt][r][c].y : 0.f);
                    if (OUT_CM) {
                        *reinterpret_cast<float2*>(out + co * LDO + 2 + 2 * i) = make_float2(ve, vo);
                    } else {                                            // (!HAS_SKIP: skip_conv_add adds the skip and rounds)
                        out[(1 + 2 * i) * LDO + co] = HAS_SKIP ? tf32r(ve) : ve;
                        out[(2 + 2 * i) * LDO + co] = HAS_SKIP ? tf32r(vo) : vo;
                    }
                }
    }
}

// X[1 + pos][co] = tf32( X[1 + pos][co] + sum_tap sum_ci E[pos + tap][ci] * W[tap][co][ci] ): a skip conv (Conv2d(1,3), padding 1, no
// bias) of one frame added onto the operand the next stage reads.  E: [1 + FIN + 1][C + 4] with zero rows at both ends (tf32-rounded);
// W: [3][C][C + 4]; X: [1 + FIN][C + 4] holding the unrounded values so far.
template <int C, int FIN>
__device__ __forceinline__ void skip_conv_add(const float* __restrict__ E, const float* __restrict__ W, float* __restrict__ X, int lane) {
    constexpr int LD = C + 4, NT = C / 8, KS = C / 8, MT = FIN / 16;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
            const float* xr = E + (mt * 16 + g + tap) * LD + t;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float a4[4] = {xr[ks * 8], xr[8 * LD + ks * 8], xr[ks * 8 + 4], xr[8 * LD + ks * 8 + 4]};
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float* wp = W + (tap * C + nt * 8 + g) * LD + ks * 8 + t;
                    mma_tf32(acc[nt], a4, wp[0], wp[4]);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float* x = X + (1 + mt * 16 + g + 8 * r) * LD + nt * 8 + 2 * t + c;
                    *x = tf32r(*x + acc[nt][2 * r + c]);
                }
    }
}

// a frame record [C][FIN] held as 8 float4 per lane (element 4 lane + 128 r + e) -> E[1 + bin][channel] with zero rows 0 and FIN + 1
template <int C, int FIN>
__device__ __forceinline__ void stage_transposed(const float4 (&v)[8], float* __restrict__ E, int lane) {
    constexpr int LD = C + 4, LPC = FIN / 4;                 // lanes per channel row
    for (int i = lane; i < LD; i += 32) { E[i] = 0.f; E[(FIN + 1) * LD + i] = 0.f; }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float* d = E + (1 + 4 * (lane % LPC)) * LD + lane / LPC + (32 / LPC) * r;
        d[0] = tf32r(v[r].x); d[LD] = tf32r(v[r].y); d[2 * LD] = tf32r(v[r].z); d[3 * LD] = tf32r(v[r].w);
    }
}

// the shared-memory image of the decoder's constants: PyTorch [Cin][Cout][1][3] -> [tap][co][ci + 4] (K contiguous), tf32-rounded;
// stage-1 weights and bias in fp32; per channel of stages 4..2: scale, shift with the conv bias folded in, PReLU slope (ReLU: 0)
__global__ void __launch_bounds__(DF_THREADS) decoder_fused_prep_kernel(const DecPrepArgs a) {
    const int tid = threadIdx.x;
    float* sm = a.image;
    const bool skc = a.wskip4 != nullptr;
    for (int i = tid; i < (skc ? WTOT_SKC : WTOT); i += DF_THREADS) sm[i] = 0.f;
    __syncthreads();
    if (skc) {                                               // Conv2d weight [co][ci][1][3] -> [tap][co][ci + 4]
        for (int i = tid; i < C4 * C4 * 3; i += DF_THREADS) {
            const int tap = i % 3, ci = (i / 3) % C4, co = i / (3 * C4);
            sm[WS4_OFF + (tap * C4 + co) * LD4 + ci] = tf32r(__ldg(a.wskip4 + i));
        }
        for (int i = tid; i < C3 * C3 * 3; i += DF_THREADS) {
            const int tap = i % 3, ci = (i / 3) % C3, co = i / (3 * C3);
            sm[WS3_OFF + (tap * C3 + co) * LD3 + ci] = tf32r(__ldg(a.wskip3 + i));
        }
    }
    for (int i = tid; i < C4 * C3 * 3; i += DF_THREADS) {
        const int tap = i % 3, co = (i / 3) % C3, ci = i / (3 * C3);
        sm[W4_OFF + (tap * C3 + co) * LD4 + ci] = tf32r(__ldg(a.w[0] + i));
    }
    for (int i = tid; i < C3 * C2 * 3; i += DF_THREADS) {
        const int tap = i % 3, co = (i / 3) % C2, ci = i / (3 * C2);
        sm[W3_OFF + (tap * C2 + co) * LD3 + ci] = tf32r(__ldg(a.w[1] + i));
    }
    for (int i = tid; i < C2 * C1 * 3; i += DF_THREADS) {
        const int tap = i % 3, co = (i / 3) % C1, ci = i / (3 * C1);
        sm[W2_OFF + (tap * C1 + co) * LD2 + ci] = tf32r(__ldg(a.w[2] + i));
    }
    if (tid < C1 * 3) sm[W1_OFF + tid] = __ldg(a.w[3] + tid);                     // [ci][tap], stage 1 runs in fp32
    if (tid == 31) sm[W1_OFF + 24] = a.bias[3] ? __ldg(a.bias[3]) : 0.f;
    if (tid < EP_N) {
        const int s = tid < C3 ? 0 : (tid < C3 + C2 ? 1 : 2), co = tid - (s == 0 ? 0 : (s == 1 ? C3 : C3 + C2));
        const float sc = a.scale[s] ? __ldg(a.scale[s] + co) : 1.f, sh = a.shift[s] ? __ldg(a.shift[s] + co) : 0.f;
        const float bv = a.bias[s] ? __ldg(a.bias[s] + co) : 0.f;
        sm[EP_OFF + tid] = sc;
        sm[EP_OFF + EP_N + tid] = fmaf(bv, sc, sh);                               // (acc + b) * sc + sh
        sm[EP_OFF + 2 * EP_N + tid] = a.alpha[s] ? __ldg(a.alpha[s] + co) : 0.f;  // ReLU = PReLU with slope 0
    }
}

template <bool SKC>
__global__ void __launch_bounds__(32 * (SKC ? DF_WARPS_SKC : DF_WARPS), 1) decoder_fused_kernel(const DecFusedArgs a) {
    constexpr int NW = SKC ? DF_WARPS_SKC : DF_WARPS, IMG = SKC ? WTOT_SKC : WTOT;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t wbar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- constants of all stages: one bulk copy of the prepared image, awaited in front of the first GEMM
    if (tid == 0) {
        tc::mbar_init(&wbar, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tc::mbar_expect_tx(&wbar, (uint32_t)(IMG * sizeof(float)));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(sm)), "l"(a.image),
                     "r"((uint32_t)(IMG * sizeof(float))), "r"(tc::smem_u32(&wbar))
                     : "memory");
    }
    bool have_w = false;

    float* bufA = sm + IMG + warp * (BUFA + BUFB);
    float* bufB = bufA + BUFA;
    const int Tc = a.t1 - a.t0;
    const long long nfr = (long long)a.B * Tc;
    for (long long fr = (long long)blockIdx.x * NW + warp; fr < nfr; fr += (long long)gridDim.x * NW) {
        const long long bsel = fr / Tc, tsel = a.t0 + (fr % Tc);
        const long long row = bsel * a.T + tsel;
        // ---- LayerNorm 2 over the frame's 1024 features (same summation order as layernorm_fwd_kernel) + skip 4 -> X4
        const float* xr = a.y2 + row * (C4 * F4);
        // SKC: the skip conv's input e4 (time-major record t*B + b); else the skip-4 values themselves
        const float* s4 = a.skip[0] + (SKC ? (tsel * a.B + bsel) : row) * (C4 * F4);
        float4 v[8], rv[8];
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            v[r] = __ldg(reinterpret_cast<const float4*>(xr + 4 * lane + 128 * r));
            rv[r] = __ldg(reinterpret_cast<const float4*>(s4 + 4 * lane + 128 * r));
            s += (v[r].x + v[r].y) + (v[r].z + v[r].w);
        }
        const float mean = warp_sum(s) / (float)(C4 * F4);
        float q = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float d0 = v[r].x - mean, d1 = v[r].y - mean, d2 = v[r].z - mean, d3 = v[r].w - mean;
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)(C4 * F4) + a.eps);
        if (SKC) stage_transposed<C4, F4>(rv, bufB, lane);                        // E4 for the skip conv
        for (int i = lane; i < LD4; i += 32) bufA[i] = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int idx = 4 * lane + 128 * r;                                   // feature = channel * 16 + bin
            const float4 gm = __ldg(reinterpret_cast<const float4*>(a.ln_g + idx)), bt = __ldg(reinterpret_cast<const float4*>(a.ln_b + idx));
            float* d = bufA + (1 + 4 * (lane & 3)) * LD4 + (lane >> 2) + 8 * r;
            if (SKC) {                                                            // skip_conv_add adds skip 4 and rounds
                d[0] = (v[r].x - mean) * rstd * gm.x + bt.x;
                d[LD4] = (v[r].y - mean) * rstd * gm.y + bt.y;
                d[2 * LD4] = (v[r].z - mean) * rstd * gm.z + bt.z;
                d[3 * LD4] = (v[r].w - mean) * rstd * gm.w + bt.w;
            } else {
                d[0] = tf32r((v[r].x - mean) * rstd * gm.x + bt.x + rv[r].x);
                d[LD4] = tf32r((v[r].y - mean) * rstd * gm.y + bt.y + rv[r].y);
                d[2 * LD4] = tf32r((v[r].z - mean) * rstd * gm.z + bt.z + rv[r].z);
                d[3 * LD4] = tf32r((v[r].w - mean) * rstd * gm.w + bt.w + rv[r].w);
            }
        }
        __syncwarp();
        SkipFrag<C3, F4> sk3;
        float4 e3[8];
        if (SKC) {                                                                // the skip-3 conv's input e3, in flight under skip conv 4 + stage 4
            const float* s3 = a.skip[1] + row * (C3 * F3);
#pragma unroll
            for (int r = 0; r < 8; ++r) e3[r] = __ldg(reinterpret_cast<const float4*>(s3 + 4 * lane + 128 * r));
        } else {
            sk3.load(a.skip[1] + row * (C3 * F3), lane);
        }
        SkipFrag<C2, F3> sk2;
        sk2.load(a.skip[2] + row * (C2 * F2), lane);
        if (!have_w) {
            tc::mbar_wait(&wbar, 0);
            have_w = true;
        }
        if (SKC) {
            skip_conv_add<C4, F4>(bufB, sm + WS4_OFF, bufA, lane);               // X4 += skip_connect_4(e4)   (:155,160)
            __syncwarp();
            convT_stage<C4, C3, F4, false, false>(bufA, bufB, sm + W4_OFF, sm + EP_OFF, sk3, lane);
            __syncwarp();
            stage_transposed<C3, F3>(e3, bufA, lane);
            __syncwarp();
            skip_conv_add<C3, F3>(bufA, sm + WS3_OFF, bufB, lane);               // X3 += skip_connect_3(e3)   (:154,161)
        } else {
            convT_stage<C4, C3, F4, false>(bufA, bufB, sm + W4_OFF, sm + EP_OFF, sk3, lane);
        }
        __syncwarp();
        SkipFrag<C1, F2> sk1;
        sk1.load(a.skip[3] + row * (C1 * F1), lane);
        convT_stage<C3, C2, F3, false>(bufB, bufA, sm + W3_OFF, sm + EP_OFF + C3, sk2, lane);
        __syncwarp();
        // the loss inputs of this frame (bins 2i, 2i+1 for i = lane + 32 r), in flight under stage 2
        float2 lr_[F1 / 32][2], lu_[F1 / 32][2];
        if (a.loss_rows) {
            const long long bb = row / a.T, tt = row - bb * a.T;
            const long long rb = bb * a.lr.sb + tt * a.lr.st, ub = bb * a.lu.sb + tt * a.lu.st;
#pragma unroll
            for (int r = 0; r < F1 / 32; ++r)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int f = 2 * (lane + 32 * r) + c;
                    lr_[r][c] = ld_cplx(a.ref, rb + f * a.lr.sf, a.lr.im_off);
                    lu_[r][c] = ld_cplx(a.unp, ub + f * a.lu.sf, a.lu.im_off);
                }
        }
        convT_stage<C2, C1, F2, true>(bufA, bufB, sm + W2_OFF, sm + EP_OFF + C3 + C2, sk1, lane);
        __syncwarp();
        // ---- stage 1: ConvTranspose2d(8 -> 1) + bias + sigmoid (model/cruse_net.py:164), fp32 FMAs; lane owns the bin pairs i = lane + 32 r
        {
            const float* w1 = sm + W1_OFF;
            const float b1 = w1[24];
            float* mrow = a.mask + row * F0;
            float lacc = 0.f;
#pragma unroll
            for (int r = 0; r < F1 / 32; ++r) {
                const int i = lane + 32 * r;
                float ve = b1, vo = b1;
#pragma unroll
                for (int ci = 0; ci < C1; ++ci) {
                    const float x0 = bufB[ci * LD1 + 2 + i], xm = bufB[ci * LD1 + 1 + i];
                    ve = fmaf(w1[3 * ci], x0, fmaf(w1[3 * ci + 2], xm, ve));
                    vo = fmaf(w1[3 * ci + 1], x0, vo);
                }
                const float m0 = sigmoidf_(ve), m1 = sigmoidf_(vo);
                *reinterpret_cast<float2*>(mrow + 2 * i) = make_float2(m0, m1);
                if (a.loss_rows) lacc += wo_male_bin(lr_[r][0], lu_[r][0], m0) + wo_male_bin(lr_[r][1], lu_[r][1], m1);
            }
            if (a.loss_rows) {
                lacc = warp_sum(lacc);
                if (lane == 0) a.loss_rows[row] = lacc;
            }
        }
        __syncwarp();
    }
}

}  // namespace
}  // namespace cruse

extern "C" long long cruse_decoder_fused_image_floats(int with_skip_convs) { return with_skip_convs ? cruse::WTOT_SKC : cruse::WTOT; }

extern "C" int cruse_decoder_fused_prep(const float* const* w, const float* const* bias, const float* const* scale, const float* const* shift,
                                        const float* const* alpha, int act, const float* wskip4, const float* wskip3, float* image,
                                        void* stream) {
    using namespace cruse;
    CRUSE_CHECK_ARG(w && bias && scale && shift && image, "decoder_fused_prep: null pointer");
    CRUSE_CHECK_ARG(act == CRUSE_ACT_RELU || act == CRUSE_ACT_PRELU, "decoder_fused_prep: activation %d (ReLU / PReLU only)", act);
    CRUSE_CHECK_ARG((wskip4 == nullptr) == (wskip3 == nullptr), "decoder_fused_prep: the weights of skip convs 4 and 3 come together or not at all");
    DecPrepArgs a;
    a.image = image; a.wskip4 = wskip4; a.wskip3 = wskip3;
    for (int s = 0; s < 4; ++s) {
        CRUSE_CHECK_ARG(w[s], "decoder_fused_prep: null weight pointer of stage %d", 4 - s);
        a.w[s] = w[s]; a.bias[s] = bias[s];
        if (s < 3) {
            a.scale[s] = scale[s]; a.shift[s] = shift[s];
            a.alpha[s] = (act == CRUSE_ACT_PRELU && alpha) ? alpha[s] : nullptr;
            CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || a.alpha[s], "decoder_fused_prep: PReLU without slopes for stage %d", 4 - s);
        }
    }
    decoder_fused_prep_kernel<<<1, DF_THREADS, 0, (cudaStream_t)stream>>>(a);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_decoder_fused_range(const float* y2, const float* ln_gamma, const float* ln_beta, float ln_eps,
                                         const float* const* skips, int skip_convs, const float* image, float* mask, const float* ref,
                                         cruse_cplx_layout lref, const float* unproc, cruse_cplx_layout lunp, float* loss_rows, int B, int T,
                                         int t_begin, int t_end, int max_ctas, void* stream) {
    using namespace cruse;
    CRUSE_CHECK_ARG(y2 && ln_gamma && ln_beta && skips && image && mask, "decoder_fused_range: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && t_begin >= 0 && t_begin < t_end && t_end <= T, "decoder_fused_range: bad sizes B=%d T=%d range [%d,%d)", B, T,
                    t_begin, t_end);
    CRUSE_CHECK_ARG(((uintptr_t)image & 15) == 0, "decoder_fused_range: image must be 16-byte aligned");
    DecFusedArgs a;
    a.y2 = y2; a.ln_g = ln_gamma; a.ln_b = ln_beta; a.eps = ln_eps; a.mask = mask; a.image = image;
    a.B = B; a.T = T; a.t0 = t_begin; a.t1 = t_end;
    CRUSE_CHECK_ARG(!loss_rows || (ref && unproc), "decoder_fused_range: loss rows requested without the clean / noisy spectra");
    a.ref = ref; a.lr = lref; a.unp = unproc; a.lu = lunp; a.loss_rows = loss_rows;
    for (int s = 0; s < 4; ++s) {
        CRUSE_CHECK_ARG(skips[s], "decoder_fused_range: null skip pointer of stage %d", 4 - s);
        a.skip[s] = skips[s];
    }
    static bool attr_set = false;
    if (!attr_set) {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(decoder_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)df_smem<false>()));
        CRUSE_CUDA_OK(cudaFuncSetAttribute(decoder_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)df_smem<true>()));
        attr_set = true;
    }
    const int nw = skip_convs ? DF_WARPS_SKC : DF_WARPS;
    const long long frames = (long long)B * (t_end - t_begin);
    long long grid = (frames + nw - 1) / nw;
    const long long cap = max_ctas > 0 ? max_ctas : sm_count();
    if (grid > cap) grid = cap;
    if (skip_convs)
        decoder_fused_kernel<true><<<(unsigned)grid, 32 * DF_WARPS_SKC, df_smem<true>(), (cudaStream_t)stream>>>(a);
    else
        decoder_fused_kernel<false><<<(unsigned)grid, 32 * DF_WARPS, df_smem<false>(), (cudaStream_t)stream>>>(a);
    CRUSE_LAUNCH_OK();
    return 0;
}

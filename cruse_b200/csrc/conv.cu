// conv.cu -- encoder / skip / decoder stages of the CRUSE U-Net on frame-major activations.
//
// Replaces nn.Conv2d((2,3), stride (1,2), pad (1,1)) + "[..., :-1, :]" + BatchNorm2d + act
// (model/cruse_net.py:138,141,149-152), the (1,3) skip convs (:143,153-156) and
// nn.ConvTranspose2d((1,3), stride (1,2)) + "[..., :-1]" + BatchNorm2d + act + skip add /
// sigmoid (:161-164) of the reference.
//
// Layout: activations are [B, T, C, F] -- every frame is one contiguous C*F record (4 KB at
// every stage of the 256-bin pyramid), so a CTA streams TT consecutive frames of one
// utterance into shared memory with fully coalesced 128-bit loads (one extra look-back frame
// for the causal (2,3) kernels), keeps the stage's whole weight tensor in shared memory, and
// every thread register-tiles CO_T output channels x TT frames for one output bin.  Bias,
// folded BatchNorm (eval) or per-channel sum / sum-of-squares partials (train), activation
// and the decoder's skip add all live in the epilogue, so a stage reads its input once and
// writes its output once (SURVEY.md App. B byte model).
#include "common.cuh"

namespace cruse {

// conv_tc.cu: tensor-core (tcgen05) implicit-GEMM instantiations for the 256-bin pyramid in eval mode
int conv_tc_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                int act, const float* addend, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride,
                int in_tm, int out_tm, int wmode, cudaStream_t st, const float* hist = nullptr, int t_begin = 0, int t_end = 0);
int convT_tc_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                 int act, const float* skip, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, cudaStream_t st,
                 int t_begin = 0, int t_end = 0);
int conv_skip_tc_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                     int act, const float* w2, float* out, float* out2, int B, int T, int Cin, int Fin, int Cout, int Fout, int out_tm,
                     cudaStream_t st, int t_begin, int t_end);
int conv_dgrad_tc_try(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin, int Fin, int Cout,
                      int Fout, int kt, cudaStream_t st);
int convT_dgrad_tc_try(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin, int Fin, int Cout,
                       int Fout, cudaStream_t st);

// conv_edge.cu: streaming kernels for the single-channel stages (Cin == 1 / Cout == 1)
int convT_edge_dgrad_try(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin, int Fin, int Cout,
                         int Fout, cudaStream_t st);
int conv_edge_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                  int act, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride, cudaStream_t st,
                  const float* hist = nullptr, int t_begin = 0, int t_end = 0);
int convT_edge_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                   int act, const float* skip, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, cudaStream_t st,
                   int t_begin = 0, int t_end = 0);

constexpr int CONV_TT = 8;    // frames per CTA
constexpr int CONV_COT = 4;   // output channels per thread
constexpr int CONV_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// forward conv, KT time taps (look-back), 3 freq taps, freq stride SF, freq pad 1.
// smem: s_in[(TT+KT-1)][Cin][Fin+2]  (zero column on both sides),  s_w[Cin][KT*3][CoutP] (CoutP = Cout
// rounded up to CONV_COT), s_stats[2*Cout]
// ---------------------------------------------------------------------------------------------
// wmode selects how `w` is read: 0 = Conv2d weight [Cout][Cin][KT][3] (forward);
//   1 = data gradient of a (1,3)/stride-1 conv: w is that conv's weight [Cin_k][Cout_k][1][3], taps flipped;
//   2 = data gradient of a ConvTranspose2d (1,3)/stride-2: w is its weight [Cout_k][Cin_k][1][3] (run with SF=2, pad_left=0).
// pad_left = zero columns left of bin 0 (1 for the forward convs); addend (optional) is added to the result.
template <int KT, int SF>
__global__ void __launch_bounds__(CONV_THREADS)
conv_fwd_kernel(const float* __restrict__ in, const float* __restrict__ hist, const float* __restrict__ w,
                const float* __restrict__ bias,
                const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ alpha,
                int act, const float* __restrict__ addend, float* __restrict__ out, float* __restrict__ stats_ws, int T, int Cin, int Fin, int Cout,
                int Fout, int pad_left, int wmode) {
    extern __shared__ float smem[];
    constexpr int TT = CONV_TT, COT = CONV_COT, NR = TT + KT - 1;
    const int FinP = Fin + 3;   // zero column left, two right (the stride-2 gradient form reads up to bin Fin+1)
    const int CoutP = (Cout + COT - 1) / COT * COT;
    float* s_in = smem;
    float* s_w = s_in + (((size_t)NR * Cin * FinP + 3) & ~(size_t)3);  // keep float4 alignment
    float* s_stats = s_w + (size_t)Cin * KT * 3 * CoutP;

    const int chunks = (T + TT - 1) / TT;
    const int b = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * TT;
    const int tid = threadIdx.x;

    // weights: PyTorch [Cout][Cin][KT][3] -> s_w[ci][tap][co]
    const int nw = Cin * KT * 3;
    for (int i = tid; i < nw * CoutP; i += blockDim.x) {
        const int co = i % CoutP, r = i / CoutP;  // r = ci*(KT*3) + tap
        float v = 0.f;
        if (co < Cout) {
            if (wmode == 0) {
                v = __ldg(w + (size_t)co * nw + r);
            } else {
                const int ci = r / 3, kf = r - ci * 3;                         // KT == 1 in the gradient modes
                v = (wmode == 1) ? __ldg(w + ((size_t)ci * Cout + co) * 3 + (2 - kf))     // [Cin_k][Cout_k][1][3], flipped
                                 : __ldg(w + ((size_t)co * Cin + ci) * 3 + kf);           // [Cout_k][Cin_k][1][3]
            }
        }
        s_w[i] = v;
    }
    if (stats_ws)
        for (int i = tid; i < 2 * Cout; i += blockDim.x) s_stats[i] = 0.f;
    // input rows t0-(KT-1) .. t0+TT-1
    const int rowlen = Cin * Fin;
    const float* inb = in + (size_t)b * T * rowlen;
    for (int r = 0; r < NR; ++r) {
        const int t = t0 - (KT - 1) + r;
        bool valid = (t >= 0 && t < T);
        float* dst = s_in + (size_t)r * Cin * FinP;
        const float* src = inb + (size_t)t * rowlen;
        if (KT == 2 && t == -1 && hist) {  // streaming: frame -1 is the last input frame of the previous chunk
            valid = true;
            src = hist + (size_t)b * rowlen;
        }
        if ((Fin & 3) == 0) {
            for (int i = tid * 4; i < rowlen; i += blockDim.x * 4) {
                float4 v = valid ? __ldg(reinterpret_cast<const float4*>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const int ci = i / Fin, f = i - ci * Fin;
                float* d = dst + ci * FinP + 1 + f;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        } else {
            for (int i = tid; i < rowlen; i += blockDim.x) {
                const int ci = i / Fin, f = i - ci * Fin;
                dst[ci * FinP + 1 + f] = valid ? __ldg(src + i) : 0.f;
            }
        }
        for (int ci = tid; ci < Cin; ci += blockDim.x) {
            dst[ci * FinP] = 0.f;
            dst[ci * FinP + Fin + 1] = 0.f;
            dst[ci * FinP + Fin + 2] = 0.f;
        }
    }
    __syncthreads();

    const int ncob = CoutP / COT;
    const int items = Fout * ncob;
    for (int item = tid; item < items; item += blockDim.x) {
        const int fo = item % Fout, cob = item / Fout;
        const int co0 = cob * COT;
        float acc[TT][COT];
#pragma unroll
        for (int t = 0; t < TT; ++t)
#pragma unroll
            for (int c = 0; c < COT; ++c) acc[t][c] = 0.f;

        const float* ip = s_in + SF * fo + (1 - pad_left);  // smem column 0 is the zero pad: tap kf=0 of output fo sits at unpadded SF*fo - pad_left
        for (int ci = 0; ci < Cin; ++ci) {
            float xv[NR][3];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const float* p = ip + ((size_t)r * Cin + ci) * FinP;
                xv[r][0] = p[0]; xv[r][1] = p[1]; xv[r][2] = p[2];
            }
            const float* wp = s_w + (size_t)ci * KT * 3 * CoutP + co0;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int kf = 0; kf < 3; ++kf) {
                    const float4 wv = *reinterpret_cast<const float4*>(wp + (kt * 3 + kf) * CoutP);
#pragma unroll
                    for (int t = 0; t < TT; ++t) {
                        const float x = xv[t + kt][kf];  // out frame t uses rows t (kt=0: t-1) .. t+KT-1
                        acc[t][0] = fmaf(wv.x, x, acc[t][0]);
                        acc[t][1] = fmaf(wv.y, x, acc[t][1]);
                        acc[t][2] = fmaf(wv.z, x, acc[t][2]);
                        acc[t][3] = fmaf(wv.w, x, acc[t][3]);
                    }
                }
        }
        // epilogue
#pragma unroll
        for (int c = 0; c < COT; ++c) {
            const int co = co0 + c;
            if (co >= Cout) continue;
            const float bv = bias ? __ldg(bias + co) : 0.f;
            const float sc = scale ? __ldg(scale + co) : 1.f, sh = scale ? __ldg(shift + co) : 0.f;
            const float al = alpha ? __ldg(alpha + co) : 0.f;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t0 + t < T) {
                    float v = acc[t][c] + bv;
                    s1 += v; s2 += v * v;
                    v = apply_act(fmaf(v, sc, sh), act, al);
                    const size_t o = (((size_t)b * T + t0 + t) * Cout + co) * Fout + fo;
                    if (addend) v += __ldg(addend + o);
                    out[o] = v;
                }
            }
            if (stats_ws) {
                atomicAdd(&s_stats[co], s1);
                atomicAdd(&s_stats[Cout + co], s2);
            }
        }
    }
    if (stats_ws) {
        __syncthreads();
        float* so = stats_ws + (size_t)blockIdx.x * 2 * Cout;
        for (int i = tid; i < 2 * Cout; i += blockDim.x) so[i] = s_stats[i];
    }
}

// ---------------------------------------------------------------------------------------------
// transposed conv (1,3), stride (1,2), no padding, cropped to Fout:
//   out[co, 2i]   = b + sum_ci W[ci,co,0]*in[ci,i] + W[ci,co,2]*in[ci,i-1]
//   out[co, 2i+1] = b + sum_ci W[ci,co,1]*in[ci,i]
// smem: s_in[TT][Cin][Fin+2], s_w[Cin][3][CoutP], s_stats[2*Cout]
// ---------------------------------------------------------------------------------------------
// KT = 2 is the data gradient of the causal (2,3)/stride-(1,2) encoder conv (w = that conv's weight
// [Cin_k][Cout_k][2][3]): time tap kt reads input frame t + 1 - kt (kt = 0 looks one frame AHEAD), and the
// result is shifted by `oshift` = 1 bins (convT output u = f + 1, because the conv padded one bin on the left).
template <int KT>
__global__ void __launch_bounds__(CONV_THREADS)
convT_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ alpha,
                 int act, const float* __restrict__ skip, float* __restrict__ out, float* __restrict__ stats_ws, int T,
                 int Cin, int Fin, int Cout, int Fout, int oshift) {
    extern __shared__ float smem[];
    constexpr int TT = CONV_TT, COT = CONV_COT, NR = CONV_TT + KT - 1;
    const int FinP = Fin + 2;
    const int CoutP = (Cout + COT - 1) / COT * COT;
    float* s_in = smem;
    float* s_w = s_in + (((size_t)NR * Cin * FinP + 3) & ~(size_t)3);  // keep float4 alignment
    float* s_stats = s_w + (size_t)Cin * KT * 3 * CoutP;
    const int chunks = (T + TT - 1) / TT;
    const int b = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * TT;
    const int tid = threadIdx.x;

    // weights: PyTorch [Cin][Cout][KT][3] -> s_w[ci][kt*3+kf][co]
    for (int i = tid; i < Cin * KT * 3 * CoutP; i += blockDim.x) {
        const int co = i % CoutP, r = i / CoutP;
        const int tap = r % (KT * 3), ci = r / (KT * 3);
        s_w[i] = (co < Cout) ? __ldg(w + ((size_t)ci * Cout + co) * (KT * 3) + tap) : 0.f;
    }
    if (stats_ws)
        for (int i = tid; i < 2 * Cout; i += blockDim.x) s_stats[i] = 0.f;
    const int rowlen = Cin * Fin;
    const float* inb = in + (size_t)b * T * rowlen;
    for (int r = 0; r < NR; ++r) {
        const int t = t0 + r;
        const bool valid = t < T;
        float* dst = s_in + (size_t)r * Cin * FinP;
        const float* src = inb + (size_t)t * rowlen;
        if ((Fin & 3) == 0) {
            for (int i = tid * 4; i < rowlen; i += blockDim.x * 4) {
                float4 v = valid ? __ldg(reinterpret_cast<const float4*>(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const int ci = i / Fin, f = i - ci * Fin;
                float* d = dst + ci * FinP + 1 + f;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        } else {
            for (int i = tid; i < rowlen; i += blockDim.x) {
                const int ci = i / Fin, f = i - ci * Fin;
                dst[ci * FinP + 1 + f] = valid ? __ldg(src + i) : 0.f;
            }
        }
        for (int ci = tid; ci < Cin; ci += blockDim.x) {
            dst[ci * FinP] = 0.f;
            dst[ci * FinP + Fin + 1] = 0.f;
        }
    }
    __syncthreads();

    const int npair = (Fout + oshift + 1) / 2;
    const int ncob = CoutP / COT;
    const int items = npair * ncob;
    for (int item = tid; item < items; item += blockDim.x) {
        const int i = item % npair, cob = item / npair;
        const int co0 = cob * COT;
        float ae[TT][COT], ao[TT][COT];
#pragma unroll
        for (int t = 0; t < TT; ++t)
#pragma unroll
            for (int c = 0; c < COT; ++c) { ae[t][c] = 0.f; ao[t][c] = 0.f; }
        for (int ci = 0; ci < Cin; ++ci) {
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const float* wp = s_w + ((size_t)ci * KT + kt) * 3 * CoutP + co0;
                const float4 w0 = *reinterpret_cast<const float4*>(wp);
                const float4 w1 = *reinterpret_cast<const float4*>(wp + CoutP);
                const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * CoutP);
#pragma unroll
                for (int t = 0; t < TT; ++t) {
                    const float* p = s_in + ((size_t)(t + (KT - 1) - kt) * Cin + ci) * FinP + i;  // p[0]=in[i-1], p[1]=in[i]
                    const float xm = p[0], x0 = p[1];
                    ae[t][0] = fmaf(w0.x, x0, fmaf(w2.x, xm, ae[t][0]));
                    ae[t][1] = fmaf(w0.y, x0, fmaf(w2.y, xm, ae[t][1]));
                    ae[t][2] = fmaf(w0.z, x0, fmaf(w2.z, xm, ae[t][2]));
                    ae[t][3] = fmaf(w0.w, x0, fmaf(w2.w, xm, ae[t][3]));
                    ao[t][0] = fmaf(w1.x, x0, ao[t][0]);
                    ao[t][1] = fmaf(w1.y, x0, ao[t][1]);
                    ao[t][2] = fmaf(w1.z, x0, ao[t][2]);
                    ao[t][3] = fmaf(w1.w, x0, ao[t][3]);
                }
            }
        }
        const int fe = 2 * i - oshift, fo = 2 * i + 1 - oshift;
        const bool has_even = fe >= 0 && fe < Fout;
        const bool has_odd = fo < Fout;
#pragma unroll
        for (int c = 0; c < COT; ++c) {
            const int co = co0 + c;
            if (co >= Cout) continue;
            const float bv = bias ? __ldg(bias + co) : 0.f;
            const float sc = scale ? __ldg(scale + co) : 1.f, sh = scale ? __ldg(shift + co) : 0.f;
            const float al = alpha ? __ldg(alpha + co) : 0.f;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                if (t0 + t < T) {
                    const size_t o = (((size_t)b * T + t0 + t) * Cout + co) * Fout;
                    float v0 = ae[t][c] + bv, v1 = ao[t][c] + bv;
                    if (has_even) {
                        s1 += v0; s2 += v0 * v0;
                        v0 = apply_act(fmaf(v0, sc, sh), act, al);
                        if (skip) v0 += __ldg(skip + o + fe);
                    }
                    if (has_odd) {
                        s1 += v1; s2 += v1 * v1;
                        v1 = apply_act(fmaf(v1, sc, sh), act, al);
                        if (skip) v1 += __ldg(skip + o + fo);
                    }
                    if (has_even && has_odd && oshift == 0 && (Fout & 1) == 0) {
                        *reinterpret_cast<float2*>(out + o + fe) = make_float2(v0, v1);
                    } else {
                        if (has_even) out[o + fe] = v0;
                        if (has_odd) out[o + fo] = v1;
                    }
                }
            }
            if (stats_ws) {
                atomicAdd(&s_stats[co], s1);
                atomicAdd(&s_stats[Cout + co], s2);
            }
        }
    }
    if (stats_ws) {
        __syncthreads();
        float* so = stats_ws + (size_t)blockIdx.x * 2 * Cout;
        for (int i = tid; i < 2 * Cout; i += blockDim.x) so[i] = s_stats[i];
    }
}

// Per-channel sum / sum of squares of a stage output z [B,T,C,F] (pre-BatchNorm) in the per-chunk partial layout the conv
// kernels emit ([nparts][2*C], one partial per 8 frames of one utterance): used when the conv itself ran on the tensor
// cores (conv_tc.cu has no statistics epilogue).
// bn_stats_pos_kernel (frames of up to 8192 floats): a thread owns NIT fixed float4 positions of the frame -- hence fixed
// channels -- and adds them over the chunk's frames with all of the chunk's loads in flight at once; the CTA then adds, per
// channel, the contiguous run of positions in order.  Deterministic, no atomics.  (bn_stats_kernel below, one warp per channel
// with one load per lane in flight, needed 8 dependent round trips per CTA on the 64-channel stages: 19 us for 53 MB.)
template <int NIT>
__global__ void __launch_bounds__(256)
bn_stats_pos_kernel(const float* __restrict__ z, float* __restrict__ stats_ws, int T, int C, int F) {
    constexpr int UF = NIT == 1 ? 8 : (NIT == 2 ? 4 : (NIT == 4 ? 2 : 1));
    const int chunks = (T + CONV_TT - 1) / CONV_TT;
    const int b = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * CONV_TT;
    const int nfr = (T - t0) < CONV_TT ? (T - t0) : CONV_TT;
    const int n4 = (C * F) >> 2;
    const float4* zb = reinterpret_cast<const float4*>(z + ((size_t)b * T + t0) * C * F);
    float s1[NIT], s2[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) s1[k] = s2[k] = 0.f;
    for (int fr0 = 0; fr0 < nfr; fr0 += UF) {
        float4 v[UF][NIT];
#pragma unroll
        for (int u = 0; u < UF; ++u)
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int j = threadIdx.x + k * 256;
                v[u][k] = (fr0 + u < nfr && j < n4) ? __ldg(zb + (size_t)(fr0 + u) * n4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
        for (int u = 0; u < UF; ++u)
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const float4 w = v[u][k];
                s1[k] += (w.x + w.y) + (w.z + w.w);
                s2[k] += (w.x * w.x + w.y * w.y) + (w.z * w.z + w.w * w.w);
            }
    }
    __shared__ float s_acc[2][NIT * 256];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        s_acc[0][threadIdx.x + k * 256] = s1[k];
        s_acc[1][threadIdx.x + k * 256] = s2[k];
    }
    __syncthreads();
    float* so = stats_ws + (size_t)blockIdx.x * 2 * C;
    const int per = F >> 2;                                       // float4 positions per channel (F % 4 == 0)
    for (int i = threadIdx.x; i < 2 * C; i += 256) {
        const int q = i / C, c = i - q * C;
        float a = 0.f;
        for (int j = c * per; j < (c + 1) * per; ++j) a += s_acc[q][j];
        so[i] = a;
    }
}

// the same for frames of any size: one warp owns a channel, lanes stride over its 8 x F values, the warp shuffle reduction has a
// fixed order
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ z, float* __restrict__ stats_ws, int T, int C, int F) {
    const int chunks = (T + CONV_TT - 1) / CONV_TT;
    const int b = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * CONV_TT;
    const int nfr = (T - t0) < CONV_TT ? (T - t0) : CONV_TT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float* zb = z + ((size_t)b * T + t0) * C * F;
    float* so = stats_ws + (size_t)blockIdx.x * 2 * C;
    const int per_fr4 = F >> 2;                                  // float4 per (frame, channel) row; F % 4 == 0
    for (int c = warp; c < C; c += nwarps) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 4                                                // four independent loads in flight per lane, same summation order
        for (int i = lane; i < nfr * per_fr4; i += 32) {
            const int fr = i / per_fr4, q = i - fr * per_fr4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(zb + ((size_t)fr * C + c) * F) + q);
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) { so[c] = s1; so[C + c] = s2; }
    }
}

static void launch_bn_stats(const float* z, float* stats_ws, int B, int T, int C, int F, cudaStream_t st) {
    const int grid = B * ((T + CONV_TT - 1) / CONV_TT);
    const int n4 = (C * F) >> 2;
    if (n4 <= 256) bn_stats_pos_kernel<1><<<grid, 256, 0, st>>>(z, stats_ws, T, C, F);
    else if (n4 <= 512) bn_stats_pos_kernel<2><<<grid, 256, 0, st>>>(z, stats_ws, T, C, F);
    else if (n4 <= 1024) bn_stats_pos_kernel<4><<<grid, 256, 0, st>>>(z, stats_ws, T, C, F);
    else if (n4 <= 2048) bn_stats_pos_kernel<8><<<grid, 256, 0, st>>>(z, stats_ws, T, C, F);
    else bn_stats_kernel<<<grid, 256, 0, st>>>(z, stats_ws, T, C, F);
}

// one block per channel: reduce per-CTA partials in double
__global__ void bn_finalize_kernel(const float* __restrict__ stats_ws, int nparts, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                                   float* save_mean, float* save_invstd) {
    const int c = blockIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int p = threadIdx.x; p < nparts; p += blockDim.x) {
        s1 += (double)stats_ws[(size_t)p * 2 * C + c];
        s2 += (double)stats_ws[(size_t)p * 2 * C + C + c];
    }
    __shared__ double sh1[32], sh2[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) { sh1[threadIdx.x >> 5] = s1; sh2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, q = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh1[i]; q += sh2[i]; }
        const double mean = a / count;
        double var = q / count - mean * mean;
        if (var < 0.0) var = 0.0;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
        scale[c] = g * invstd;
        shift[c] = bt - (float)mean * g * invstd;
        if (save_mean) save_mean[c] = (float)mean;
        if (save_invstd) save_invstd[c] = invstd;
        if (running_mean) {
            const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float inv = 1.0f / sqrtf(var[c] + eps);
    const float g = gamma ? gamma[c] : 1.f;
    scale[c] = g * inv;
    shift[c] = (beta ? beta[c] : 0.f) - mean[c] * g * inv;
}

__global__ void bn_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const float* __restrict__ alpha, int act,
                                  const float* __restrict__ skip, float* __restrict__ y, long long total, int C, int F) {
    // vectorised when F % 4 == 0 (a float4 never straddles a channel)
    if ((F & 3) == 0) {
        const long long n4 = total >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const int c = (int)(((i << 2) / F) % C);
            const float sc = __ldg(scale + c), sh = __ldg(shift + c), al = alpha ? __ldg(alpha + c) : 0.f;
            float4 v = __ldg(reinterpret_cast<const float4*>(z) + i);
            v.x = apply_act(fmaf(v.x, sc, sh), act, al);
            v.y = apply_act(fmaf(v.y, sc, sh), act, al);
            v.z = apply_act(fmaf(v.z, sc, sh), act, al);
            v.w = apply_act(fmaf(v.w, sc, sh), act, al);
            if (skip) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(skip) + i);
                v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
            }
            reinterpret_cast<float4*>(y)[i] = v;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int c = (int)((i / F) % C);
            float v = apply_act(fmaf(__ldg(z + i), __ldg(scale + c), __ldg(shift + c)), act, alpha ? __ldg(alpha + c) : 0.f);
            if (skip) v += __ldg(skip + i);
            y[i] = v;
        }
    }
}

static size_t conv_smem_bytes(int KT, int Cin, int Fin, int Cout) {
    const int CoutP = (Cout + CONV_COT - 1) / CONV_COT * CONV_COT;
    return sizeof(float) * (((size_t)(CONV_TT + KT - 1) * Cin * (Fin + 3) + 3 & ~(size_t)3) + (size_t)Cin * KT * 3 * CoutP + 2 * (size_t)Cout);
}

}  // namespace cruse

using namespace cruse;

extern "C" int cruse_conv_nparts(int B, int T) { return B * ((T + CONV_TT - 1) / CONV_TT); }

extern "C" int cruse_conv_fwd(const float* in, const float* hist, const float* w, const float* bias, const float* scale,
                              const float* shift, const float* alpha, int act, float* out, float* stats_ws, int B,
                              int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride, void* stream) {
    CRUSE_CHECK_ARG(in && w && out, "conv_fwd: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "conv_fwd: bad sizes");
    CRUSE_CHECK_ARG((kt == 2 && fstride == 2) || (kt == 1 && fstride == 1), "conv_fwd: supported (kt,fstride) are (2,2) and (1,1), got (%d,%d)", kt, fstride);
    const int expF = (Fin + 2 - 3) / fstride + 1;
    CRUSE_CHECK_ARG(Fout == expF, "conv_fwd: Fout=%d, expected %d", Fout, expF);
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fwd: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "conv_fwd: PReLU needs alpha");
    cudaStream_t st = (cudaStream_t)stream;
    if (!hist && stats_ws && (Fout & 3) == 0 && !scale && act == CRUSE_ACT_NONE) {
        // train-mode stage: the conv (+ bias) on the tensor cores, then one pass over z for the BatchNorm partial sums
        const int rc = conv_tc_try(in, w, bias, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, nullptr, out, B, T, Cin, Fin, Cout, Fout, kt, fstride, 0, 0, 0, st);
        if (rc < 0) return rc;
        int done = rc;
        if (!done) {          // stage 1 (Cin == 1): the exact-fp32 streaming kernel, statistics by the same deterministic pass
            done = conv_edge_try(in, w, bias, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, out, B, T, Cin, Fin, Cout, Fout, kt, fstride, st);
            if (done < 0) { set_error("conv_fwd: streaming stage-1 kernel launch failed"); return done; }
        }
        if (done == 1) {
            launch_bn_stats(out, stats_ws, B, T, Cout, Fout, st);
            CRUSE_LAUNCH_OK();
            return 0;
        }
    }
    if (!stats_ws) {     // eval-mode stage of the 256-bin pyramid (hist: streaming chunk): tcgen05 implicit GEMM (conv_tc.cu)
        int rc = conv_tc_try(in, w, bias, scale, shift, alpha, act, nullptr, out, B, T, Cin, Fin, Cout, Fout, kt, fstride, 0, 0, 0, st, hist);
        if (rc) return rc < 0 ? rc : 0;
        rc = conv_edge_try(in, w, bias, scale, shift, alpha, act, out, B, T, Cin, Fin, Cout, Fout, kt, fstride, st, hist);
        if (rc) { if (rc < 0) set_error("conv_fwd: streaming stage-1 kernel launch failed"); return rc < 0 ? rc : 0; }
    }
    const size_t smem = conv_smem_bytes(kt, Cin, Fin, Cout);
    CRUSE_CHECK_ARG(smem <= 227 * 1024, "conv_fwd: stage (Cin=%d,Fin=%d,Cout=%d) needs %zu B shared memory", Cin, Fin, Cout, smem);
    const int grid = cruse_conv_nparts(B, T);
    if (kt == 2) {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_fwd_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_kernel<2, 2><<<grid, CONV_THREADS, smem, st>>>(in, hist, w, bias, scale, shift, alpha, act, nullptr, out, stats_ws, T, Cin, Fin, Cout, Fout, 1, 0);
    } else {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_fwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_kernel<1, 1><<<grid, CONV_THREADS, smem, st>>>(in, nullptr, w, bias, scale, shift, alpha, act, nullptr, out, stats_ws, T, Cin, Fin, Cout, Fout, 1, 0);
    }
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_conv_fwd_tm(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                                 const float* alpha, int act, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout,
                                 int kt, int fstride, int in_time_major, int out_time_major, void* stream) {
    CRUSE_CHECK_ARG(in && w && out, "conv_fwd_tm: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "conv_fwd_tm: bad sizes");
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fwd_tm: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "conv_fwd_tm: PReLU needs alpha");
    const int rc = conv_tc_try(in, w, bias, scale, shift, alpha, act, nullptr, out, B, T, Cin, Fin, Cout, Fout, kt, fstride,
                               in_time_major ? 1 : 0, out_time_major ? 1 : 0, 0, (cudaStream_t)stream);
    if (rc < 0) return rc;
    CRUSE_CHECK_ARG(rc == 1, "conv_fwd_tm: no tensor-core instantiation for kt=%d fstride=%d Cin=%d Cout=%d Fin=%d (or conv mode is fp32)", kt,
                    fstride, Cin, Cout, Fin);
    return 0;
}

// Eval-mode stages restricted to the output frames [t_begin, t_end) of every utterance (tensor-core / streaming kernels only).
extern "C" int cruse_conv_fwd_range(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                                    const float* alpha, int act, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout,
                                    int kt, int fstride, int in_time_major, int out_time_major, int t_begin, int t_end, void* stream) {
    CRUSE_CHECK_ARG(in && w && out, "conv_fwd_range: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "conv_fwd_range: bad sizes");
    CRUSE_CHECK_ARG(t_begin >= 0 && t_begin < t_end && t_end <= T, "conv_fwd_range: bad frame range [%d,%d) of %d", t_begin, t_end, T);
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fwd_range: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "conv_fwd_range: PReLU needs alpha");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = conv_tc_try(in, w, bias, scale, shift, alpha, act, nullptr, out, B, T, Cin, Fin, Cout, Fout, kt, fstride,
                         in_time_major ? 1 : 0, out_time_major ? 1 : 0, 0, st, nullptr, t_begin, t_end);
    if (rc == 0 && !in_time_major && !out_time_major)
        rc = conv_edge_try(in, w, bias, scale, shift, alpha, act, out, B, T, Cin, Fin, Cout, Fout, kt, fstride, st, nullptr, t_begin, t_end);
    if (rc < 0) { set_error("conv_fwd_range: kernel launch failed"); return rc; }
    CRUSE_CHECK_ARG(rc == 1, "conv_fwd_range: no tensor-core / streaming instantiation for kt=%d fstride=%d Cin=%d Cout=%d Fin=%d (or conv mode is fp32)",
                    kt, fstride, Cin, Cout, Fin);
    return 0;
}

// Eval-mode encoder stage with the (1,3) skip conv of its INPUT fused in (tensor-core instantiations only): one pass over `in`
// produces out = act(BN(conv2x3/s2(in))) and out_skip = conv1x3(in; w_skip) -- model/cruse_net.py:149-152 and :153-155 together.
extern "C" int cruse_conv_skip_fwd(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                                   const float* alpha, int act, const float* w_skip, float* out, float* out_skip, int B, int T,
                                   int Cin, int Fin, int Cout, int Fout, int out_time_major, int t_begin, int t_end, void* stream) {
    CRUSE_CHECK_ARG(in && w && out && w_skip && out_skip, "conv_skip_fwd: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0 && Fout == (Fin + 2 - 3) / 2 + 1, "conv_skip_fwd: bad sizes");
    CRUSE_CHECK_ARG((t_begin == 0 && t_end == 0) || (t_begin >= 0 && t_begin < t_end && t_end <= T), "conv_skip_fwd: bad frame range [%d,%d) of %d", t_begin, t_end, T);
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_skip_fwd: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "conv_skip_fwd: PReLU needs alpha");
    const int rc = conv_skip_tc_try(in, w, bias, scale, shift, alpha, act, w_skip, out, out_skip, B, T, Cin, Fin, Cout, Fout,
                                    out_time_major ? 1 : 0, (cudaStream_t)stream, t_begin, t_end);
    if (rc < 0) { set_error("conv_skip_fwd: kernel launch failed"); return rc; }
    CRUSE_CHECK_ARG(rc == 1, "conv_skip_fwd: no fused tensor-core instantiation for Cin=%d Cout=%d Fin=%d (or conv mode is fp32)", Cin, Cout, Fin);
    return 0;
}

extern "C" int cruse_convT_fwd_range(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                                     const float* alpha, int act, const float* skip, float* out, int B, int T, int Cin, int Fin,
                                     int Cout, int Fout, int t_begin, int t_end, void* stream) {
    CRUSE_CHECK_ARG(in && w && out, "convT_fwd_range: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "convT_fwd_range: bad sizes");
    CRUSE_CHECK_ARG(t_begin >= 0 && t_begin < t_end && t_end <= T, "convT_fwd_range: bad frame range [%d,%d) of %d", t_begin, t_end, T);
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "convT_fwd_range: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "convT_fwd_range: PReLU needs alpha");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = convT_tc_try(in, w, bias, scale, shift, alpha, act, skip, out, B, T, Cin, Fin, Cout, Fout, st, t_begin, t_end);
    if (rc == 0) rc = convT_edge_try(in, w, bias, scale, shift, alpha, act, skip, out, B, T, Cin, Fin, Cout, Fout, st, t_begin, t_end);
    if (rc < 0) { set_error("convT_fwd_range: kernel launch failed"); return rc; }
    CRUSE_CHECK_ARG(rc == 1, "convT_fwd_range: no tensor-core / streaming instantiation for Cin=%d Cout=%d Fin=%d (or conv mode is fp32)", Cin, Cout, Fin);
    return 0;
}

extern "C" int cruse_convT_fwd(const float* in, const float* w, const float* bias, const float* scale,
                               const float* shift, const float* alpha, int act, const float* skip, float* out,
                               float* stats_ws, int B, int T, int Cin, int Fin, int Cout, int Fout, void* stream) {
    CRUSE_CHECK_ARG(in && w && out, "convT_fwd: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "convT_fwd: bad sizes");
    CRUSE_CHECK_ARG(Fout > 0 && Fout <= 2 * Fin + 1, "convT_fwd: Fout=%d must be in (0, 2*Fin+1=%d]", Fout, 2 * Fin + 1);
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "convT_fwd: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "convT_fwd: PReLU needs alpha");
    if (stats_ws && (Fout & 3) == 0 && !scale && act == CRUSE_ACT_NONE && !skip) {
        const int rc = convT_tc_try(in, w, bias, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, nullptr, out, B, T, Cin, Fin, Cout, Fout, (cudaStream_t)stream);
        if (rc < 0) return rc;
        if (rc == 1) {
            launch_bn_stats(out, stats_ws, B, T, Cout, Fout, (cudaStream_t)stream);
            CRUSE_LAUNCH_OK();
            return 0;
        }
    }
    if (!stats_ws) {
        int rc = convT_tc_try(in, w, bias, scale, shift, alpha, act, skip, out, B, T, Cin, Fin, Cout, Fout, (cudaStream_t)stream);
        if (rc) return rc < 0 ? rc : 0;
        rc = convT_edge_try(in, w, bias, scale, shift, alpha, act, skip, out, B, T, Cin, Fin, Cout, Fout, (cudaStream_t)stream);
        if (rc) { if (rc < 0) set_error("convT_fwd: streaming last-stage kernel launch failed"); return rc < 0 ? rc : 0; }
    }
    const int CoutP = (Cout + CONV_COT - 1) / CONV_COT * CONV_COT;
    const size_t smem = sizeof(float) * ((((size_t)CONV_TT * Cin * (Fin + 2) + 3) & ~(size_t)3) + (size_t)Cin * 3 * CoutP + 2 * (size_t)Cout);
    CRUSE_CHECK_ARG(smem <= 227 * 1024, "convT_fwd: stage needs %zu B shared memory", smem);
    CRUSE_CUDA_OK(cudaFuncSetAttribute(convT_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = cruse_conv_nparts(B, T);
    convT_fwd_kernel<1><<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(in, w, bias, scale, shift, alpha, act, skip, out, stats_ws, T,
                                                                         Cin, Fin, Cout, Fout, 0);
    CRUSE_LAUNCH_OK();
    return 0;
}

// ---- data gradients: the forward kernels run "backwards" with re-indexed weights -------------------
extern "C" int cruse_conv_dgrad(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin,
                                int Fin, int Cout, int Fout, int kt, int fstride, void* stream) {
    CRUSE_CHECK_ARG(dz && w && din, "conv_dgrad: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "conv_dgrad: bad sizes");
    CRUSE_CHECK_ARG((kt == 2 && fstride == 2) || (kt == 1 && fstride == 1), "conv_dgrad: supported (kt,fstride) are (2,2) and (1,1), got (%d,%d)", kt, fstride);
    CRUSE_CHECK_ARG(Fout == (Fin + 2 - 3) / fstride + 1, "conv_dgrad: Fout=%d does not match Fin=%d", Fout, Fin);
    const int grid = cruse_conv_nparts(B, T);
    cudaStream_t st = (cudaStream_t)stream;
    if (kt == 1) {
        // din[ci,f] = sum_co sum_kf W[co,ci,0,kf] dz[co,f+1-kf]: a (1,3) conv over dz with flipped taps, channels swapped
        {   // tf32 mode: the same implicit GEMM as the forward skip conv (conv_tc.cu), weights read transposed + flipped
            const int rc = conv_tc_try(dz, w, nullptr, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, addend, din, B, T, Cout, Fout, Cin, Fin, 1, 1, 0, 0, 1, st);
            if (rc) return rc < 0 ? rc : 0;
        }
        const size_t smem = conv_smem_bytes(1, Cout, Fout, Cin);
        CRUSE_CHECK_ARG(smem <= 227 * 1024, "conv_dgrad: stage needs %zu B shared memory", smem);
        CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_fwd_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_kernel<1, 1><<<grid, CONV_THREADS, smem, st>>>(dz, nullptr, w, nullptr, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, addend, din,
                                                                nullptr, T, Cout, Fout, Cin, Fin, 1, 1);
    } else {
        // transposed conv with two time taps (frame t and t+1) over dz, output shifted by the conv's left pad
        {
            const int rc = conv_dgrad_tc_try(dz, w, addend, din, B, T, Cin, Fin, Cout, Fout, kt, st);
            if (rc) return rc < 0 ? rc : 0;
        }
        const int CoutP = (Cin + CONV_COT - 1) / CONV_COT * CONV_COT;
        const size_t smem = sizeof(float) * ((((size_t)(CONV_TT + 1) * Cout * (Fout + 2) + 3) & ~(size_t)3) + (size_t)Cout * 6 * CoutP + 2 * (size_t)Cin);
        CRUSE_CHECK_ARG(smem <= 227 * 1024, "conv_dgrad: stage needs %zu B shared memory", smem);
        CRUSE_CUDA_OK(cudaFuncSetAttribute(convT_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        convT_fwd_kernel<2><<<grid, CONV_THREADS, smem, st>>>(dz, w, nullptr, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, addend, din, nullptr, T,
                                                              Cout, Fout, Cin, Fin, 1);
    }
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_convT_dgrad(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin,
                                 int Fin, int Cout, int Fout, void* stream) {
    CRUSE_CHECK_ARG(dz && w && din, "convT_dgrad: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "convT_dgrad: bad sizes");
    CRUSE_CHECK_ARG(Fout > 0 && Fout <= 2 * Fin + 1 && Fout >= 2 * Fin - 1, "convT_dgrad: Fout=%d must be within [2*Fin-1, 2*Fin+1] (Fin=%d)", Fout, Fin);
    // din[ci,i] = sum_co sum_k W[ci,co,0,k] dz[co,2i+k]: a stride-2 conv over dz without left padding
    {
        int rc = convT_edge_dgrad_try(dz, w, addend, din, B, T, Cin, Fin, Cout, Fout, (cudaStream_t)stream);
        if (rc) { if (rc < 0) set_error("convT_dgrad: streaming last-stage kernel launch failed"); return rc < 0 ? rc : 0; }
        rc = convT_dgrad_tc_try(dz, w, addend, din, B, T, Cin, Fin, Cout, Fout, (cudaStream_t)stream);
        if (rc) return rc < 0 ? rc : 0;
    }
    const size_t smem = conv_smem_bytes(1, Cout, Fout, Cin);
    CRUSE_CHECK_ARG(smem <= 227 * 1024, "convT_dgrad: stage needs %zu B shared memory", smem);
    CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_fwd_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_fwd_kernel<1, 2><<<cruse_conv_nparts(B, T), CONV_THREADS, smem, (cudaStream_t)stream>>>(
        dz, nullptr, w, nullptr, nullptr, nullptr, nullptr, CRUSE_ACT_NONE, addend, din, nullptr, T, Cout, Fout, Cin, Fin, 0, 2);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_bn_finalize(const float* stats_ws, int nparts, int C, double count, const float* gamma,
                                 const float* beta, float eps, float momentum, float* running_mean,
                                 float* running_var, float* scale, float* shift, float* save_mean,
                                 float* save_invstd, void* stream) {
    CRUSE_CHECK_ARG(stats_ws && scale && shift, "bn_finalize: null pointer");
    CRUSE_CHECK_ARG(nparts > 0 && C > 0 && count > 0, "bn_finalize: bad sizes");
    CRUSE_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats go together");
    bn_finalize_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(stats_ws, nparts, C, count, gamma, beta, eps, momentum, running_mean,
                                                            running_var, scale, shift, save_mean, save_invstd);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                             const float* running_var, float eps, float* scale, float* shift, int C, void* stream) {
    CRUSE_CHECK_ARG(running_mean && running_var && scale && shift, "bn_fold: null pointer");
    CRUSE_CHECK_ARG(C > 0, "bn_fold: bad C");
    bn_fold_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps, scale, shift, C);
    CRUSE_LAUNCH_OK();
    return 0;
}

namespace cruse {
struct BnFoldMany {
    const float* gamma[8];
    const float* beta[8];
    const float* mean[8];
    const float* var[8];
    float* scale[8];
    float* shift[8];
    float eps[8];
    int C[8];
};
__global__ void bn_fold_many_kernel(const BnFoldMany a) {
    const int i = blockIdx.x;
    for (int c = threadIdx.x; c < a.C[i]; c += blockDim.x) {
        const float inv = 1.0f / sqrtf(a.var[i][c] + a.eps[i]);
        const float g = a.gamma[i] ? a.gamma[i][c] : 1.f;
        a.scale[i][c] = g * inv;
        a.shift[i][c] = (a.beta[i] ? a.beta[i][c] : 0.f) - a.mean[i][c] * g * inv;
    }
}
}  // namespace cruse

extern "C" int cruse_bn_fold_many(const float* const* gamma, const float* const* beta, const float* const* running_mean,
                                  const float* const* running_var, const float* eps, float* const* scale, float* const* shift,
                                  const int* C, int n, void* stream) {
    CRUSE_CHECK_ARG(running_mean && running_var && scale && shift && eps && C, "bn_fold_many: null pointer");
    CRUSE_CHECK_ARG(n > 0 && n <= 8, "bn_fold_many: n=%d must be in [1, 8]", n);
    cruse::BnFoldMany a;
    for (int i = 0; i < n; ++i) {
        CRUSE_CHECK_ARG(running_mean[i] && running_var[i] && scale[i] && shift[i] && C[i] > 0, "bn_fold_many: bad entry %d", i);
        a.gamma[i] = gamma ? gamma[i] : nullptr;
        a.beta[i] = beta ? beta[i] : nullptr;
        a.mean[i] = running_mean[i];
        a.var[i] = running_var[i];
        a.scale[i] = scale[i];
        a.shift[i] = shift[i];
        a.eps[i] = eps[i];
        a.C[i] = C[i];
    }
    cruse::bn_fold_many_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(a);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_bn_act_fwd(const float* z, const float* scale, const float* shift, const float* alpha, int act,
                                const float* skip, float* y, long long n_frames, int C, int F, void* stream) {
    CRUSE_CHECK_ARG(z && scale && shift && y, "bn_act_fwd: null pointer");
    CRUSE_CHECK_ARG(n_frames > 0 && C > 0 && F > 0, "bn_act_fwd: bad sizes");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "bn_act_fwd: PReLU needs alpha");
    const long long total = n_frames * C * F;
    long long blocks = (total / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    bn_act_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, scale, shift, alpha, act, skip, y, total, C, F);
    CRUSE_LAUNCH_OK();
    return 0;
}

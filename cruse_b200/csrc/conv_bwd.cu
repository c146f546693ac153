// conv_bwd.cu -- weight gradients of the encoder / skip / decoder convolutions (SURVEY.md section 8 row a9;
// on the reference path these are cuDNN wgrad calls made by autograd for model/cruse_net.py:138-143,149-164).
//
//   conv  (kt,3)/stride (1,s):  dW[co,ci,kt,kf] = sum_{b,t,fo} dz[b,t,co,fo] * in[b, t-(KT-1)+kt, ci, s*fo-1+kf]
//   convT (1,3)/stride (1,2):   dW[ci,co,0,k]   = sum_{b,t,i}  dz[b,t,co,2i+k] * in[b,t,ci,i]
//   dbias[co] = sum dz[b,t,co,:]
//
// Same tiling as the forward kernels: a CTA walks chunks of 8 consecutive frames of one utterance
// (persistent, grid-stride), stages the input frames and the dz frames in shared memory with coalesced
// 128-bit loads, and every thread owns a (4 out-channels x 1 in-channel x all taps) block of dW for a
// slice of the chunk's positions.  Per-CTA sums live in shared memory for the whole launch and are
// written once as a partial; cruse_colsum adds the partials in a fixed order (deterministic, no global
// atomics).  The data gradients reuse the forward kernels (conv.cu, cruse_conv_dgrad / cruse_convT_dgrad).
#include "common.cuh"

namespace cruse {

// conv_wgrad_tc.cu: the same reduction as a split-K tcgen05 GEMM (tf32) for the 256-bin pyramid
int conv_wgrad_tc_try(const float* x, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt,
                      int fstride, int pitch, int max_grid, cudaStream_t st);

int convT_wgrad_tc_try(const float* x, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int pitch,
                       int max_grid, cudaStream_t st);

// conv_edge.cu: streaming kernels for the single-channel stages (exact fp32, either numeric mode)
int conv_edge_wgrad_try(const float* in, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt,
                        int fstride, int pitch, int max_grid, cudaStream_t st);
int convT_edge_wgrad_try(const float* in, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int pitch,
                         int max_grid, cudaStream_t st);

constexpr int WG_TT = 8;
constexpr int WG_THREADS = 256;

__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src_base, int t_first, int nrows, int T, int C,
                                           int F, int FP, int left, int tid) {
    // rows t_first .. t_first+nrows-1 of one utterance ([T][C][F]) -> dst[row][c][FP], data at column `left`, rest zero
    const int rowlen = C * F;
    for (int r = 0; r < nrows; ++r) {
        const int t = t_first + r;
        const bool valid = t >= 0 && t < T;
        float* d = dst + (size_t)r * C * FP;
        const float* s = src_base + (size_t)t * rowlen;
        if ((F & 3) == 0) {
            for (int i = tid * 4; i < rowlen; i += WG_THREADS * 4) {
                const float4 v = valid ? __ldg(reinterpret_cast<const float4*>(s + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = i / F, f = i - c * F;
                float* q = d + c * FP + left + f;
                q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
            }
        } else {
            for (int i = tid; i < rowlen; i += WG_THREADS) {
                const int c = i / F, f = i - c * F;
                d[c * FP + left + f] = valid ? __ldg(s + i) : 0.f;
            }
        }
        for (int i = tid; i < C * (FP - F); i += WG_THREADS) {
            const int c = i / (FP - F), j = i - c * (FP - F);
            d[c * FP + (j < left ? j : F + j)] = 0.f;
        }
    }
}

// MODE 0: conv with KT time taps, freq stride SF, left pad 1.  MODE 1: convT (KT = 1, SF ignored).
template <int MODE, int KT, int SF>
__global__ void __launch_bounds__(WG_THREADS)
conv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dz, float* __restrict__ ws, int B, int T, int Cin,
                  int Fin, int Cout, int Fout) {
    extern __shared__ float smem[];
    constexpr int TT = WG_TT, NR = TT + KT - 1, NTAP = KT * 3;
    const int FinP = Fin + 3, FoutP = Fout + 3;     // in: 1 zero column left; dz (convT): zero columns right
    const int CoutP = (Cout + 3) & ~3;
    const int nW = Cout * Cin * NTAP;
    float* s_in = smem;
    float* s_dz = s_in + (((size_t)NR * Cin * FinP + 3) & ~(size_t)3);
    float* s_dw = s_dz + (((size_t)TT * Cout * FoutP + 3) & ~(size_t)3);   // [nW + Cout]
    const int tid = threadIdx.x;
    for (int i = tid; i < nW + Cout; i += WG_THREADS) s_dw[i] = 0.f;

    const int n_items = (CoutP / 4) * Cin;
    const int nsl = n_items >= WG_THREADS ? 1 : WG_THREADS / n_items;   // position slices when there are few outputs
    const int chunks_per_b = (T + TT - 1) / TT;
    const int nchunks = B * chunks_per_b;
    const int npos_f = (MODE == 0) ? Fout : Fin;

    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int b = chunk / chunks_per_b, t0 = (chunk % chunks_per_b) * TT;
        __syncthreads();   // previous chunk's readers are done
        stage_rows(s_in, in + (size_t)b * T * Cin * Fin, t0 - (KT - 1), NR, T, Cin, Fin, FinP, 1, tid);
        stage_rows(s_dz, dz + (size_t)b * T * Cout * Fout, t0, TT, T, Cout, Fout, FoutP, 0, tid);
        __syncthreads();
        const int npos = TT * npos_f;
        for (int w = tid; w < n_items * nsl; w += WG_THREADS) {
            const int item = w % n_items, sl = w / n_items;
            const int ci = item % Cin, cob = item / Cin;
            const int co0 = cob * 4;
            float acc[NTAP][4], accb[4];
#pragma unroll
            for (int k = 0; k < NTAP; ++k)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[k][c] = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) accb[c] = 0.f;
            for (int pos = sl; pos < npos; pos += nsl) {
                const int t = pos / npos_f, f = pos - t * npos_f;
                if (MODE == 0) {
                    float d[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) d[c] = (co0 + c < Cout) ? s_dz[((size_t)t * Cout + co0 + c) * FoutP + f] : 0.f;
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt) {
                        const float* ip = s_in + ((size_t)(t + kt) * Cin + ci) * FinP + SF * f;   // padded col of tap kf=0
#pragma unroll
                        for (int kf = 0; kf < 3; ++kf) {
                            const float x = ip[kf];
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[kt * 3 + kf][c] = fmaf(d[c], x, acc[kt * 3 + kf][c]);
                        }
                    }
                    if (ci == 0) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) accb[c] += d[c];
                    }
                } else {
                    const float x = s_in[((size_t)t * Cin + ci) * FinP + 1 + f];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (co0 + c < Cout) {
                            const float* dp = s_dz + ((size_t)t * Cout + co0 + c) * FoutP + 2 * f;
                            acc[0][c] = fmaf(dp[0], x, acc[0][c]);
                            acc[1][c] = fmaf(dp[1], x, acc[1][c]);
                            acc[2][c] = fmaf(dp[2], x, acc[2][c]);
                            if (ci == 0) accb[c] += dp[0] + dp[1];
                        }
                    }
                }
            }
            // flush to the CTA accumulators
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int co = co0 + c;
                if (co >= Cout) continue;
#pragma unroll
                for (int k = 0; k < NTAP; ++k) {
                    const int idx = (MODE == 0) ? ((co * Cin + ci) * NTAP + k) : ((ci * Cout + co) * 3 + k);
                    if (nsl == 1) s_dw[idx] += acc[k][c]; else atomicAdd(&s_dw[idx], acc[k][c]);
                }
                if (ci == 0) {
                    if (nsl == 1) s_dw[nW + co] += accb[c]; else atomicAdd(&s_dw[nW + co], accb[c]);
                }
            }
        }
    }
    __syncthreads();
    float* o = ws + (size_t)blockIdx.x * (nW + Cout);
    for (int i = tid; i < nW + Cout; i += WG_THREADS) o[i] = s_dw[i];
}

// out[j] (+)= sum_p ws[p*pitch + j], j < n   (double accumulation, fixed order).
// LANES = 1: one thread per column walks the parts (the split-K planes of the GEMMs: a dozen parts, 200 k columns).
// LANES = 8: a block covers 32 columns with 8 part lanes -- lane q adds the parts q, q+8, q+16, ... (four loads in flight), the
// eight lane sums are then added in lane order: for the 300-600 per-CTA partials of the LayerNorm / conv weight-gradient kernels,
// where one thread walking all parts alone took 15-60 us, partly on the dependent chain of the backward pass.
template <int LANES>
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ ws, int nparts, int pitch, int n, float* __restrict__ out,
                                                     int accumulate) {
    constexpr int COLS = 256 / LANES;
    __shared__ double sh[LANES > 1 ? LANES : 1][LANES > 1 ? COLS + 1 : 1];
    const int cx = threadIdx.x % COLS, q = threadIdx.x / COLS;
    const int j = blockIdx.x * COLS + cx;
    double s = 0.0;
    if (j < n) {
        const float* col = ws + j;
        int p = q;
        for (; p + 3 * LANES < nparts; p += 4 * LANES) {
            const float a = col[(size_t)p * pitch], b = col[(size_t)(p + LANES) * pitch], c = col[(size_t)(p + 2 * LANES) * pitch],
                        d = col[(size_t)(p + 3 * LANES) * pitch];
            s += (double)a; s += (double)b; s += (double)c; s += (double)d;
        }
        for (; p < nparts; p += LANES) s += (double)col[(size_t)p * pitch];
    }
    if constexpr (LANES > 1) {
        sh[q][cx] = s;
        __syncthreads();
        if (q != 0) return;
        s = 0.0;
#pragma unroll
        for (int k = 0; k < LANES; ++k) s += sh[k][cx];
    }
    if (j < n) out[j] = (float)(accumulate ? (double)out[j] + s : s);
}

static void launch_colsum(const float* ws, int nparts, int pitch, int n, float* out, int accumulate, cudaStream_t st) {
    if (nparts >= 64) colsum_kernel<8><<<(n + 31) / 32, 256, 0, st>>>(ws, nparts, pitch, n, out, accumulate);
    else colsum_kernel<1><<<(n + 255) / 256, 256, 0, st>>>(ws, nparts, pitch, n, out, accumulate);
}

static size_t wgrad_smem(int KT, int Cin, int Fin, int Cout, int Fout) {
    return sizeof(float) * ((((size_t)(WG_TT + KT - 1) * Cin * (Fin + 3) + 3) & ~(size_t)3) + (((size_t)WG_TT * Cout * (Fout + 3) + 3) & ~(size_t)3) +
                            (size_t)Cout * Cin * KT * 3 + Cout);
}

static int wgrad_grid(int B, int T, size_t smem) {
    const int nchunks = B * ((T + WG_TT - 1) / WG_TT);
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    const int cap = sm_count() * per_sm;
    return nchunks < cap ? nchunks : cap;
}

}  // namespace cruse

using namespace cruse;

extern "C" int cruse_colsum(const float* ws, int nparts, int n, float* out, int accumulate, void* stream) {
    CRUSE_CHECK_ARG(ws && out && nparts > 0 && n > 0, "colsum: bad arguments");
    launch_colsum(ws, nparts, n, n, out, accumulate, (cudaStream_t)stream);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" size_t cruse_conv_wgrad_ws_bytes(int B, int T, int Cin, int Fin, int Cout, int Fout, int kt) {
    const size_t smem = wgrad_smem(kt, Cin, Fin, Cout, Fout);
    return sizeof(float) * (size_t)wgrad_grid(B, T, smem) * ((size_t)Cout * Cin * kt * 3 + Cout);
}

extern "C" int cruse_conv_wgrad(const float* in, const float* dz, float* dw, float* dbias, void* ws, int B, int T, int Cin,
                                int Fin, int Cout, int Fout, int kt, int fstride, void* stream) {
    CRUSE_CHECK_ARG(in && dz && dw && ws, "conv_wgrad: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "conv_wgrad: bad sizes");
    CRUSE_CHECK_ARG((kt == 2 && fstride == 2) || (kt == 1 && fstride == 1), "conv_wgrad: supported (kt,fstride) are (2,2) and (1,1), got (%d,%d)", kt, fstride);
    CRUSE_CHECK_ARG(Fout == (Fin + 2 - 3) / fstride + 1, "conv_wgrad: Fout=%d does not match Fin=%d", Fout, Fin);
    const size_t smem = wgrad_smem(kt, Cin, Fin, Cout, Fout);
    CRUSE_CHECK_ARG(smem <= 227 * 1024, "conv_wgrad: stage needs %zu B shared memory", smem);
    int grid = wgrad_grid(B, T, smem);
    const int nW = Cout * Cin * kt * 3;
    cudaStream_t st = (cudaStream_t)stream;
    int np_tc = conv_edge_wgrad_try(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout, kt, fstride, nW + Cout, grid, st);
    if (np_tc < 0) { set_error("conv_wgrad: streaming stage-1 kernel launch failed"); return np_tc; }
    if (np_tc == 0 && cruse_conv_get_mode() == 1) {       // tf32 mode: tensor-core split-K GEMM; same partial layout, same fixed-order reduction
        np_tc = conv_wgrad_tc_try(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout, kt, fstride, nW + Cout, grid, st);
        if (np_tc < 0) return np_tc;
    }
    if (np_tc > 0) {
        grid = np_tc;
    } else if (kt == 2) {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<0, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_kernel<0, 2, 2><<<grid, WG_THREADS, smem, st>>>(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout);
    } else {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_kernel<0, 1, 1><<<grid, WG_THREADS, smem, st>>>(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout);
    }
    CRUSE_LAUNCH_OK();
    // partial layout per CTA: [nW weights | Cout bias sums]; the two column sums read it with a row pitch of nW+Cout
    {
        const int n = nW + Cout;
        // weights
        launch_colsum((const float*)ws, grid, n, nW, dw, 0, st);
        CRUSE_LAUNCH_OK();
        if (dbias) {
            launch_colsum((const float*)ws + nW, grid, n, Cout, dbias, 0, st);
            CRUSE_LAUNCH_OK();
        }
    }
    return 0;
}

extern "C" size_t cruse_convT_wgrad_ws_bytes(int B, int T, int Cin, int Fin, int Cout, int Fout) {
    const size_t smem = wgrad_smem(1, Cin, Fin, Cout, Fout);
    return sizeof(float) * (size_t)wgrad_grid(B, T, smem) * ((size_t)Cout * Cin * 3 + Cout);
}

extern "C" int cruse_convT_wgrad(const float* in, const float* dz, float* dw, float* dbias, void* ws, int B, int T, int Cin,
                                 int Fin, int Cout, int Fout, void* stream) {
    CRUSE_CHECK_ARG(in && dz && dw && ws, "convT_wgrad: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && Cin > 0 && Cout > 0 && Fin > 0, "convT_wgrad: bad sizes");
    CRUSE_CHECK_ARG(Fout >= 2 * Fin - 2 && Fout <= 2 * Fin && Fout > 0, "convT_wgrad: Fout=%d must be in [2*Fin-2, 2*Fin] (Fin=%d)", Fout, Fin);
    const size_t smem = wgrad_smem(1, Cin, Fin, Cout, Fout);
    CRUSE_CHECK_ARG(smem <= 227 * 1024, "convT_wgrad: stage needs %zu B shared memory", smem);
    int grid = wgrad_grid(B, T, smem);
    const int nW = Cout * Cin * 3, n = nW + Cout;
    cudaStream_t st = (cudaStream_t)stream;
    int np_tc = convT_edge_wgrad_try(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout, n, grid, st);
    if (np_tc < 0) { set_error("convT_wgrad: streaming last-stage kernel launch failed"); return np_tc; }
    if (np_tc == 0 && cruse_conv_get_mode() == 1) {
        np_tc = convT_wgrad_tc_try(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout, n, grid, st);
        if (np_tc < 0) return np_tc;
    }
    if (np_tc > 0) {
        grid = np_tc;
    } else {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<1, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_wgrad_kernel<1, 1, 2><<<grid, WG_THREADS, smem, st>>>(in, dz, (float*)ws, B, T, Cin, Fin, Cout, Fout);
        CRUSE_LAUNCH_OK();
    }
    launch_colsum((const float*)ws, grid, n, nW, dw, 0, st);
    CRUSE_LAUNCH_OK();
    if (dbias) {
        launch_colsum((const float*)ws + nW, grid, n, Cout, dbias, 0, st);
        CRUSE_LAUNCH_OK();
    }
    return 0;
}

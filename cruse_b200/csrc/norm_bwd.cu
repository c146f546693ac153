// norm_bwd.cu -- backward of the normalisation / activation epilogues:
//   * y = act(BatchNorm2d(z)) (+ skip)   model/cruse_net.py:141-142,149-152,161-163   (autograd of BatchNorm2d + ReLU/PReLU)
//   * nn.LayerNorm(D)                    model/cruse_net.py:32-33,46,51
// Both are pure HBM streams with a per-channel (per-column) reduction.  BatchNorm backward needs the
// batch sums before any element can be finished, so it is the classic two passes over (dy, z):
//   pass 1  S1[c] = sum da, S2[c] = sum da*xhat, S3[c] = sum dy*a*[a<=0]   (da = dy*act'(a), a = z*scale+shift)
//   pass 2  dz = A[c] * (da - M1[c] - xhat*M2[c])
// Frames are [C][F] records; a thread always meets the same channels (its offsets inside a frame are
// fixed), so the partial sums live in registers for the whole launch and hit shared memory once.
#include "common.cuh"

namespace cruse {

constexpr int NB_THREADS = 256;
constexpr int NB_MAXIT = 8;          // a frame holds at most NB_MAXIT * 256 * 4 floats (8192) on the vector path

__device__ __forceinline__ float act_grad(float a, float dy, int act, float alpha) {
    switch (act) {
        case CRUSE_ACT_RELU: return a > 0.f ? dy : 0.f;
        case CRUSE_ACT_PRELU: return a > 0.f ? dy : alpha * dy;
        case CRUSE_ACT_SIGMOID: { const float s = sigmoidf_(a); return dy * s * (1.f - s); }
        default: return dy;
    }
}

// VEC = 4 (F % 4 == 0: a float4 never straddles a channel) or 1.  NIT = vector elements of a frame per thread (compile time:
// the per-thread state is NIT-sized register arrays; sized for the largest frame they cost 94 registers = 2 CTAs per SM, and a
// thread had only its own two 16-byte loads in flight -- 2.5 TB/s).  UF frames are loaded together before any is used, so a
// thread of the 1024-float frames of the 256-bin pyramid (NIT = 1) keeps 8 loads in flight.
// Pass 1 ends with a fixed-order reduction: every thread leaves its sums in shared memory indexed by its position in the
// frame, channel c then adds its contiguous range [c*F/VEC, (c+1)*F/VEC) in order (no shared-memory float atomics: those are
// compare-and-swap loops, here up to 32 threads deep on one address, and their order is not fixed).
template <int VEC, bool APPLY, int NIT>
__global__ void __launch_bounds__(NB_THREADS)
bn_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ alpha, int act, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ coef, float* __restrict__ dz,
                  float* __restrict__ partials, long long n_frames, int C, int F) {
    constexpr int UF = NIT == 1 ? 4 : (NIT == 2 ? 2 : 1);
    const int CF = C * F;
    float s1[NIT], s2[NIT], s3[NIT];
    float sc[NIT], sh[NIT], al[NIT], mu[NIT], is[NIT], cA[NIT], cM1[NIT], cM2[NIT];
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
        s1[k] = s2[k] = s3[k] = 0.f;
        const int e = (threadIdx.x + k * NB_THREADS) * VEC;
        const int c = e < CF ? e / F : 0;
        sc[k] = scale ? __ldg(scale + c) : 1.f;
        sh[k] = shift ? __ldg(shift + c) : 0.f;
        al[k] = alpha ? __ldg(alpha + c) : 0.f;
        mu[k] = mean ? __ldg(mean + c) : 0.f;
        is[k] = invstd ? __ldg(invstd + c) : 1.f;
        cA[k] = (APPLY && coef) ? __ldg(coef + c) : 1.f;
        cM1[k] = (APPLY && coef) ? __ldg(coef + C + c) : 0.f;
        cM2[k] = (APPLY && coef) ? __ldg(coef + 2 * C + c) : 0.f;
    }
    const long long stride = gridDim.x;
    for (long long fr0 = blockIdx.x; fr0 < n_frames; fr0 += stride * UF) {
        float dv[UF][NIT][VEC], zv[UF][NIT][VEC];
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            const long long fr = fr0 + u * stride;
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int e = (threadIdx.x + k * NB_THREADS) * VEC;
                if (fr < n_frames && e < CF) {
                    if constexpr (VEC == 4) {
                        const float4 a4 = __ldg(reinterpret_cast<const float4*>(dy + fr * CF + e)), b4 = __ldg(reinterpret_cast<const float4*>(z + fr * CF + e));
                        dv[u][k][0] = a4.x; dv[u][k][1] = a4.y; dv[u][k][2] = a4.z; dv[u][k][3] = a4.w;
                        zv[u][k][0] = b4.x; zv[u][k][1] = b4.y; zv[u][k][2] = b4.z; zv[u][k][3] = b4.w;
                    } else {
                        dv[u][k][0] = __ldg(dy + fr * CF + e);
                        zv[u][k][0] = __ldg(z + fr * CF + e);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UF; ++u) {                       // frames in the order fr0, fr0 + stride, ...: the per-thread sums do not depend on UF
            const long long fr = fr0 + u * stride;
            if (fr >= n_frames) break;
#pragma unroll
            for (int k = 0; k < NIT; ++k) {
                const int e = (threadIdx.x + k * NB_THREADS) * VEC;
                if (e >= CF) continue;
                float ov[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float a = fmaf(zv[u][k][v], sc[k], sh[k]);
                    const float da = act_grad(a, dv[u][k][v], act, al[k]);
                    const float xh = (zv[u][k][v] - mu[k]) * is[k];
                    if (APPLY) {
                        ov[v] = cA[k] * (da - cM1[k] - xh * cM2[k]);
                    } else {
                        s1[k] += da;
                        s2[k] += da * xh;
                        if (act == CRUSE_ACT_PRELU && a <= 0.f) s3[k] += dv[u][k][v] * a;
                    }
                }
                if (APPLY) {
                    if constexpr (VEC == 4) *reinterpret_cast<float4*>(dz + fr * CF + e) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                    else dz[fr * CF + e] = ov[0];
                }
            }
        }
    }
    if (!APPLY) {
        extern __shared__ float s_acc[];   // [3][NIT * NB_THREADS] sums by position in the frame
        constexpr int NPOS = NIT * NB_THREADS;
#pragma unroll
        for (int k = 0; k < NIT; ++k) {
            const int j = threadIdx.x + k * NB_THREADS;
            const bool in = j * VEC < CF;
            s_acc[j] = in ? s1[k] : 0.f;
            s_acc[NPOS + j] = in ? s2[k] : 0.f;
            s_acc[2 * NPOS + j] = in ? s3[k] : 0.f;
        }
        __syncthreads();
        float* o = partials + (size_t)blockIdx.x * 3 * C;
        for (int i = threadIdx.x; i < 3 * C; i += NB_THREADS) {
            const int q = i / C, c = i - q * C;
            const int j0 = (c * F + VEC - 1) / VEC, j1 = ((c + 1) * F + VEC - 1) / VEC;      // VEC == 4: F % 4 == 0
            float a = 0.f;
            for (int j = j0; j < j1; ++j) a += s_acc[q * NPOS + j];
            o[i] = a;
        }
    }
}

template <int VEC, bool APPLY>
static void launch_bn_act_bwd(int grid, cudaStream_t st, const float* dy, const float* z, const float* scale, const float* shift,
                              const float* alpha, int act, const float* mean, const float* invstd, const float* coef, float* dz,
                              float* partials, long long n_frames, int C, int F) {
    const int nit = (C * F + NB_THREADS * VEC - 1) / (NB_THREADS * VEC);
#define CRUSE_BN_BWD(NIT_)                                                                                                            \
    bn_act_bwd_kernel<VEC, APPLY, NIT_><<<grid, NB_THREADS, APPLY ? 0 : sizeof(float) * 3 * NIT_ * NB_THREADS, st>>>(                  \
        dy, z, scale, shift, alpha, act, mean, invstd, coef, dz, partials, n_frames, C, F)
    if (nit <= 1) CRUSE_BN_BWD(1);
    else if (nit <= 2) CRUSE_BN_BWD(2);
    else if (nit <= 4) CRUSE_BN_BWD(4);
    else CRUSE_BN_BWD(8);
#undef CRUSE_BN_BWD
}

// one block per channel: sums the partials (double), emits dgamma/dbeta/dalpha and the pass-2 coefficients
__global__ void __launch_bounds__(128)
bn_bwd_finalize_kernel(const float* __restrict__ partials, int nparts, int C, double count, const float* __restrict__ gamma,
                       const float* __restrict__ invstd, int training, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ dalpha, float* __restrict__ coef) {
    const int c = blockIdx.x;
    double a = 0.0, b = 0.0, d = 0.0;
    for (int p = threadIdx.x; p < nparts; p += blockDim.x) {
        const float* q = partials + (size_t)p * 3 * C;
        a += (double)q[c];
        b += (double)q[C + c];
        d += (double)q[2 * C + c];
    }
    __shared__ double sh[3][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; sh[2][threadIdx.x >> 5] = d; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S1 = 0, S2 = 0, S3 = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { S1 += sh[0][i]; S2 += sh[1][i]; S3 += sh[2][i]; }
        if (dgamma) dgamma[c] = (float)S2;
        if (dbeta) dbeta[c] = (float)S1;
        if (dalpha) dalpha[c] = (float)S3;
        const float g = gamma ? gamma[c] : 1.f;
        coef[c] = g * invstd[c];
        coef[C + c] = training ? (float)(S1 / count) : 0.f;
        coef[2 * C + c] = training ? (float)(S2 / count) : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row.  dx = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*gamma.
// dgamma/dbeta column partials stay in registers (a lane always meets the same columns) and are
// combined across the CTA's warps in shared memory -> one [2*D] partial per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXIT = 8;   // D <= 8*128 = 1024 on the vector path

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                     float* __restrict__ partials, long long rows, int D) {
    extern __shared__ float s_part[];   // [nwarps][2*D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nit = (D + 127) / 128;
    float4 dg[LN_MAXIT], db[LN_MAXIT], gm[LN_MAXIT];
#pragma unroll
    for (int k = 0; k < LN_MAXIT; ++k) {
        dg[k] = db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int i = lane * 4 + k * 128;
        gm[k] = (k < nit && i < D) ? (gamma ? __ldg(reinterpret_cast<const float4*>(gamma + i)) : make_float4(1.f, 1.f, 1.f, 1.f))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long row = (long long)blockIdx.x * nwarps + warp; row < rows; row += (long long)gridDim.x * nwarps) {
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        const float* xr = x + row * D;
        const float* dr = dy + row * D;
        // all 16 loads of the row are issued before the first value is used: with the loads inside the loop below, behind its
        // data-dependent early exit, each 128-column block waited for the previous one's arithmetic (8 round trips per row, 7 us)
        float4 xh[LN_MAXIT], g[LN_MAXIT];
#pragma unroll
        for (int k = 0; k < LN_MAXIT; ++k) {
            const int i = lane * 4 + k * 128;
            const bool in = i < D;                      // i >= k*128, so this also covers k >= nit
            xh[k] = in ? __ldg(reinterpret_cast<const float4*>(xr + i)) : make_float4(mu, mu, mu, mu);
            g[k] = in ? __ldg(reinterpret_cast<const float4*>(dr + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < LN_MAXIT; ++k) {
            const float4 xv = xh[k], dv = g[k];          // out of range: xhat = 0, dy = 0 -> contributes nothing
            xh[k] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            g[k] = make_float4(dv.x * gm[k].x, dv.y * gm[k].y, dv.z * gm[k].z, dv.w * gm[k].w);
            a += (g[k].x + g[k].y) + (g[k].z + g[k].w);
            b += (g[k].x * xh[k].x + g[k].y * xh[k].y) + (g[k].z * xh[k].z + g[k].w * xh[k].w);
            dg[k].x += dv.x * xh[k].x; dg[k].y += dv.y * xh[k].y; dg[k].z += dv.z * xh[k].z; dg[k].w += dv.w * xh[k].w;
            db[k].x += dv.x; db[k].y += dv.y; db[k].z += dv.z; db[k].w += dv.w;
        }
        const float m1 = warp_sum(a) / (float)D, m2 = warp_sum(b) / (float)D;
#pragma unroll
        for (int k = 0; k < LN_MAXIT; ++k) {
            if (k >= nit) break;
            const int i = lane * 4 + k * 128;
            if (i >= D) continue;
            float4 o;
            o.x = rs * (g[k].x - m1 - xh[k].x * m2);
            o.y = rs * (g[k].y - m1 - xh[k].y * m2);
            o.z = rs * (g[k].z - m1 - xh[k].z * m2);
            o.w = rs * (g[k].w - m1 - xh[k].w * m2);
            *reinterpret_cast<float4*>(dx + row * D + i) = o;
        }
    }
    // every warp leaves its column sums in its own [2*D] slice, the CTA adds the slices in warp order (shared-memory float
    // atomics -- compare-and-swap loops, eight warps deep here -- would leave the summation order open)
    float* mine = s_part + (size_t)warp * 2 * D;
#pragma unroll
    for (int k = 0; k < LN_MAXIT; ++k) {
        if (k >= nit) break;
        const int i = lane * 4 + k * 128;
        if (i >= D) continue;
        *reinterpret_cast<float4*>(mine + i) = dg[k];
        *reinterpret_cast<float4*>(mine + D + i) = db[k];
    }
    __syncthreads();
    float* o = partials + (size_t)blockIdx.x * 2 * D;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
        float a = 0.f;
        for (int w = 0; w < nwarps; ++w) a += s_part[(size_t)w * 2 * D + i];
        o[i] = a;
    }
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s = __ldg(y + i);
        dz[i] = __ldg(dy + i) * s * (1.f - s);
    }
}

static int nb_grid(long long n_frames) {
    long long g = (long long)sm_count() * 4;
    if (g > n_frames) g = n_frames;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace cruse

using namespace cruse;

extern "C" int cruse_bn_bwd_nparts(long long n_frames) { return nb_grid(n_frames); }

extern "C" int cruse_bn_act_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift,
                                       const float* alpha, int act, const float* mean, const float* invstd, float* partials,
                                       long long n_frames, int C, int F, void* stream) {
    CRUSE_CHECK_ARG(dy && z && partials, "bn_act_bwd_reduce: null pointer");
    CRUSE_CHECK_ARG(n_frames > 0 && C > 0 && F > 0, "bn_act_bwd_reduce: bad sizes");
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "bn_act_bwd_reduce: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "bn_act_bwd_reduce: PReLU needs alpha");
    const int vec = (F & 3) == 0 ? 4 : 1;
    CRUSE_CHECK_ARG((long long)C * F <= (long long)NB_MAXIT * NB_THREADS * vec, "bn_act_bwd_reduce: frame of %d x %d floats is too large", C, F);
    const int grid = nb_grid(n_frames);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec == 4)
        launch_bn_act_bwd<4, false>(grid, st, dy, z, scale, shift, alpha, act, mean, invstd, nullptr, nullptr, partials, n_frames, C, F);
    else
        launch_bn_act_bwd<1, false>(grid, st, dy, z, scale, shift, alpha, act, mean, invstd, nullptr, nullptr, partials, n_frames, C, F);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_bn_bwd_finalize(const float* partials, int nparts, int C, double count, const float* gamma,
                                     const float* invstd, int training, float* dgamma, float* dbeta, float* dalpha,
                                     float* coef, void* stream) {
    CRUSE_CHECK_ARG(partials && invstd && coef, "bn_bwd_finalize: null pointer");
    CRUSE_CHECK_ARG(nparts > 0 && C > 0 && count > 0, "bn_bwd_finalize: bad sizes");
    bn_bwd_finalize_kernel<<<C, 128, 0, (cudaStream_t)stream>>>(partials, nparts, C, count, gamma, invstd, training, dgamma, dbeta, dalpha, coef);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_bn_act_bwd_apply(const float* dy, const float* z, const float* scale, const float* shift,
                                      const float* alpha, int act, const float* mean, const float* invstd, const float* coef,
                                      float* dz, long long n_frames, int C, int F, void* stream) {
    CRUSE_CHECK_ARG(dy && z && dz, "bn_act_bwd_apply: null pointer");
    CRUSE_CHECK_ARG(n_frames > 0 && C > 0 && F > 0, "bn_act_bwd_apply: bad sizes");
    CRUSE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "bn_act_bwd_apply: scale and shift go together");
    CRUSE_CHECK_ARG(act != CRUSE_ACT_PRELU || alpha, "bn_act_bwd_apply: PReLU needs alpha");
    const int vec = (F & 3) == 0 ? 4 : 1;
    CRUSE_CHECK_ARG((long long)C * F <= (long long)NB_MAXIT * NB_THREADS * vec, "bn_act_bwd_apply: frame of %d x %d floats is too large", C, F);
    long long g = (long long)sm_count() * 8;
    if (g > n_frames) g = n_frames;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec == 4)
        launch_bn_act_bwd<4, true>((int)g, st, dy, z, scale, shift, alpha, act, mean, invstd, coef, dz, nullptr, n_frames, C, F);
    else
        launch_bn_act_bwd<1, true>((int)g, st, dy, z, scale, shift, alpha, act, mean, invstd, coef, dz, nullptr, n_frames, C, F);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_sigmoid_bwd(const float* dy, const float* y, float* dz, long long n, void* stream) {
    CRUSE_CHECK_ARG(dy && y && dz && n > 0, "sigmoid_bwd: bad arguments");
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    sigmoid_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, dz, n);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_layernorm_bwd_nparts(long long rows) {
    long long g = (long long)sm_count() * 2;
    const long long need = (rows + 7) / 8;
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" int cruse_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                   float* dx, float* partials, long long rows, int D, void* stream) {
    CRUSE_CHECK_ARG(dy && x && mean && rstd && dx && partials, "layernorm_bwd: null pointer");
    CRUSE_CHECK_ARG(rows > 0 && D > 0 && (D % 4) == 0 && D <= LN_MAXIT * 128, "layernorm_bwd: bad sizes rows=%lld D=%d (D%%4==0, D<=%d)", rows, D, LN_MAXIT * 128);
    const int grid = cruse_layernorm_bwd_nparts(rows);
    const size_t smem = sizeof(float) * 8 * 2 * D;          // one [2*D] slice per warp
    CRUSE_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layernorm_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dy, x, gamma, mean, rstd, dx, partials, rows, D);
    CRUSE_LAUNCH_OK();
    return 0;
}

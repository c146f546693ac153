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

// VEC = 4 (F % 4 == 0: a float4 never straddles a channel) or 1
template <int VEC, bool APPLY>
__global__ void __launch_bounds__(NB_THREADS)
bn_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ alpha, int act, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ coef, float* __restrict__ dz,
                  float* __restrict__ partials, long long n_frames, int C, int F) {
    const int CF = C * F;
    const int nit = (CF + NB_THREADS * VEC - 1) / (NB_THREADS * VEC);
    float s1[NB_MAXIT], s2[NB_MAXIT], s3[NB_MAXIT];
    float sc[NB_MAXIT], sh[NB_MAXIT], al[NB_MAXIT], mu[NB_MAXIT], is[NB_MAXIT], cA[NB_MAXIT], cM1[NB_MAXIT], cM2[NB_MAXIT];
#pragma unroll
    for (int k = 0; k < NB_MAXIT; ++k) {
        s1[k] = s2[k] = s3[k] = 0.f;
        const int e = (threadIdx.x + k * NB_THREADS) * VEC;
        const int c = (k < nit && e < CF) ? e / F : 0;
        sc[k] = scale ? __ldg(scale + c) : 1.f;
        sh[k] = shift ? __ldg(shift + c) : 0.f;
        al[k] = alpha ? __ldg(alpha + c) : 0.f;
        mu[k] = mean ? __ldg(mean + c) : 0.f;
        is[k] = invstd ? __ldg(invstd + c) : 1.f;
        cA[k] = (APPLY && coef) ? __ldg(coef + c) : 1.f;
        cM1[k] = (APPLY && coef) ? __ldg(coef + C + c) : 0.f;
        cM2[k] = (APPLY && coef) ? __ldg(coef + 2 * C + c) : 0.f;
    }
    for (long long fr = blockIdx.x; fr < n_frames; fr += gridDim.x) {
        const float* dyf = dy + fr * CF;
        const float* zf = z + fr * CF;
#pragma unroll
        for (int k = 0; k < NB_MAXIT; ++k) {
            if (k >= nit) break;
            const int e = (threadIdx.x + k * NB_THREADS) * VEC;
            if (e >= CF) continue;
            float dv[VEC], zv[VEC], ov[VEC];
            if constexpr (VEC == 4) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(dyf + e)), b4 = __ldg(reinterpret_cast<const float4*>(zf + e));
                dv[0] = a4.x; dv[1] = a4.y; dv[2] = a4.z; dv[3] = a4.w;
                zv[0] = b4.x; zv[1] = b4.y; zv[2] = b4.z; zv[3] = b4.w;
            } else {
                dv[0] = __ldg(dyf + e);
                zv[0] = __ldg(zf + e);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float a = fmaf(zv[v], sc[k], sh[k]);
                const float da = act_grad(a, dv[v], act, al[k]);
                const float xh = (zv[v] - mu[k]) * is[k];
                if (APPLY) {
                    ov[v] = cA[k] * (da - cM1[k] - xh * cM2[k]);
                } else {
                    s1[k] += da;
                    s2[k] += da * xh;
                    if (act == CRUSE_ACT_PRELU && a <= 0.f) s3[k] += dv[v] * a;
                }
            }
            if (APPLY) {
                if constexpr (VEC == 4) *reinterpret_cast<float4*>(dz + fr * CF + e) = make_float4(ov[0], ov[1], ov[2], ov[3]);
                else dz[fr * CF + e] = ov[0];
            }
        }
    }
    if (!APPLY) {
        extern __shared__ float s_acc[];   // [3*C]
        for (int i = threadIdx.x; i < 3 * C; i += NB_THREADS) s_acc[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NB_MAXIT; ++k) {
            if (k >= nit) break;
            const int e = (threadIdx.x + k * NB_THREADS) * VEC;
            if (e >= CF) continue;
            const int c = e / F;
            atomicAdd(&s_acc[c], s1[k]);
            atomicAdd(&s_acc[C + c], s2[k]);
            if (act == CRUSE_ACT_PRELU) atomicAdd(&s_acc[2 * C + c], s3[k]);
        }
        __syncthreads();
        float* o = partials + (size_t)blockIdx.x * 3 * C;
        for (int i = threadIdx.x; i < 3 * C; i += NB_THREADS) o[i] = s_acc[i];
    }
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
    extern __shared__ float s_part[];   // [2*D]
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
        float4 xh[LN_MAXIT], g[LN_MAXIT];
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < LN_MAXIT; ++k) {
            if (k >= nit) break;
            const int i = lane * 4 + k * 128;
            if (i >= D) { xh[k] = g[k] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
            const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + i)), dv = __ldg(reinterpret_cast<const float4*>(dr + i));
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
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) s_part[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LN_MAXIT; ++k) {
        if (k >= nit) break;
        const int i = lane * 4 + k * 128;
        if (i >= D) continue;
        atomicAdd(&s_part[i + 0], dg[k].x); atomicAdd(&s_part[i + 1], dg[k].y); atomicAdd(&s_part[i + 2], dg[k].z); atomicAdd(&s_part[i + 3], dg[k].w);
        atomicAdd(&s_part[D + i + 0], db[k].x); atomicAdd(&s_part[D + i + 1], db[k].y); atomicAdd(&s_part[D + i + 2], db[k].z); atomicAdd(&s_part[D + i + 3], db[k].w);
    }
    __syncthreads();
    float* o = partials + (size_t)blockIdx.x * 2 * D;
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) o[i] = s_part[i];
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
    const size_t smem = sizeof(float) * 3 * C;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec == 4)
        bn_act_bwd_kernel<4, false><<<grid, NB_THREADS, smem, st>>>(dy, z, scale, shift, alpha, act, mean, invstd, nullptr, nullptr, partials, n_frames, C, F);
    else
        bn_act_bwd_kernel<1, false><<<grid, NB_THREADS, smem, st>>>(dy, z, scale, shift, alpha, act, mean, invstd, nullptr, nullptr, partials, n_frames, C, F);
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
        bn_act_bwd_kernel<4, true><<<(int)g, NB_THREADS, 0, st>>>(dy, z, scale, shift, alpha, act, mean, invstd, coef, dz, nullptr, n_frames, C, F);
    else
        bn_act_bwd_kernel<1, true><<<(int)g, NB_THREADS, 0, st>>>(dy, z, scale, shift, alpha, act, mean, invstd, coef, dz, nullptr, n_frames, C, F);
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
    layernorm_bwd_kernel<<<grid, 256, sizeof(float) * 2 * D, (cudaStream_t)stream>>>(dy, x, gamma, mean, rstd, dx, partials, rows, D);
    CRUSE_LAUNCH_OK();
    return 0;
}

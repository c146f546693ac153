// frontend.cu -- the step in FRONT of the hot path: input feature norms and on-the-fly mixing on the device,
// so that an 8-GPU job is not fed by a CPU dataloader.
//
//   cruse_feature_norm   train_base/model/base_model.py:202-300  offline_laplace / cumulative_laplace / offline_gaussian /
//                        cumulative_layer norms of a magnitude spectrogram (reference layout [B,C,F,T]; here frame-major [B,T,F],
//                        C = 1: the running statistics are over all bins of the frames 0..t)
//   cruse_rir_conv       dataset/dataset.py:244-247  scipy.signal.fftconvolve(y, rir)[:len(y)] as a direct causal convolution
//   cruse_snr_mix        dataset/dataset.py:236-264  peak normalisation, SNR scalar, sum, output level
//
// All of them are HBM streams with per-utterance reductions; reductions have a fixed order (deterministic).
#include "common.cuh"

namespace cruse {
namespace {

// ---------------------------------------------------------------------------------------------------------------------
// feature norms: one CTA per utterance.  (a) per-frame sums (one warp per frame), (b) running totals over time by warp 0 (double),
// (c) the normalisation pass (the input is re-read: it was just streamed through L2).
// mode 0 offline_laplace, 1 cumulative_laplace, 2 offline_gaussian, 3 cumulative_layer
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
feature_norm_kernel(const float* __restrict__ x, float* __restrict__ y, int T, int F, int mode, float eps_cum) {
    extern __shared__ float sm[];
    float* a = sm;            // [T]  sum x of frame t      -> mean (or 0) to subtract
    float* q = sm + T;        // [T]  sum x^2 of frame t    -> factor to multiply with
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float* xb = x + (size_t)b * T * F;
    float* yb = y + (size_t)b * T * F;
    for (int t = warp; t < T; t += nwarps) {
        float s = 0.f, s2 = 0.f;
        for (int f = lane; f < F; f += 32) {
            const float v = __ldg(xb + (size_t)t * F + f);
            s += v;
            s2 = fmaf(v, v, s2);
        }
        s = warp_sum(s);
        s2 = warp_sum(s2);
        if (lane == 0) { a[t] = s; q[t] = s2; }
    }
    __syncthreads();
    if (warp == 0) {
        double cs = 0.0, cp = 0.0;                       // running totals carried between groups of 32 frames
        if (mode == 0 || mode == 2) {                    // offline: utterance-level statistics
            double s = 0.0, s2 = 0.0;
            for (int t = lane; t < T; t += 32) { s += (double)a[t]; s2 += (double)q[t]; }
            for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            const double n = (double)T * F, mu = s / n;
            float sub, mul;
            if (mode == 0) { sub = 0.f; mul = (float)(1.0 / (mu + 1e-5)); }                      // :213
            else {
                const double var = (s2 - n * mu * mu) / (n - 1.0);                                // torch.std: unbiased (:257)
                sub = (float)mu;
                mul = (float)(1.0 / (sqrt(var > 0.0 ? var : 0.0) + 1e-5));
            }
            __syncwarp();
            for (int t = lane; t < T; t += 32) { a[t] = sub; q[t] = mul; }
        } else {
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int t = t0 + lane;
                double s = t < T ? (double)a[t] : 0.0, s2 = t < T ? (double)q[t] : 0.0;
                for (int o = 1; o < 32; o <<= 1) {       // inclusive scan of the 32 frames
                    const double u = __shfl_up_sync(0xffffffffu, s, o), u2 = __shfl_up_sync(0xffffffffu, s2, o);
                    if (lane >= o) { s += u; s2 += u2; }
                }
                s += cs;
                s2 += cp;
                cs = __shfl_sync(0xffffffffu, s, 31);
                cp = __shfl_sync(0xffffffffu, s2, 31);
                if (t < T) {
                    const double cnt = (double)F * (t + 1), mean = s / cnt;
                    if (mode == 1) { a[t] = 0.f; q[t] = (float)(1.0 / (mean + (double)eps_cum)); }                 // :240-243
                    else {
                        const double var = (s2 - 2.0 * mean * s) / cnt + mean * mean;                               // :292, literally
                        a[t] = (float)mean;
                        q[t] = (float)(1.0 / sqrt(var + (double)eps_cum));                                          // :293
                    }
                }
            }
        }
    }
    __syncthreads();
    const long long n = (long long)T * F;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const int t = (int)(i / F);
        yb[i] = (__ldg(xb + i) - a[t]) * q[t];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// y[b, n] = sum_{k < R} rir[b or 0][k] * x[b, n - k]   (n < L): fftconvolve(x, rir)[:L].  A thread owns 8 consecutive outputs and
// slides a register window over the input while it walks the impulse response: one shared-memory read of x and one (broadcast) of
// the tap per 8 FMAs.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RC_TN = 2048;      // outputs per CTA (256 threads x 8)
constexpr int RC_RC = 1024;      // taps per pass

__global__ void __launch_bounds__(256)
rir_conv_kernel(const float* __restrict__ x, const float* __restrict__ rir, float* __restrict__ y, int L, int R, long long rir_stride) {
    __shared__ float xs[RC_TN + RC_RC];
    __shared__ float rs[RC_RC];
    const int b = blockIdx.y, n0 = blockIdx.x * RC_TN;
    const float* xb = x + (size_t)b * L;
    const float* rb = rir + (size_t)b * rir_stride;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int o0 = 8 * threadIdx.x;                                   // my outputs: n0 + o0 + j
    for (int k0 = 0; k0 < R; k0 += RC_RC) {
        // the input samples this pass touches: n - k for n in [n0, n0 + TN), k in [k0, k0 + RC)  ->  [n0 - k0 - RC + 1, n0 - k0 + TN)
        const int base = n0 - k0 - RC_RC + 1;
        __syncthreads();
        for (int i = threadIdx.x; i < RC_TN + RC_RC - 1; i += blockDim.x) {
            const int m = base + i;
            xs[i] = (m >= 0 && m < L) ? __ldg(xb + m) : 0.f;
        }
        for (int i = threadIdx.x; i < RC_RC; i += blockDim.x) rs[i] = (k0 + i < R) ? __ldg(rb + k0 + i) : 0.f;
        __syncthreads();
        // output n0 + o0 + j with tap k0 + kk reads xs[(o0 + j) - kk + RC - 1]
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = xs[o0 + j + RC_RC - 1];     // kk = 0
        for (int kk = 0; kk < RC_RC; ++kk) {
            const float r = rs[kk];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(r, w[j], acc[j]);
#pragma unroll
            for (int j = 7; j > 0; --j) w[j] = w[j - 1];
            w[0] = (kk + 1 < RC_RC) ? xs[o0 - (kk + 1) + RC_RC - 1] : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (n0 + o0 + j < L) y[(size_t)b * L + n0 + o0 + j] = acc[j];
}

// ---------------------------------------------------------------------------------------------------------------------
// snr_mix: pass 1 = per-(utterance, chunk) partials (max|c|, max|n|, sum cc, sum nn, sum cn); finish = the three scalars of the
// utterance; pass 2 = the two output streams
// ---------------------------------------------------------------------------------------------------------------------
constexpr int MIX_CHUNKS = 32;

__global__ void __launch_bounds__(256)
snr_mix_partial_kernel(const float* __restrict__ clean, const float* __restrict__ noise, float* __restrict__ part, int L) {
    __shared__ float sh[5][8];
    const int b = blockIdx.y, j = blockIdx.x;
    const int len = (L + MIX_CHUNKS - 1) / MIX_CHUNKS;
    const int lo = j * len, hi = (lo + len < L) ? lo + len : L;
    const float* c = clean + (size_t)b * L;
    const float* n = noise + (size_t)b * L;
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float cv = __ldg(c + i), nv = __ldg(n + i);
        v[0] = fmaxf(v[0], fabsf(cv)); v[1] = fmaxf(v[1], fabsf(nv));
        v[2] = fmaf(cv, cv, v[2]); v[3] = fmaf(nv, nv, v[3]); v[4] = fmaf(cv, nv, v[4]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        float r = v[q];
        for (int o = 16; o > 0; o >>= 1) {
            const float u = __shfl_xor_sync(0xffffffffu, r, o);
            r = q < 2 ? fmaxf(r, u) : r + u;
        }
        if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = r;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float r = 0.f;
        for (int w = 0; w < 8; ++w) r = threadIdx.x < 2 ? fmaxf(r, sh[threadIdx.x][w]) : r + sh[threadIdx.x][w];
        part[((size_t)b * MIX_CHUNKS + j) * 5 + threadIdx.x] = r;
    }
}

// coef[b] = (a, g, sc): clean_out = sc * a * clean, noisy_out = sc * (a * clean + g * noise)
__global__ void __launch_bounds__(256)
snr_mix_finish_kernel(const float* __restrict__ part, const float* __restrict__ snr_db, const float* __restrict__ level_db,
                      float* __restrict__ coef, int B, int L, float eps_f) {
    const double eps = (double)eps_f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double pc = 0.0, pn = 0.0, cc = 0.0, nn = 0.0, cn = 0.0;
        for (int j = 0; j < MIX_CHUNKS; ++j) {
            const float* p = part + ((size_t)b * MIX_CHUNKS + j) * 5;
            pc = fmax(pc, (double)p[0]); pn = fmax(pn, (double)p[1]);
            cc += (double)p[2]; nn += (double)p[3]; cn += (double)p[4];
        }
        const double a = 1.0 / (pc + eps), bn = 1.0 / (pn + eps);                     // dataset.py:250,254
        const double clean_rms = sqrt(a * a * cc / L), noise_rms = sqrt(bn * bn * nn / L);   // :252,256
        const double g = bn * (clean_rms / pow(10.0, (double)snr_db[b] / 20.0) / (noise_rms + eps));   // :257-258
        double sc = 1.0;
        if (level_db) {
            const double noisy_rms = sqrt((a * a * cc + 2.0 * a * g * cn + g * g * nn) / L);
            sc = pow(10.0, (double)level_db[b] / 20.0) / (noisy_rms + eps);
        }
        coef[3 * b] = (float)a; coef[3 * b + 1] = (float)g; coef[3 * b + 2] = (float)sc;
    }
}

__global__ void __launch_bounds__(256)
snr_mix_apply_kernel(const float* __restrict__ clean, const float* __restrict__ noise, const float* __restrict__ coef,
                     float* __restrict__ noisy_out, float* __restrict__ clean_out, long long total, int L) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / L;
        const float a = __ldg(coef + 3 * b), g = __ldg(coef + 3 * b + 1), sc = __ldg(coef + 3 * b + 2);
        const float c = a * __ldg(clean + i);
        noisy_out[i] = sc * fmaf(g, __ldg(noise + i), c);
        if (clean_out) clean_out[i] = sc * c;
    }
}

}  // namespace
}  // namespace cruse

using namespace cruse;

extern "C" int cruse_feature_norm(const float* x, float* y, int B, int T, int F, int mode, void* stream) {
    CRUSE_CHECK_ARG(x && y, "feature_norm: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && F > 0 && mode >= 0 && mode <= 3, "feature_norm: bad sizes / mode B=%d T=%d F=%d mode=%d", B, T, F, mode);
    const size_t smem = sizeof(float) * 2 * (size_t)T;
    CRUSE_CHECK_ARG(smem <= 200 * 1024, "feature_norm: T=%d frames do not fit the per-utterance scan (<= 25600)", T);
    CRUSE_CUDA_OK(cudaFuncSetAttribute(feature_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    feature_norm_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(x, y, T, F, mode, 1.1920929e-07f);   // train_base/constant.py:8
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_rir_conv(const float* x, const float* rir, float* y, int B, int L, int R, long long rir_stride, void* stream) {
    CRUSE_CHECK_ARG(x && rir && y && x != y, "rir_conv: null pointer / in-place call");
    CRUSE_CHECK_ARG(B > 0 && L > 0 && R > 0 && (rir_stride == 0 || rir_stride >= R), "rir_conv: bad sizes B=%d L=%d R=%d stride=%lld", B, L, R, rir_stride);
    rir_conv_kernel<<<dim3((L + RC_TN - 1) / RC_TN, B), 256, 0, (cudaStream_t)stream>>>(x, rir, y, L, R, rir_stride);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" size_t cruse_snr_mix_ws_bytes(int B) { return sizeof(float) * (size_t)(B > 0 ? B : 0) * (5 * MIX_CHUNKS + 3); }

extern "C" int cruse_snr_mix(const float* clean, const float* noise, const float* snr_db, const float* level_db, float* noisy_out,
                             float* clean_out, void* ws, int B, int L, float eps, void* stream) {
    CRUSE_CHECK_ARG(clean && noise && snr_db && noisy_out && ws, "snr_mix: null pointer");
    CRUSE_CHECK_ARG(B > 0 && L > 0, "snr_mix: bad sizes B=%d L=%d", B, L);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = static_cast<float*>(ws);
    float* coef = part + (size_t)B * MIX_CHUNKS * 5;
    snr_mix_partial_kernel<<<dim3(MIX_CHUNKS, B), 256, 0, st>>>(clean, noise, part, L);
    CRUSE_LAUNCH_OK();
    snr_mix_finish_kernel<<<1, 256, 0, st>>>(part, snr_db, level_db, coef, B, L, eps);
    CRUSE_LAUNCH_OK();
    const long long total = (long long)B * L;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    snr_mix_apply_kernel<<<(unsigned)blocks, 256, 0, st>>>(clean, noise, coef, noisy_out, clean_out, total, L);
    CRUSE_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// 16-bit PCM -> float32 on the device: out = in / 32768, the conversion soundfile / librosa apply when the reference reads a wav file
// (dataset/dataset.py:20, train_base/acoustics/feature.py:110-114).  The HOST-buffer entry points take the samples in the files' own
// format, so a step moves half the bytes over PCIe (the limit of an 8-GPU box: DESIGN.md section 6).  Eight samples per thread.
// ---------------------------------------------------------------------------------------------
namespace cruse {
__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const short* __restrict__ in, float* __restrict__ out, long long n, float scale) {
    const long long n8 = n >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(in) + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float4 o[2];
        float* of = reinterpret_cast<float*>(o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            of[2 * k] = (float)(short)(w[k] & 0xffff) * scale;
            of[2 * k + 1] = (float)(short)(w[k] >> 16) * scale;
        }
        reinterpret_cast<float4*>(out)[2 * i] = o[0];
        reinterpret_cast<float4*>(out)[2 * i + 1] = o[1];
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) out[(n8 << 3) + threadIdx.x] = (float)in[(n8 << 3) + threadIdx.x] * scale;
}
}  // namespace cruse

extern "C" int cruse_pcm16_to_float(const short* in, float* out, long long n, void* stream) {
    using namespace cruse;
    CRUSE_CHECK_ARG(in && out && n > 0, "pcm16_to_float: null pointer / empty input");
    CRUSE_CHECK_ARG(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "pcm16_to_float: buffers must be 16-byte aligned");
    long long blocks = ((n >> 3) + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pcm16_to_float_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n, 1.0f / 32768.0f);
    CRUSE_LAUNCH_OK();
    return 0;
}

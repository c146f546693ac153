// spectral.cu -- framed real FFT front-end (STFT) and mask*spectrum + overlap-add iSTFT back-end.
//
// Replaces torch.stft / torch.istft at train_base/acoustics/feature.py:22-30,53-61 and
// utils/utils.py:390-400,417-433,448-454 of the reference.
//
// Design (HBM-bound stages, SURVEY.md App. B: 4.4 KB / 3.3 KB per frame):
//  * one warp owns one frame: hop-strided samples are read coalesced (float2), windowed,
//    packed as N/2 complex points, transformed by a mixed-radix (4,2,5,3) Stockham FFT that
//    lives entirely in that warp's shared-memory ping-pong buffers (only __syncwarp between
//    passes), then split into the N/2+1 real-FFT bins and written as coalesced float2.
//  * the magnitude the network consumes (utils/utils.py:400) is emitted by the same kernel.
//  * iSTFT: a CTA owns a run of consecutive frames of one utterance (+ halo frames that are
//    recomputed), so the overlap-add is a shared-memory gather with no atomics; the mask
//    multiply (PreProcess.masking) is fused into the spectrum load and the hann^2 envelope
//    division into the store.
#include "common.cuh"

namespace cruse {

struct FftPlan {
    int nfac;
    int fac[16];
};

static int make_plan(int M, FftPlan& p) {
    p.nfac = 0;
    int n = M;
    while (n % 4 == 0) { p.fac[p.nfac++] = 4; n /= 4; }
    while (n % 2 == 0) { p.fac[p.nfac++] = 2; n /= 2; }
    while (n % 5 == 0) { p.fac[p.nfac++] = 5; n /= 5; }
    while (n % 3 == 0) { p.fac[p.nfac++] = 3; n /= 3; }
    return n == 1 ? 0 : -1;
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// tw[j] = exp(-2*pi*i*j/N), j in [0,N)
__device__ __forceinline__ void build_twiddles(float2* tw, int N) {
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double s, c;
        sincospi(2.0 * (double)j / (double)N, &s, &c);
        tw[j] = make_float2((float)c, (float)(-s));
    }
}

// Complex FFT of size M executed by one warp on shared-memory buffers x -> (x|y).
// Stockham autosort, decimation in frequency:  for radix r, n_cur = r*m, stride s:
//   y[q + s*(r*p + k)] = W_{n_cur}^{p*k} * sum_j x[q + s*(p + m*j)] * W_r^{j*k}
// INV conjugates every root (unnormalised inverse).  Returns the buffer holding the result.
template <bool INV>
__device__ float2* warp_fft(float2* x, float2* y, const float2* __restrict__ tw, int M, const FftPlan& plan, int lane) {
    const int N = 2 * M;
    int n_cur = M, s = 1;
    for (int f = 0; f < plan.nfac; ++f) {
        const int r = plan.fac[f];
        const int m = n_cur / r;
        const int twstep = N / n_cur;
        const int nb = M / r;
        for (int i = lane; i < nb; i += 32) {
            const int p = i / s, q = i - p * s;
            const float2* xi = x + q + s * p;
            float2* yo = y + q + s * r * p;
            const int sm = s * m;
            if (r == 4) {
                float2 a0 = xi[0], a1 = xi[sm], a2 = xi[2 * sm], a3 = xi[3 * sm];
                float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
                float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // d * (+-i)
                float2 b0 = cadd(t0, t2), b1 = cadd(t1, t3), b2 = csub(t0, t2), b3 = csub(t1, t3);
                float2 w1 = tw[p * twstep], w2 = tw[2 * p * twstep], w3 = tw[3 * p * twstep];
                if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
                yo[0] = b0;
                yo[s] = cmul(b1, w1);
                yo[2 * s] = cmul(b2, w2);
                yo[3 * s] = cmul(b3, w3);
            } else if (r == 2) {
                float2 a0 = xi[0], a1 = xi[sm];
                float2 w1 = tw[p * twstep];
                if (INV) w1.y = -w1.y;
                yo[0] = cadd(a0, a1);
                yo[s] = cmul(csub(a0, a1), w1);
            } else {  // r = 3 or 5: direct small DFT through the root table (r divides N)
                float2 a[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) a[j] = (j < r) ? xi[j * sm] : make_float2(0.f, 0.f);
                const int rstep = N / r;
                for (int k = 0; k < r; ++k) {
                    float2 acc = a[0];
#pragma unroll
                    for (int j = 1; j < 5; ++j) {
                        if (j < r) {
                            float2 w = tw[((j * k) % r) * rstep];
                            if (INV) w.y = -w.y;
                            acc = cadd(acc, cmul(a[j], w));
                        }
                    }
                    float2 wk = tw[p * k * twstep];
                    if (INV) wk.y = -wk.y;
                    yo[k * s] = cmul(acc, wk);
                }
            }
        }
        __syncwarp();
        float2* tmp = x; x = y; y = tmp;
        n_cur = m;
        s *= r;
    }
    return x;
}

// The same transform specialised for M = 256 = 4^4 (n_fft = 512, the geometry of the benchmark): four radix-4 passes
// with compile-time strides -- no divisions, the four inputs of a butterfly are always 64 apart, the last pass needs no
// twiddles.  Result is back in x (even number of ping-pong passes).
template <bool INV, int S>
__device__ __forceinline__ void fft256_pass(const float2* __restrict__ x, float2* __restrict__ y, const float2* __restrict__ tw, int lane) {
    constexpr int TWSTEP = 2 * S;                      // N / n_cur = 512 / (256 / S)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        const int p = i / S, q = i % S;                // S is a power of two: shift / mask
        const float2 a0 = x[i], a1 = x[i + 64], a2 = x[i + 128], a3 = x[i + 192];
        const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
        const float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);  // d * (+-i)
        const float2 b0 = cadd(t0, t2), b1 = cadd(t1, t3), b2 = csub(t0, t2), b3 = csub(t1, t3);
        float2* yo = y + q + 4 * S * p;
        if (S == 64) {                                 // p == 0: all roots are 1
            yo[0] = b0; yo[S] = b1; yo[2 * S] = b2; yo[3 * S] = b3;
        } else {
            float2 w1 = tw[p * TWSTEP], w2 = tw[2 * p * TWSTEP], w3 = tw[3 * p * TWSTEP];
            if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            yo[0] = b0;
            yo[S] = cmul(b1, w1);
            yo[2 * S] = cmul(b2, w2);
            yo[3 * S] = cmul(b3, w3);
        }
    }
    __syncwarp();
}

template <bool INV>
__device__ __forceinline__ float2* warp_fft256(float2* x, float2* y, const float2* __restrict__ tw, int lane) {
    fft256_pass<INV, 1>(x, y, tw, lane);
    fft256_pass<INV, 4>(y, x, tw, lane);
    fft256_pass<INV, 16>(x, y, tw, lane);
    fft256_pass<INV, 64>(y, x, tw, lane);
    return x;
}

// ---------------------------------------------------------------------------------------------------------------------
// M = 256 (n_fft = 512) in REGISTERS: every lane holds 8 of the 256 complex points, the transform is radix 8 x 8 x 4 with the
// butterflies in registers and only TWO exchanges through a small shared-memory pad (the ping-pong version above does four full
// passes through shared memory: 8 loads + 6 twiddle loads + 8 stores per lane and pass, the stores of the first two passes 8- and
// 4-way bank conflicted; measured, that made the STFT / iSTFT kernels instruction bound at 6-9 % of the HBM roofline).
// Same Stockham index algebra as warp_fft (so the results agree to rounding):
//   pass 1 (r 8, m 32, s 1):  lane i:            in x[i + 32 j]            out y1[8 i + k]          * W256^(i k)
//   pass 2 (r 8, m 4,  s 8):  lane i = 8 p + q:  in y1[i + 32 j]           out y2[q + 64 p + 8 k]   * W32^(p k)
//   pass 3 (r 4, m 1, s 64):  lane i:            in y2[i + 64 j], y2[i + 32 + 64 j]   out Z[i + 64 k], Z[i + 32 + 64 k]
// Input and output are both "lane-strided" (element lane + 32 j in register j): coalesced global accesses on either side.
// The per-lane twiddles depend on the lane only and are loaded once per kernel.
// ---------------------------------------------------------------------------------------------------------------------
struct FftRegTw {
    float2 w1[7];      // W256^(lane * k),     k = 1..7
    float2 w2[7];      // W32^((lane / 8) * k)
};
constexpr int FFT_XCH = 320;                       // float2 slots of one warp's exchange pad

__device__ __forceinline__ void fft_reg_tw_init(FftRegTw& t, const float2* __restrict__ tw512, int lane) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        t.w1[k - 1] = tw512[2 * lane * k];          // <= 2*31*7 = 434 < 512
        t.w2[k - 1] = tw512[16 * (lane >> 3) * k];  // <= 16*3*7 = 336
    }
}

template <bool INV>
__device__ __forceinline__ float2 mul_i(float2 a) {   // a * (-i) forward, a * (+i) inverse
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ void dft4_reg(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 e0 = cadd(a0, a2), e1 = csub(a0, a2), o0 = cadd(a1, a3), o1 = mul_i<INV>(csub(a1, a3));
    a0 = cadd(e0, o0); a1 = cadd(e1, o1); a2 = csub(e0, o0); a3 = csub(e1, o1);
}

// v[k] <- sum_j v[j] W8^(jk)   (W8 = exp(-+ 2 pi i / 8)), natural order in and out
template <bool INV>
__device__ __forceinline__ void dft8_reg(float2 (&v)[8]) {
    const float h = 0.70710678118654752f;
    float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    // b_j *= W8^j
    b1 = INV ? make_float2(h * (b1.x - b1.y), h * (b1.x + b1.y)) : make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));
    b2 = mul_i<INV>(b2);
    b3 = INV ? make_float2(-h * (b3.x + b3.y), h * (b3.x - b3.y)) : make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));
    dft4_reg<INV>(a0, a1, a2, a3);                  // even outputs 0, 2, 4, 6
    dft4_reg<INV>(b0, b1, b2, b3);                  // odd outputs 1, 3, 5, 7
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

template <bool INV>
__device__ __forceinline__ void warp_fft256_reg(float2 (&z)[8], const FftRegTw& t, float2* __restrict__ xch, int lane) {
    // ---- pass 1
    dft8_reg<INV>(z);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        float2 w = t.w1[k - 1];
        if (INV) w.y = -w.y;
        z[k] = cmul(z[k], w);
    }
    // exchange 1: y1[8 i + k] sits at slot 10 i + k (80-byte lane pitch: the four 16-byte stores of a lane are conflict free)
    {
        float4* wa = reinterpret_cast<float4*>(xch + 10 * lane);
#pragma unroll
        for (int k = 0; k < 4; ++k) wa[k] = make_float4(z[2 * k].x, z[2 * k].y, z[2 * k + 1].x, z[2 * k + 1].y);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = xch[10 * ((lane >> 3) + 4 * j) + (lane & 7)];      // y1[lane + 32 j]
    __syncwarp();
    // ---- pass 2
    dft8_reg<INV>(z);
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        float2 w = t.w2[k - 1];
        if (INV) w.y = -w.y;
        z[k] = cmul(z[k], w);
    }
    // exchange 2: y2[e] sits at slot e + 8 (e >> 6)  (72-slot pitch per 64: the four p blocks of a store hit different banks)
    {
        const int base2 = (lane & 7) + 72 * (lane >> 3);
#pragma unroll
        for (int k = 0; k < 8; ++k) xch[base2 + 8 * k] = z[k];                               // y2[q + 64 p + 8 k]
    }
    __syncwarp();
    float2 a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = xch[lane + 72 * j];                                                          // y2[lane + 64 j]
        b[j] = xch[lane + 32 + 72 * j];                                                     // y2[lane + 32 + 64 j]
    }
    __syncwarp();
    // ---- pass 3
    dft4_reg<INV>(a[0], a[1], a[2], a[3]);
    dft4_reg<INV>(b[0], b[1], b[2], b[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) { z[2 * k] = a[k]; z[2 * k + 1] = b[k]; }                    // Z[lane + 32 j] in z[j]
}

// value of register j' of lane (32 - lane) & 31 where the source element is Z[(256 - (lane + 32 j)) & 255]
__device__ __forceinline__ float2 fft_reg_partner(const float2 (&z)[8], int j, int lane) {
    const int src = (32 - lane) & 31;
    float2 v;
    // lanes != 0: element 256 - lane - 32 j = (32 - lane) + 32 (7 - j): register 7 - j of the partner lane
    v.x = __shfl_sync(0xffffffffu, z[7 - j].x, src);
    v.y = __shfl_sync(0xffffffffu, z[7 - j].y, src);
    if (lane == 0) v = z[(8 - j) & 7];               // lane 0: element (256 - 32 j) & 255 = 32 ((8 - j) & 7), its own register
    return v;
}

__device__ __forceinline__ float load_padded(const float* __restrict__ xb, int i, int L, int pad_mode) {
    if (i < 0) {
        if (pad_mode == CRUSE_PAD_CONSTANT) return 0.f;
        i = -i;
    } else if (i >= L) {
        if (pad_mode == CRUSE_PAD_CONSTANT) return 0.f;
        i = 2 * (L - 1) - i;
    }
    return __ldg(xb + i);
}

__global__ void __launch_bounds__(256)
stft_fwd_kernel(const float* __restrict__ wav, const float* __restrict__ window, float* __restrict__ spec,
                float* __restrict__ mag, int B, int L, int N, int hop, int T, int pad_mode, int mag_bins,
                float mag_eps, FftPlan plan) {
    extern __shared__ float2 sm_[];
    const int M = N >> 1, NF = M + 1;
    float2* tw = sm_;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float2* x = sm_ + N + (size_t)warp * 2 * M;
    float2* y = x + M;
    build_twiddles(tw, N);
    __syncthreads();

    const long long nframes = (long long)B * T;
    for (long long fr = (long long)blockIdx.x * nwarps + warp; fr < nframes; fr += (long long)gridDim.x * nwarps) {
        const int b = (int)(fr / T), t = (int)(fr - (long long)b * T);
        const float* xb = wav + (long long)b * L;
        const int start = t * hop - M;
        const bool interior = (start >= 0) && (start + N <= L) && ((((long long)b * L + start) & 1) == 0);
        if (interior) {
            const float2* x2 = reinterpret_cast<const float2*>(xb + start);
            const float2* w2 = reinterpret_cast<const float2*>(window);
            for (int n = lane; n < M; n += 32) {
                float2 v = __ldg(x2 + n), w = __ldg(w2 + n);
                x[n] = make_float2(v.x * w.x, v.y * w.y);
            }
        } else {
            for (int n = lane; n < M; n += 32) {
                float v0 = load_padded(xb, start + 2 * n, L, pad_mode);
                float v1 = load_padded(xb, start + 2 * n + 1, L, pad_mode);
                x[n] = make_float2(v0 * __ldg(window + 2 * n), v1 * __ldg(window + 2 * n + 1));
            }
        }
        __syncwarp();
        const float2* Z = warp_fft<false>(x, y, tw, M, plan, lane);
        float2* so = reinterpret_cast<float2*>(spec) + fr * NF;
        float* mo = mag ? mag + fr * mag_bins : nullptr;
        for (int k = lane; k <= M; k += 32) {
            float2 zk = Z[k == M ? 0 : k];
            float2 zc = Z[k == 0 ? 0 : M - k];
            zc.y = -zc.y;
            float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            float2 xo = make_float2(d.y, -d.x);  // d / i
            float2 X = cadd(xe, cmul(tw[k], xo));
            so[k] = X;
            if (mo && k < mag_bins) mo[k] = sqrtf(X.x * X.x + X.y * X.y + mag_eps);
        }
        __syncwarp();
    }
}

// n_fft = 512 (the geometry of the benchmark): one warp per frame, the 256-point complex FFT in registers (warp_fft256_reg).
// Every global access is lane-strided (element lane + 32 j): full 256-byte warp transactions for the samples, the window, the
// spectrum and the magnitude.
__global__ void __launch_bounds__(256, 3)
stft512_fwd_kernel(const float* __restrict__ wav, const float* __restrict__ window, float* __restrict__ spec,
                   float* __restrict__ mag, int B, int L, int hop, int T, int pad_mode, int mag_bins, float mag_eps) {
    constexpr int N = 512, M = 256, NF = 257;
    extern __shared__ float2 sm_[];
    float2* tw = sm_;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float2* xch = sm_ + N + (size_t)warp * FFT_XCH;
    build_twiddles(tw, N);
    __syncthreads();
    FftRegTw rt;
    fft_reg_tw_init(rt, tw, lane);
    const float2* w2 = reinterpret_cast<const float2*>(window);

    const long long nframes = (long long)B * T;
    for (long long fr = (long long)blockIdx.x * nwarps + warp; fr < nframes; fr += (long long)gridDim.x * nwarps) {
        const int b = (int)(fr / T), t = (int)(fr - (long long)b * T);
        const float* xb = wav + (long long)b * L;
        const int start = t * hop - M;
        const bool interior = (start >= 0) && (start + N <= L) && ((((long long)b * L + start) & 1) == 0);
        float2 z[8];
        if (interior) {
            const float2* x2 = reinterpret_cast<const float2*>(xb + start);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 v = __ldg(x2 + lane + 32 * j), w = __ldg(w2 + lane + 32 * j);
                z[j] = make_float2(v.x * w.x, v.y * w.y);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = lane + 32 * j;
                const float2 w = __ldg(w2 + n);
                z[j] = make_float2(load_padded(xb, start + 2 * n, L, pad_mode) * w.x, load_padded(xb, start + 2 * n + 1, L, pad_mode) * w.y);
            }
        }
        warp_fft256_reg<false>(z, rt, xch, lane);
        float2* so = reinterpret_cast<float2*>(spec) + fr * NF;
        float* mo = mag ? mag + fr * mag_bins : nullptr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = lane + 32 * j;
            const float2 zk = z[j];
            float2 zc = fft_reg_partner(z, j, lane);
            zc.y = -zc.y;
            const float2 xe = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            const float2 xo = make_float2(d.y, -d.x);  // d / i
            const float2 X = cadd(xe, cmul(tw[k], xo));
            so[k] = X;
            if (mo && k < mag_bins) mo[k] = sqrtf(X.x * X.x + X.y * X.y + mag_eps);
        }
        if (lane == 0) {                                // Nyquist bin: X[256] = Re Z[0] - Im Z[0]
            const float2 X = make_float2(z[0].x - z[0].y, 0.f);
            so[M] = X;
            if (mo && M < mag_bins) mo[M] = sqrtf(X.x * X.x + mag_eps);
        }
    }
}

// grid (chunks, B).  FC frames per chunk + halo recomputed frames in front.
__global__ void __launch_bounds__(256)
mask_istft_kernel(const float* __restrict__ spec, const float* __restrict__ mask, const float* __restrict__ window,
                  float* __restrict__ est_spec, float* __restrict__ wav, int B, int L, int N, int hop, int T,
                  int mask_bins, int FC, int halo, FftPlan plan, int c_begin) {
    extern __shared__ float2 sm_[];
    const int M = N >> 1, NF = M + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float2* tw = sm_;
    float2* x = sm_ + N + (size_t)warp * 2 * M;
    float2* y = x + M;
    float* frames = reinterpret_cast<float*>(sm_ + N + (size_t)nwarps * 2 * M);
    const int b = blockIdx.y, t0 = (blockIdx.x + c_begin) * FC, tbeg = t0 - halo, nfr = FC + halo;
    build_twiddles(tw, N);
    __syncthreads();
    const float inv_m = 1.0f / (float)M;

    for (int fi = warp; fi < nfr; fi += nwarps) {
        const int t = tbeg + fi;
        if (t < 0 || t >= T) continue;  // never read by the gather below
        const long long fr = (long long)b * T + t;
        const float2* si = reinterpret_cast<const float2*>(spec) + fr * NF;
        const float* mi = mask ? mask + fr * mask_bins : nullptr;
        if (est_spec && t >= t0) {
            float2* eo = reinterpret_cast<float2*>(est_spec) + fr * NF;
            for (int k = lane; k <= M; k += 32) {
                float2 v = __ldg(si + k);
                float mk = (mi && k < mask_bins) ? __ldg(mi + k) : 1.f;
                eo[k] = make_float2(v.x * mk, v.y * mk);
            }
        }
        if (wav) {
            for (int k = lane; k < M; k += 32) {
                float2 xk = __ldg(si + k), xm = __ldg(si + (M - k));
                float mk = (mi && k < mask_bins) ? __ldg(mi + k) : 1.f;
                float mm = (mi && (M - k) < mask_bins) ? __ldg(mi + (M - k)) : 1.f;
                xk.x *= mk; xk.y *= mk; xm.x *= mm; xm.y *= mm;
                if (k == 0) { xk.y = 0.f; xm.y = 0.f; }  // c2r ignores imag of DC and Nyquist
                xm.y = -xm.y;                             // conj(X[M-k])
                float2 xe = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y + xm.y));
                float2 d = make_float2(0.5f * (xk.x - xm.x), 0.5f * (xk.y - xm.y));
                float2 w = tw[k];
                w.y = -w.y;  // W_N^{-k}
                float2 xo = cmul(d, w);
                x[k] = make_float2(xe.x - xo.y, xe.y + xo.x);  // Xe + i*Xo
            }
            __syncwarp();
            const float2* z = (M == 256) ? warp_fft256<true>(x, y, tw, lane) : warp_fft<true>(x, y, tw, M, plan, lane);
            float* fb = frames + (size_t)fi * N;
            for (int n = lane; n < M; n += 32) {
                float2 v = z[n];
                fb[2 * n] = v.x * inv_m * __ldg(window + 2 * n);
                fb[2 * n + 1] = v.y * inv_m * __ldg(window + 2 * n + 1);
            }
            __syncwarp();
        }
    }
    if (!wav) return;
    __syncthreads();

    const bool last = (t0 + FC >= T);
    const long long m_lo = (long long)t0 * hop;
    long long m_hi = (long long)(t0 + FC) * hop;
    if (last) {
        long long a = (long long)(T - 1) * hop + N, c = (long long)L + M;
        m_hi = a > c ? a : c;
    }
    float* wo = wav + (long long)b * L;
    // overlap-add by (frame slot i, offset r inside the hop): sample m = (t0+i)*hop + r is covered by frame t0+i at offset r,
    // frame t0+i-1 at offset r+hop, ... while the offset stays below N -- no division per sample
    const int nslots = (int)((m_hi - m_lo + hop - 1) / hop);
    for (int i = 0; i < nslots; ++i) {
        const int tt = t0 + i;
        for (int r = threadIdx.x; r < hop; r += blockDim.x) {
            const long long m = m_lo + (long long)i * hop + r;
            const long long sidx = m - M;
            if (m >= m_hi || sidx < 0 || sidx >= L) continue;
            float sum = 0.f, env = 0.f;
            int t = tt;
            for (int off = r; off < N; off += hop, --t) {
                if (t < 0) break;
                if (t > T - 1) continue;
                sum += frames[(size_t)(t - tbeg) * N + off];
                const float w = __ldg(window + off);
                env += w * w;
            }
            wo[sidx] = env > 1e-11f ? sum / env : 0.f;
        }
    }
}

// n_fft = 512: the same CTA structure (a run of FC frames of one utterance + halo frames recomputed, overlap-add as a shared-memory
// gather) with the inverse transform in registers (warp_fft256_reg) and lane-strided global accesses; the noisy spectrum is
// read ONCE per frame for both the enhanced spectrum and the waveform.
__global__ void __launch_bounds__(256, 3)
mask_istft512_kernel(const float* __restrict__ spec, const float* __restrict__ mask, const float* __restrict__ window,
                     float* __restrict__ est_spec, float* __restrict__ wav, int B, int L, int hop, int T,
                     int mask_bins, int FC, int halo, int c_begin) {
    constexpr int N = 512, M = 256, NF = 257;
    extern __shared__ float2 sm_[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float2* tw = sm_;
    float2* xch = sm_ + N + (size_t)warp * FFT_XCH;
    float* frames = reinterpret_cast<float*>(sm_ + N + (size_t)nwarps * FFT_XCH);
    const int b = blockIdx.y, t0 = (blockIdx.x + c_begin) * FC, tbeg = t0 - halo, nfr = FC + halo;
    build_twiddles(tw, N);
    __syncthreads();
    FftRegTw rt;
    fft_reg_tw_init(rt, tw, lane);
    const float inv_m = 1.0f / (float)M;
    const float2* w2 = reinterpret_cast<const float2*>(window);

    for (int fi = warp; fi < nfr; fi += nwarps) {
        const int t = tbeg + fi;
        if (t < 0 || t >= T) continue;  // never read by the gather below
        const long long fr = (long long)b * T + t;
        const float2* si = reinterpret_cast<const float2*>(spec) + fr * NF;
        const float* mi = mask ? mask + fr * mask_bins : nullptr;
        const bool want_est = est_spec && t >= t0;
        float2* eo = reinterpret_cast<float2*>(est_spec) + fr * NF;
        float2 z[8];
        float2 xk[8], xm[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {                 // all 16 spectrum loads (and the mask) of the frame in flight together
            const int k = lane + 32 * j;
            xk[j] = __ldg(si + k);
            xm[j] = __ldg(si + (M - k));
            const float mk = (mi && k < mask_bins) ? __ldg(mi + k) : 1.f;
            const float mm = (mi && (M - k) < mask_bins) ? __ldg(mi + (M - k)) : 1.f;
            xk[j].x *= mk; xk[j].y *= mk; xm[j].x *= mm; xm[j].y *= mm;
        }
        if (want_est) {
#pragma unroll
            for (int j = 0; j < 8; ++j) eo[lane + 32 * j] = xk[j];
            if (lane == 0) eo[M] = xm[0];             // bin 256 = X[M - 0] of lane 0
        }
        if (wav) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = lane + 32 * j;
                float2 a = xk[j], c = xm[j];
                if (k == 0) { a.y = 0.f; c.y = 0.f; }  // c2r ignores imag of DC and Nyquist
                c.y = -c.y;                            // conj(X[M-k])
                const float2 xe = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y + c.y));
                const float2 d = make_float2(0.5f * (a.x - c.x), 0.5f * (a.y - c.y));
                float2 w = tw[k];
                w.y = -w.y;  // W_N^{-k}
                const float2 xo = cmul(d, w);
                z[j] = make_float2(xe.x - xo.y, xe.y + xo.x);  // Xe + i*Xo
            }
            warp_fft256_reg<true>(z, rt, xch, lane);
            float2* fb = reinterpret_cast<float2*>(frames + (size_t)fi * N);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = lane + 32 * j;
                const float2 w = __ldg(w2 + n);
                fb[n] = make_float2(z[j].x * inv_m * w.x, z[j].y * inv_m * w.y);
            }
        }
    }
    if (!wav) return;
    __syncthreads();

    const bool last = (t0 + FC >= T);
    const long long m_lo = (long long)t0 * hop;
    long long m_hi = (long long)(t0 + FC) * hop;
    if (last) {
        long long a = (long long)(T - 1) * hop + N, c = (long long)L + M;
        m_hi = a > c ? a : c;
    }
    float* wo = wav + (long long)b * L;
    const int nslots = (int)((m_hi - m_lo + hop - 1) / hop);
    for (int i = 0; i < nslots; ++i) {
        const int tt = t0 + i;
        for (int r = threadIdx.x; r < hop; r += blockDim.x) {
            const long long m = m_lo + (long long)i * hop + r;
            const long long sidx = m - M;
            if (m >= m_hi || sidx < 0 || sidx >= L) continue;
            float sum = 0.f, env = 0.f;
            int t = tt;
            for (int off = r; off < N; off += hop, --t) {
                if (t < 0) break;
                if (t > T - 1) continue;
                sum += frames[(size_t)(t - tbeg) * N + off];
                const float w = __ldg(window + off);
                env += w * w;
            }
            wo[sidx] = env > 1e-11f ? sum / env : 0.f;
        }
    }
}

__global__ void mask_bwd_kernel(const float* __restrict__ dest, const float* __restrict__ spec,
                                const float* __restrict__ gscale, const float* __restrict__ mask, float* __restrict__ dmask,
                                long long rows, int NF, int mask_bins) {
    const float g = gscale ? __ldg(gscale) : 1.f;
    const long long total = rows * mask_bins;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / mask_bins;
        const int f = (int)(i - r * mask_bins);
        const float2 d = __ldg(reinterpret_cast<const float2*>(dest) + r * NF + f);
        const float2 x = __ldg(reinterpret_cast<const float2*>(spec) + r * NF + f);
        float v = g * (d.x * x.x + d.y * x.y);
        if (mask) { const float m = __ldg(mask + i); v *= m * (1.f - m); }   // through the sigmoid of cruse_net.py:164
        dmask[i] = v;
    }
}

}  // namespace cruse

using namespace cruse;

extern "C" int cruse_stft_fwd(const float* wav, const float* window, float* spec, float* mag, int B, int L, int n_fft,
                              int hop, int T, int pad_mode, int mag_bins, float mag_eps, void* stream) {
    CRUSE_CHECK_ARG(wav && window && spec, "stft_fwd: null pointer");
    CRUSE_CHECK_ARG(B > 0 && L > 0 && hop > 0 && n_fft >= 4 && (n_fft % 2) == 0, "stft_fwd: bad sizes B=%d L=%d n_fft=%d hop=%d", B, L, n_fft, hop);
    CRUSE_CHECK_ARG(T == 1 + L / hop, "stft_fwd: T=%d must equal 1 + L/hop = %d (center=True)", T, 1 + L / hop);
    CRUSE_CHECK_ARG(pad_mode == CRUSE_PAD_CONSTANT || L > n_fft / 2, "stft_fwd: reflect padding needs L > n_fft/2");
    CRUSE_CHECK_ARG(mag_bins >= 0 && mag_bins <= n_fft / 2 + 1, "stft_fwd: mag_bins out of range");
    FftPlan plan;
    CRUSE_CHECK_ARG(make_plan(n_fft / 2, plan) == 0, "stft_fwd: n_fft/2=%d must factor into 2,3,5", n_fft / 2);
    const int threads = 256, nwarps = threads / 32;
    const long long nframes = (long long)B * T;
    if (n_fft == 512) {
        const size_t smem512 = sizeof(float2) * ((size_t)n_fft + (size_t)nwarps * FFT_XCH);
        long long blocks = (nframes + nwarps - 1) / nwarps;
        const long long cap = (long long)sm_count() * 3;         // persistent: 3 CTAs per SM, each builds its twiddle table once
        if (blocks > cap) blocks = cap;
        stft512_fwd_kernel<<<(unsigned)blocks, threads, smem512, (cudaStream_t)stream>>>(wav, window, spec, mag_bins > 0 ? mag : nullptr, B, L,
                                                                                        hop, T, pad_mode, mag_bins, mag_eps);
        CRUSE_LAUNCH_OK();
        return 0;
    }
    const size_t smem = sizeof(float2) * ((size_t)n_fft + (size_t)nwarps * n_fft);
    CRUSE_CHECK_ARG(smem <= 200 * 1024, "stft_fwd: n_fft=%d too large for on-chip FFT", n_fft);
    CRUSE_CUDA_OK(cudaFuncSetAttribute(stft_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (nframes + nwarps - 1) / nwarps;
    const long long cap = (long long)sm_count() * 6;  // persistent-ish (6 CTAs of 36 KB fit an SM): each CTA builds its twiddle table once
    if (blocks > cap) blocks = cap;
    stft_fwd_kernel<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(wav, window, spec, mag_bins > 0 ? mag : nullptr, B, L, n_fft,
                                                                             hop, T, pad_mode, mag_bins, mag_eps, plan);
    CRUSE_LAUNCH_OK();
    return 0;
}

static int istft_chunk_frames(int n_fft, int hop) {
    const int nwarps = 256 / 32, halo = (n_fft - 1) / hop;
    // n_fft = 512: one frame per warp (FC + halo = 8): twice the CTAs of the two-round version, half the serial work in each --
    // the frame ranges that follow the decoder are short and latency bound
    const int FC = (n_fft == 512 ? 1 : 2) * nwarps - halo;
    return FC < 1 ? nwarps : FC;
}

// frames per CTA of cruse_mask_istft_fwd: CTA c produces est_spec of the frames [c*FC, (c+1)*FC) and the samples they start,
// reading the mask of the frames [c*FC - (n_fft-1)/hop, (c+1)*FC)
extern "C" int cruse_mask_istft_chunk_frames(int n_fft, int hop) {
    if (n_fft < 4 || hop <= 0 || hop > n_fft) return -1;
    return istft_chunk_frames(n_fft, hop);
}

static int mask_istft_launch(const float* spec, const float* mask, const float* window, float* est_spec, float* wav, int B, int L,
                             int n_fft, int hop, int T, int mask_bins, int c_begin, int c_end, void* stream);

// the CTAs [c_begin, c_end) of cruse_mask_istft_fwd only (c_end <= ceil(T / chunk_frames)): lets mask*X + iSTFT follow the
// decoder range by range; the union of disjoint CTA ranges covering [0, ceil(T/FC)) is bit-identical to the whole call
extern "C" int cruse_mask_istft_fwd_range(const float* spec, const float* mask, const float* window, float* est_spec,
                                          float* wav, int B, int L, int n_fft, int hop, int T, int mask_bins, int c_begin,
                                          int c_end, void* stream) {
    return mask_istft_launch(spec, mask, window, est_spec, wav, B, L, n_fft, hop, T, mask_bins, c_begin, c_end, stream);
}

extern "C" int cruse_mask_istft_fwd(const float* spec, const float* mask, const float* window, float* est_spec,
                                    float* wav, int B, int L, int n_fft, int hop, int T, int mask_bins, void* stream) {
    return mask_istft_launch(spec, mask, window, est_spec, wav, B, L, n_fft, hop, T, mask_bins, 0, -1, stream);
}

static int mask_istft_launch(const float* spec, const float* mask, const float* window, float* est_spec, float* wav, int B, int L,
                             int n_fft, int hop, int T, int mask_bins, int c_begin, int c_end, void* stream) {
    CRUSE_CHECK_ARG(spec && window, "mask_istft_fwd: null pointer");
    CRUSE_CHECK_ARG(est_spec || wav, "mask_istft_fwd: nothing to compute");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && hop > 0 && n_fft >= 4 && (n_fft % 2) == 0 && hop <= n_fft, "mask_istft_fwd: bad sizes");
    CRUSE_CHECK_ARG(mask_bins >= 0 && mask_bins <= n_fft / 2 + 1, "mask_istft_fwd: mask_bins out of range");
    CRUSE_CHECK_ARG(!wav || L > 0, "mask_istft_fwd: L must be positive");
    FftPlan plan;
    CRUSE_CHECK_ARG(make_plan(n_fft / 2, plan) == 0, "mask_istft_fwd: n_fft/2=%d must factor into 2,3,5", n_fft / 2);
    const int threads = 256, nwarps = threads / 32;
    const int halo = (n_fft - 1) / hop;
    const int FC = istft_chunk_frames(n_fft, hop);
    const size_t smem = sizeof(float2) * ((size_t)n_fft + (size_t)nwarps * n_fft) + sizeof(float) * (size_t)(FC + halo) * n_fft;
    CRUSE_CHECK_ARG(smem <= 220 * 1024, "mask_istft_fwd: n_fft=%d / hop=%d need too much shared memory", n_fft, hop);
    const int nchunks = (T + FC - 1) / FC;
    if (c_end < 0) c_end = nchunks;
    CRUSE_CHECK_ARG(c_begin >= 0 && c_begin < c_end && c_end <= nchunks, "mask_istft_fwd: CTA range [%d,%d) outside [0,%d)", c_begin, c_end, nchunks);
    dim3 grid(c_end - c_begin, B);
    if (n_fft == 512) {
        const size_t smem512 = sizeof(float2) * ((size_t)n_fft + (size_t)nwarps * FFT_XCH) + sizeof(float) * (size_t)(FC + halo) * n_fft;
        static bool attr512 = false;
        if (!attr512) {
            CRUSE_CUDA_OK(cudaFuncSetAttribute(mask_istft512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem512));
            attr512 = true;
        }
        mask_istft512_kernel<<<grid, threads, smem512, (cudaStream_t)stream>>>(spec, mask_bins > 0 ? mask : nullptr, window, est_spec, wav, B, L,
                                                                              hop, T, mask_bins, FC, halo, c_begin);
        CRUSE_LAUNCH_OK();
        return 0;
    }
    CRUSE_CUDA_OK(cudaFuncSetAttribute(mask_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mask_istft_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(spec, mask_bins > 0 ? mask : nullptr, window, est_spec, wav, B, L,
                                                                    n_fft, hop, T, mask_bins, FC, halo, plan, c_begin);
    CRUSE_LAUNCH_OK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// backward of the iSTFT (a9 for time-domain losses).  wav = istft(spec) is linear: irfft per frame, window, overlap-add,
// divide by the window-power envelope, trim n_fft/2.  Its adjoint is therefore an STFT: with g[s] = dwav[s] / env[s + n_fft/2]
// (0 where the forward wrote 0), dspec[t,k] = c_k / n_fft * rfft(window * frame_t(g, zero padded))[k], c_0 = c_{n_fft/2} = 1,
// c_k = 2 otherwise (a real inverse FFT sees every interior bin twice).  The transform itself is cruse_stft_fwd with constant
// padding; the two small kernels here do the envelope division and the c_k / n_fft scale.
// ---------------------------------------------------------------------------------------------
namespace cruse {
__global__ void __launch_bounds__(256)
istft_bwd_prep_kernel(const float* __restrict__ dwav, const float* __restrict__ window, float* __restrict__ g, long long total, int L,
                      int N, int hop, int T) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(i % L);
        const int m = s + (N >> 1);
        int t = m / hop;
        if (t > T - 1) t = T - 1;
        float env = 0.f;
        for (; t >= 0; --t) {
            const int off = m - t * hop;
            if (off >= N) break;
            const float w = __ldg(window + off);
            env += w * w;
        }
        g[i] = env > 1e-11f ? __ldg(dwav + i) / env : 0.f;
    }
}

__global__ void __launch_bounds__(256)
rfft_adjoint_scale_kernel(float2* __restrict__ dspec, long long total, int NF, float inv_n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % NF);
        const float sc = (k == 0 || k == NF - 1) ? inv_n : 2.f * inv_n;
        float2 v = dspec[i];
        dspec[i] = make_float2(v.x * sc, v.y * sc);
    }
}
}  // namespace cruse

extern "C" size_t cruse_istft_bwd_ws_bytes(int B, int L) { return sizeof(float) * (size_t)(B > 0 ? B : 0) * (size_t)(L > 0 ? L : 0); }

extern "C" int cruse_istft_bwd(const float* dwav, const float* window, float* dspec, void* ws, int B, int L, int n_fft, int hop,
                               int T, void* stream) {
    CRUSE_CHECK_ARG(dwav && window && dspec && ws, "istft_bwd: null pointer");
    CRUSE_CHECK_ARG(B > 0 && L > 0 && hop > 0 && n_fft >= 4 && (n_fft % 2) == 0 && hop <= n_fft, "istft_bwd: bad sizes B=%d L=%d n_fft=%d hop=%d", B, L, n_fft, hop);
    CRUSE_CHECK_ARG(T == 1 + L / hop, "istft_bwd: T=%d must equal 1 + L/hop = %d", T, 1 + L / hop);
    cudaStream_t st = (cudaStream_t)stream;
    float* g = static_cast<float*>(ws);
    const long long total = (long long)B * L;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)cruse::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cruse::istft_bwd_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(dwav, window, g, total, L, n_fft, hop, T);
    CRUSE_LAUNCH_OK();
    if (int rc = cruse_stft_fwd(g, window, dspec, nullptr, B, L, n_fft, hop, T, CRUSE_PAD_CONSTANT, 0, 0.f, stream)) return rc;
    const int NF = n_fft / 2 + 1;
    const long long n = (long long)B * T * NF;
    blocks = (n + 255) / 256;
    if (blocks > cap) blocks = cap;
    cruse::rfft_adjoint_scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<float2*>(dspec), n, NF, 1.0f / (float)n_fft);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_mask_bwd(const float* dest, const float* spec, const float* gscale, const float* mask, float* dmask,
                              int B, int T, int NF, int mask_bins, void* stream) {
    CRUSE_CHECK_ARG(dest && spec && dmask, "mask_bwd: null pointer");
    CRUSE_CHECK_ARG(mask_bins > 0 && mask_bins <= NF, "mask_bwd: mask_bins out of range");
    const long long rows = (long long)B * T, total = rows * mask_bins;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    mask_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dest, spec, gscale, mask, dmask, rows, NF, mask_bins);
    CRUSE_LAUNCH_OK();
    return 0;
}

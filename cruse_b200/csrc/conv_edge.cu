// conv_edge.cu -- the two stages of the U-Net that have a single channel on one side: the first encoder stage
// (Conv2d(1 -> C, (2,3), stride (1,2), pad (1,1)) + "[..., :-1, :]" + folded BN + act, model/cruse_net.py:138,141,149-152)
// and the last decoder stage (ConvTranspose2d(C -> 1, (1,3), stride (1,2)) + "[..., :-1]" + sigmoid, :164).
//
// With one channel on one side there is no GEMM to speak of (K = 6 resp. N = 2): these are pure HBM streams --
// 1 KB in / 4 KB out per frame (stage 1) and 4 KB in / 1 KB out (last stage) -- so they are written as streaming
// kernels: one thread per (frame, bin), every global access a fully coalesced 128/256-byte warp transaction, the
// neighbouring bin through a warp shuffle, all weights and folded BN parameters in registers.  Exact fp32.
// Used for eval-mode whole-utterance calls (no history frame, no batch statistics); everything else stays on conv.cu.
#include "common.cuh"

namespace cruse {
namespace {

// ---------------------------------------------------------------------------------------------
// stage 1: in [frames, 1, Fin] (Fin = 2*Fout), out [frames, COUT, Fout];  out[t] uses in[t-1] (kt = 0) and in[t] (kt = 1)
// block = Fout threads = one frame; frame t-1 of the same utterance is re-read through L1/L2
// ---------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128)
enc1_stream_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ alpha, int act,
                   float* __restrict__ out, int T, int Fout, long long nframes, const float* __restrict__ hist, int t_begin, int Tc) {
    __shared__ float s_w[COUT * 6 + 3 * COUT];
    for (int i = threadIdx.x; i < COUT * 6; i += blockDim.x) s_w[i] = __ldg(w + i);            // [COUT][1][2][3]
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
        const float sc = scale ? __ldg(scale + i) : 1.f;
        s_w[COUT * 6 + i] = sc;
        s_w[COUT * 7 + i] = fmaf(bias ? __ldg(bias + i) : 0.f, sc, shift ? __ldg(shift + i) : 0.f);
        s_w[COUT * 8 + i] = alpha ? __ldg(alpha + i) : 0.f;
    }
    __syncthreads();
    const int Fin = 2 * Fout;
    const int lane = threadIdx.x & 31;
    for (long long i = blockIdx.x; i < nframes; i += gridDim.x) {       // nframes = B * Tc frames of the range [t_begin, t_begin + Tc)
        const int t = t_begin + (int)(i % Tc);
        const long long fr = (i / Tc) * T + t;
        for (int fo = threadIdx.x; fo < Fout; fo += blockDim.x) {        // Fout is a multiple of 32: warps stay converged
            const float* cur = in + fr * Fin + 2 * fo;
            const float2 c = __ldg(reinterpret_cast<const float2*>(cur));
            // frame t-1: the previous frame of the utterance, or (t == 0) the frame carried from the previous chunk of a stream
            const float* prv = t > 0 ? cur - Fin : (hist ? hist + (fr / T) * Fin + 2 * fo : nullptr);
            float2 p = make_float2(0.f, 0.f);
            if (prv) p = __ldg(reinterpret_cast<const float2*>(prv));
            // bin 2fo-1 = the odd element of the previous thread (previous warp's last lane: one extra 4-byte load)
            float cl = __shfl_up_sync(0xffffffffu, c.y, 1), pl = __shfl_up_sync(0xffffffffu, p.y, 1);
            if (lane == 0) {
                cl = fo > 0 ? __ldg(cur - 1) : 0.f;
                pl = (fo > 0 && prv) ? __ldg(prv - 1) : 0.f;
            }
            float* o = out + fr * (long long)(COUT * Fout) + fo;
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                const float* wc = s_w + co * 6;
                float v = wc[0] * pl;
                v = fmaf(wc[1], p.x, v); v = fmaf(wc[2], p.y, v);
                v = fmaf(wc[3], cl, v); v = fmaf(wc[4], c.x, v); v = fmaf(wc[5], c.y, v);
                o[(size_t)co * Fout] = apply_act(fmaf(v, s_w[COUT * 6 + co], s_w[COUT * 7 + co]), act, s_w[COUT * 8 + co]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// last stage: in [frames, CIN, Fin], out [frames, 1, 2*Fin]:
//   out[2i] = b + sum_ci W[ci,0,0] x[ci,i] + W[ci,0,2] x[ci,i-1];   out[2i+1] = b + sum_ci W[ci,0,1] x[ci,i]
// ---------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(128)
dec1_stream_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ alpha, int act,
                   float* __restrict__ out, int Fin, long long nframes, int T, int t_begin, int Tc) {
    __shared__ float s_w[CIN * 3];
    for (int i = threadIdx.x; i < CIN * 3; i += blockDim.x) s_w[i] = __ldg(w + i);             // [CIN][1][1][3]
    __syncthreads();
    const float sc = scale ? __ldg(scale) : 1.f;
    const float sh = fmaf(bias ? __ldg(bias) : 0.f, sc, shift ? __ldg(shift) : 0.f);
    const float al = alpha ? __ldg(alpha) : 0.f;
    const int lane = threadIdx.x & 31;
    for (long long j = blockIdx.x; j < nframes; j += gridDim.x) {       // nframes = B * Tc frames of the range [t_begin, t_begin + Tc)
        const long long fr = (j / Tc) * T + t_begin + (j % Tc);
        for (int i = threadIdx.x; i < Fin; i += blockDim.x) {
            const float* x = in + fr * (long long)(CIN * Fin) + i;
            float xv[CIN];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) xv[ci] = __ldg(x + (size_t)ci * Fin);
            float e = 0.f, o = 0.f;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                float xm = __shfl_up_sync(0xffffffffu, xv[ci], 1);
                if (lane == 0) xm = i > 0 ? __ldg(x + (size_t)ci * Fin - 1) : 0.f;
                e = fmaf(s_w[ci * 3 + 0], xv[ci], e);
                e = fmaf(s_w[ci * 3 + 2], xm, e);
                o = fmaf(s_w[ci * 3 + 1], xv[ci], o);
            }
            *reinterpret_cast<float2*>(out + fr * (long long)(2 * Fin) + 2 * i) =
                make_float2(apply_act(fmaf(e, sc, sh), act, al), apply_act(fmaf(o, sc, sh), act, al));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward of the same two stages (training step).  The weight gradients are reductions over every (frame, bin) of
// 6 resp. 3 taps x 8 channels: each thread keeps all of them in registers while the CTA strides over frames, one block
// reduction at the end, one partial row [weights | bias] per CTA (summed in fixed order by colsum_kernel: deterministic).
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_reduce_store(float (&acc)[N], float* s_red, float* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp * N + j] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < nwarps; ++w) v += s_red[w * N + j];
        dst[j] = v;
    }
}

// stage 1: dW[co][0][kt][kf] = sum dz[co,t,fo] * x[t-1+kt, 2fo-1+kf], dbias[co] = sum dz[co,t,fo]
template <int COUT>
__global__ void __launch_bounds__(128)
enc1_wgrad_stream_kernel(const float* __restrict__ in, const float* __restrict__ dz, float* __restrict__ ws, int T, int Fout,
                         long long nframes, int pitch) {
    __shared__ float s_red[4 * COUT * 7];
    float acc[COUT * 7];                                   // [co][6 taps] then [co] bias
#pragma unroll
    for (int j = 0; j < COUT * 7; ++j) acc[j] = 0.f;
    const int Fin = 2 * Fout;
    const int lane = threadIdx.x & 31;
    for (long long fr = blockIdx.x; fr < nframes; fr += gridDim.x) {
        const int t = (int)(fr % T);
        for (int fo = threadIdx.x; fo < Fout; fo += blockDim.x) {
            const float* cur = in + fr * Fin + 2 * fo;
            const float2 c = __ldg(reinterpret_cast<const float2*>(cur));
            float2 p = make_float2(0.f, 0.f);
            if (t > 0) p = __ldg(reinterpret_cast<const float2*>(cur - Fin));
            float cl = __shfl_up_sync(0xffffffffu, c.y, 1), pl = __shfl_up_sync(0xffffffffu, p.y, 1);
            if (lane == 0) {
                cl = fo > 0 ? __ldg(cur - 1) : 0.f;
                pl = (fo > 0 && t > 0) ? __ldg(cur - Fin - 1) : 0.f;
            }
            const float* g = dz + fr * (long long)(COUT * Fout) + fo;
            float gv[COUT];
#pragma unroll
            for (int co = 0; co < COUT; ++co) gv[co] = __ldg(g + (size_t)co * Fout);
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                acc[co * 6 + 0] = fmaf(gv[co], pl, acc[co * 6 + 0]);
                acc[co * 6 + 1] = fmaf(gv[co], p.x, acc[co * 6 + 1]);
                acc[co * 6 + 2] = fmaf(gv[co], p.y, acc[co * 6 + 2]);
                acc[co * 6 + 3] = fmaf(gv[co], cl, acc[co * 6 + 3]);
                acc[co * 6 + 4] = fmaf(gv[co], c.x, acc[co * 6 + 4]);
                acc[co * 6 + 5] = fmaf(gv[co], c.y, acc[co * 6 + 5]);
                acc[COUT * 6 + co] += gv[co];
            }
        }
    }
    block_reduce_store<COUT * 7>(acc, s_red, ws + (size_t)blockIdx.x * pitch);
}

// last stage: dW[ci][0][k] = sum_i x[ci,i] * dz[2i+k] (2i+k < 2*Fin), dbias = sum dz
template <int CIN>
__global__ void __launch_bounds__(128)
dec1_wgrad_stream_kernel(const float* __restrict__ in, const float* __restrict__ dz, float* __restrict__ ws, int Fin,
                         long long nframes, int pitch) {
    __shared__ float s_red[4 * (CIN * 3 + 1)];
    float acc[CIN * 3 + 1];
#pragma unroll
    for (int j = 0; j < CIN * 3 + 1; ++j) acc[j] = 0.f;
    const int lane = threadIdx.x & 31;
    for (long long fr = blockIdx.x; fr < nframes; fr += gridDim.x) {
        for (int i = threadIdx.x; i < Fin; i += blockDim.x) {
            const float* gp = dz + fr * (long long)(2 * Fin) + 2 * i;
            const float2 g = __ldg(reinterpret_cast<const float2*>(gp));
            float gn = __shfl_down_sync(0xffffffffu, g.x, 1);
            if (lane == 31) gn = i + 1 < Fin ? __ldg(gp + 2) : 0.f;
            const float* x = in + fr * (long long)(CIN * Fin) + i;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float xv = __ldg(x + (size_t)ci * Fin);
                acc[ci * 3 + 0] = fmaf(xv, g.x, acc[ci * 3 + 0]);
                acc[ci * 3 + 1] = fmaf(xv, g.y, acc[ci * 3 + 1]);
                acc[ci * 3 + 2] = fmaf(xv, gn, acc[ci * 3 + 2]);
            }
            acc[CIN * 3] += g.x + g.y;
        }
    }
    block_reduce_store<CIN * 3 + 1>(acc, s_red, ws + (size_t)blockIdx.x * pitch);
}

// last stage: din[ci,i] = sum_k W[ci,0,k] dz[2i+k]
template <int CIN>
__global__ void __launch_bounds__(128)
dec1_dgrad_stream_kernel(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ din, int Fin, long long nframes) {
    __shared__ float s_w[CIN * 3];
    for (int i = threadIdx.x; i < CIN * 3; i += blockDim.x) s_w[i] = __ldg(w + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (long long fr = blockIdx.x; fr < nframes; fr += gridDim.x) {
        for (int i = threadIdx.x; i < Fin; i += blockDim.x) {
            const float* gp = dz + fr * (long long)(2 * Fin) + 2 * i;
            const float2 g = __ldg(reinterpret_cast<const float2*>(gp));
            float gn = __shfl_down_sync(0xffffffffu, g.x, 1);
            if (lane == 31) gn = i + 1 < Fin ? __ldg(gp + 2) : 0.f;
            float* o = din + fr * (long long)(CIN * Fin) + i;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci)
                o[(size_t)ci * Fin] = fmaf(s_w[ci * 3 + 2], gn, fmaf(s_w[ci * 3 + 1], g.y, s_w[ci * 3] * g.x));
        }
    }
}

}  // namespace

// Weight / bias gradient of stage 1: returns the number of partial rows written to ws (pitch floats each), 0 = not this shape.
int conv_edge_wgrad_try(const float* in, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt,
                        int fstride, int pitch, int max_grid, cudaStream_t st) {
    if (!(Cin == 1 && Cout == 8 && kt == 2 && fstride == 2 && Fin == 2 * Fout && (Fout % 32) == 0 && Fout <= 128)) return 0;
    if ((reinterpret_cast<uintptr_t>(in) & 7) != 0 || max_grid < 1) return 0;
    const long long nframes = (long long)B * T;
    long long grid = (long long)sm_count() * 4;
    if (grid > nframes) grid = nframes;
    if (grid > max_grid) grid = max_grid;
    enc1_wgrad_stream_kernel<8><<<(int)grid, Fout < 128 ? Fout : 128, 0, st>>>(in, dz, ws, T, Fout, nframes, pitch);
    return cudaGetLastError() == cudaSuccess ? (int)grid : -3;
}

// Weight / bias gradient of the last decoder stage (same contract).
int convT_edge_wgrad_try(const float* in, const float* dz, float* ws, int B, int T, int Cin, int Fin, int Cout, int Fout, int pitch,
                         int max_grid, cudaStream_t st) {
    if (!(Cout == 1 && Cin == 8 && Fout == 2 * Fin && (Fin % 32) == 0)) return 0;
    if ((reinterpret_cast<uintptr_t>(dz) & 7) != 0 || max_grid < 1) return 0;
    const long long nframes = (long long)B * T;
    long long grid = (long long)sm_count() * 4;
    if (grid > nframes) grid = nframes;
    if (grid > max_grid) grid = max_grid;
    dec1_wgrad_stream_kernel<8><<<(int)grid, Fin < 128 ? Fin : 128, 0, st>>>(in, dz, ws, Fin, nframes, pitch);
    return cudaGetLastError() == cudaSuccess ? (int)grid : -3;
}

// Data gradient of the last decoder stage: 1 when launched here.
int convT_edge_dgrad_try(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin, int Fin, int Cout,
                         int Fout, cudaStream_t st) {
    if (!(Cout == 1 && Cin == 8 && Fout == 2 * Fin && (Fin % 32) == 0 && addend == nullptr)) return 0;
    if ((reinterpret_cast<uintptr_t>(dz) & 7) != 0) return 0;
    const long long nframes = (long long)B * T;
    const long long cap = (long long)sm_count() * 16;
    const int grid = (int)(nframes < cap ? nframes : cap);
    dec1_dgrad_stream_kernel<8><<<grid, Fin < 128 ? Fin : 128, 0, st>>>(dz, w, din, Fin, nframes);
    return cudaGetLastError() == cudaSuccess ? 1 : -3;
}

// Returns 1 when the stage was launched here, 0 when the shape is not one of these (caller runs the general kernel).
int conv_edge_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                  int act, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride, cudaStream_t st,
                  const float* hist, int t_begin, int t_end) {
    if ((reinterpret_cast<uintptr_t>(hist) & 7) != 0) return 0;
    if (t_end <= 0) { t_begin = 0; t_end = T; }
    if (t_begin < 0 || t_begin >= t_end || t_end > T) return -1;
    if (!(Cin == 1 && Cout == 8 && kt == 2 && fstride == 2 && Fin == 2 * Fout && (Fout % 32) == 0 && Fout <= 128)) return 0;
    if ((reinterpret_cast<uintptr_t>(in) & 7) != 0) return 0;
    const long long nframes = (long long)B * (t_end - t_begin);
    const long long cap = (long long)sm_count() * 16;
    const int grid = (int)(nframes < cap ? nframes : cap);
    enc1_stream_kernel<8><<<grid, Fout < 128 ? Fout : 128, 0, st>>>(in, w, bias, scale, shift, alpha, act, out, T, Fout, nframes, hist, t_begin,
                                                                  t_end - t_begin);
    return cudaGetLastError() == cudaSuccess ? 1 : -3;
}

int convT_edge_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                   int act, const float* skip, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, cudaStream_t st,
                   int t_begin, int t_end) {
    if (!(Cout == 1 && Cin == 8 && Fout == 2 * Fin && (Fin % 32) == 0 && skip == nullptr)) return 0;
    if ((reinterpret_cast<uintptr_t>(out) & 7) != 0) return 0;
    if (t_end <= 0) { t_begin = 0; t_end = T; }
    if (t_begin < 0 || t_begin >= t_end || t_end > T) return -1;
    const long long nframes = (long long)B * (t_end - t_begin);
    const long long cap = (long long)sm_count() * 16;
    const int grid = (int)(nframes < cap ? nframes : cap);
    dec1_stream_kernel<8><<<grid, Fin < 128 ? Fin : 128, 0, st>>>(in, w, bias, scale, shift, alpha, act, out, Fin, nframes, T, t_begin,
                                                                 t_end - t_begin);
    return cudaGetLastError() == cudaSuccess ? 1 : -3;
}

}  // namespace cruse

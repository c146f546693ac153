// loss.cu -- weighted-magnitude loss wo_male (loss_func/loss.py:121-148 of the reference),
// forward and d/d(est) in one streaming pass.
//
// Replaces ~20 ATen elementwise launches + one reduction on the reference path.  Pure HBM
// stream: 3 complex reads (+1 complex write for the gradient) per (b,t,f) bin; magnitudes, the
// IAM weight exp(alpha/(beta+|S|/|X|)) and the |log10| error stay in registers.  Reduction is
// warp shuffle -> shared memory -> one partial per CTA; a second 1-CTA launch sums the
// partials in double in a fixed order (deterministic, no atomics, graph-capturable).
#include "common.cuh"

namespace cruse {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_PARTS = 148 * 8;

__device__ __forceinline__ float2 ld_cplx(const float* __restrict__ p, long long off, long long im_off) {
    if (im_off == 1 && ((off & 1) == 0)) return __ldg(reinterpret_cast<const float2*>(p + off));
    return make_float2(__ldg(p + off), __ldg(p + off + im_off));
}

__global__ void __launch_bounds__(LOSS_THREADS)
wo_male_partial_kernel(const float* __restrict__ ref, cruse_cplx_layout lr, const float* __restrict__ est,
                       cruse_cplx_layout le, const float* __restrict__ unp, cruse_cplx_layout lu,
                       float* __restrict__ dest, float* __restrict__ partials, int T, int F, long long total,
                       float inv_count, const float* __restrict__ mask, int t_begin, int Tc) {
    // Tc < T: only the frames [t_begin, t_begin + Tc) of every utterance (total = B * Tc * F); the partial sums of several
    // ranges are added up by sum_partials_kernel
    // mask != NULL: the estimate is mask[b,t,f] * unproc[b,t,f] computed on the fly (PreProcess.masking, utils/utils.py:417-433,
    // fused into the loss so that it does not have to wait for -- and runs beside -- the mask*spectrum + iSTFT kernel)
    const float alpha = 2.f, beta = 1.f;              // loss.py:126-128 (gamma = 1)
    const float inv_ln10 = 0.43429448190325176f;
    float acc = 0.f;
    // one (b, t) row of F bins per CTA iteration; (b, t) advance incrementally -- no 64-bit division per element
    const long long rows = total / F;
    int t = t_begin + (int)(blockIdx.x % Tc);
    long long b = blockIdx.x / Tc;
    for (long long bt = blockIdx.x; bt < rows; bt += gridDim.x) {
      const long long rbase = b * lr.sb + t * lr.st, ebase = b * le.sb + t * le.st, ubase = b * lu.sb + t * lu.st;
      for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const long long i = (b * T + t) * F + f;
        const float2 r = ld_cplx(ref, rbase + f * lr.sf, lr.im_off);
        const long long eoff = ebase + f * le.sf;
        const float2 u = ld_cplx(unp, ubase + f * lu.sf, lu.im_off);
        float2 e;
        if (mask) {
            const float mk = __ldg(mask + i);
            e = make_float2(u.x * mk, u.y * mk);
        } else {
            e = ld_cplx(est, eoff, le.im_off);
        }
        const float mr = sqrtf(r.x * r.x + r.y * r.y);
        const float me = sqrtf(e.x * e.x + e.y * e.y);
        const float mu = sqrtf(u.x * u.x + u.y * u.y);
        const float iam = mr / mu;                       // :142, no eps (inf -> w = 1, 0/0 -> NaN as in torch)
        const float w = expf(alpha / (beta + iam));      // :143
        const float d = log10f(me + 1.f) - log10f(mr + 1.f);
        acc += w * fabsf(d);
        if (dest) {
            // d|d|/d me = sign(d) / ((me+1) ln10);  d me / d(re,im) = (re,im)/me  (0 at me == 0)
            const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            const float gme = w * sgn * inv_ln10 / (me + 1.f) * inv_count;
            const float s = me > 0.f ? gme / me : 0.f;
            if (le.im_off == 1 && ((eoff & 1) == 0)) {
                *reinterpret_cast<float2*>(dest + eoff) = make_float2(s * e.x, s * e.y);
            } else {
                dest[eoff] = s * e.x;
                dest[eoff + le.im_off] = s * e.y;
            }
        }
      }
      t += (int)(gridDim.x % Tc);
      b += gridDim.x / Tc;
      if (t >= t_begin + Tc) { t -= Tc; ++b; }
    }
    __shared__ float sh[LOSS_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < LOSS_THREADS / 32 ? sh[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) partials[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partials, int n, double scale,
                                                         float* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
    __shared__ double sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += sh[i];
        out[0] = (float)(a * scale);
    }
}

}  // namespace cruse

using namespace cruse;

extern "C" size_t cruse_wo_male_ws_bytes(void) { return sizeof(float) * LOSS_MAX_PARTS; }

static int wo_male_launch(const float* ref, cruse_cplx_layout lref, const float* est, cruse_cplx_layout lest, const float* unproc,
                          cruse_cplx_layout lunp, const float* mask, float* dest, float* loss, void* ws, int B, int T, int F,
                          void* stream) {
    CRUSE_CHECK_ARG(ref && (est || mask) && unproc && loss && ws, "wo_male: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && F > 0, "wo_male: bad sizes B=%d T=%d F=%d", B, T, F);
    const long long total = (long long)B * T * F;
    long long blocks = (long long)B * T;                       // one (b, t) row per CTA iteration
    long long cap = (long long)sm_count() * 8;
    if (cap > LOSS_MAX_PARTS) cap = LOSS_MAX_PARTS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    const double inv = 1.0 / (double)total;
    wo_male_partial_kernel<<<(unsigned)blocks, LOSS_THREADS, 0, st>>>(ref, lref, est, lest, unproc, lunp, dest, (float*)ws, T, F, total,
                                                                      (float)inv, mask, 0, T);
    CRUSE_LAUNCH_OK();
    sum_partials_kernel<<<1, 256, 0, st>>>((const float*)ws, (int)blocks, inv, loss);
    CRUSE_LAUNCH_OK();
    return 0;
}

// The masked loss over the frames [t_begin, t_end) of every utterance: nparts partial sums into ws[p_off, p_off + nparts)
// (no final reduction); cruse_wo_male_finish adds up the partials of all ranges.  Lets the loss follow the decoder range by
// range instead of waiting for the whole mask.
extern "C" int cruse_wo_male_masked_partial_range(const float* ref, cruse_cplx_layout lref, const float* mask, const float* unproc,
                                                  cruse_cplx_layout lunp, void* ws, int p_off, int nparts, int B, int T, int F,
                                                  int t_begin, int t_end, void* stream) {
    CRUSE_CHECK_ARG(ref && mask && unproc && ws, "wo_male_masked_partial_range: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && F > 0 && t_begin >= 0 && t_begin < t_end && t_end <= T, "wo_male_masked_partial_range: bad sizes B=%d T=%d F=%d range [%d,%d)",
                    B, T, F, t_begin, t_end);
    CRUSE_CHECK_ARG(p_off >= 0 && nparts > 0 && p_off + nparts <= LOSS_MAX_PARTS, "wo_male_masked_partial_range: partials [%d,%d) outside the workspace (%d)",
                    p_off, p_off + nparts, LOSS_MAX_PARTS);
    const int Tc = t_end - t_begin;
    long long blocks = (long long)B * Tc;
    if (blocks > nparts) blocks = nparts;
    const long long total = (long long)B * Tc * F;
    cudaStream_t st = (cudaStream_t)stream;
    if (blocks < nparts)      // fewer rows than partial slots: the unused slots must not hold stale values
        CRUSE_CUDA_OK(cudaMemsetAsync((float*)ws + p_off + blocks, 0, sizeof(float) * (size_t)(nparts - blocks), st));
    wo_male_partial_kernel<<<(unsigned)blocks, LOSS_THREADS, 0, st>>>(ref, lref, nullptr, lunp, unproc, lunp, nullptr, (float*)ws + p_off, T, F,
                                                                      total, 0.f, mask, t_begin, Tc);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_wo_male_finish(const void* ws, int nparts, int B, int T, int F, float* loss, void* stream) {
    CRUSE_CHECK_ARG(ws && loss && nparts > 0 && nparts <= LOSS_MAX_PARTS && B > 0 && T > 0 && F > 0, "wo_male_finish: bad arguments");
    sum_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const float*)ws, nparts, 1.0 / ((double)B * T * F), loss);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_wo_male_fwd_bwd(const float* ref, cruse_cplx_layout lref, const float* est, cruse_cplx_layout lest,
                                     const float* unproc, cruse_cplx_layout lunp, float* dest, float* loss, void* ws,
                                     int B, int T, int F, void* stream) {
    CRUSE_CHECK_ARG(est, "wo_male: null pointer");
    return wo_male_launch(ref, lref, est, lest, unproc, lunp, nullptr, dest, loss, ws, B, T, F, stream);
}

extern "C" int cruse_wo_male_masked_fwd(const float* ref, cruse_cplx_layout lref, const float* mask, const float* unproc,
                                        cruse_cplx_layout lunp, float* loss, void* ws, int B, int T, int F, void* stream) {
    CRUSE_CHECK_ARG(mask, "wo_male_masked: null pointer");
    return wo_male_launch(ref, lref, nullptr, lunp, unproc, lunp, mask, nullptr, loss, ws, B, T, F, stream);
}

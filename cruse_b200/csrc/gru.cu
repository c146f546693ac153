// gru.cu -- grouped GRU bottleneck (model/cruse_net.py:14-55 of the reference): input projection
// GEMM, the sequential recurrence, and the LayerNorm that follows each layer.
//
// Replaces 8x nn.GRU(H,H) (cuDNN RNN / ATen on the reference path), torch.chunk/stack/flatten/cat
// (:42-45,48-50) and nn.LayerNorm (:46,51).
//
// Recurrence design (the latency-critical kernel of the whole path):
//   one thread-block CLUSTER of 8 CTAs owns (group g, a slice of 8 utterances) for all T steps.
//   W_hh[g] (3H x H fp32 = 768 KB at H=256) is split by hidden unit across the 8 CTAs and lives
//   in REGISTERS for the whole sequence: a thread holds the r,z,n rows of one hidden unit for a
//   1/8 slice of k (96 weights at H=256).  h_{t-1} of the 8 utterances sits in shared memory
//   (double buffered); each step is 768 FFMA/thread against broadcast float4 reads of h, a
//   3-stage warp-shuffle transpose-reduce over the 8 k-slices that leaves lane `s` holding the
//   gate pre-activations of utterance s, fp32 gate math, and a DSMEM scatter of the new h slice
//   to all 8 CTAs with st.async, whose byte credits complete an mbarrier in each receiving CTA
//   (double-buffered h, no fence and no cluster barrier inside the time loop).  x-projections
//   (and biases) come precomputed from the GEMM and are prefetched one step ahead.
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace cruse {

struct GroupPtrs {
    const float* p[CRUSE_MAX_GROUPS];
};

// ---------------------------------------------------------------------------------------------
// xproj[m, g, n] = sum_k x[m, g*H + k] * w_ih[g][n, k] + b_ih[g][n] + (n < 2H ? b_hh[g][n] : 0)
// fp32 SIMT GEMM, 128x64 tile, 8x4 per thread.  (tensor-core variant: see gru_tc.cu)
// ---------------------------------------------------------------------------------------------
constexpr int GEMM_BM = 128, GEMM_BN = 64, GEMM_BK = 16;

__global__ void __launch_bounds__(256)
gru_ih_gemm_kernel(const float* __restrict__ x, GroupPtrs w_ih, GroupPtrs b_ih, GroupPtrs b_hh,
                   float* __restrict__ xproj, int M, int G, int H) {
    __shared__ __align__(16) float As[GEMM_BK][GEMM_BM + 4];
    __shared__ __align__(16) float Ws[GEMM_BK][GEMM_BN + 4];
    const int g = blockIdx.z;
    const int m0 = blockIdx.x * GEMM_BM, n0 = blockIdx.y * GEMM_BN;
    const int N = 3 * H, K = H, lda = G * H;
    const float* A = x + (size_t)g * H;
    const float* W = w_ih.p[g];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GEMM_BK) {
        // A tile: 128 rows x 16 k -> 512 float4, 2 per thread
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int idx = tid + l * 256;
            const int r = idx >> 2, kq = (idx & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M && k0 + kq < K) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k0 + kq));
            As[kq + 0][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
        }
        {
            const int r = tid >> 2, kq = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < N && k0 + kq < K) v = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + r) * K + k0 + kq));
            Ws[kq + 0][r] = v.x; Ws[kq + 1][r] = v.y; Ws[kq + 2][r] = v.z; Ws[kq + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GEMM_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const float* bi = b_ih.p[g];
    const float* bh = b_hh.p[g];
    float bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        bias[j] = 0.f;
        if (n < N) bias[j] = (bi ? __ldg(bi + n) : 0.f) + ((bh && n < 2 * H) ? __ldg(bh + n) : 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
        float* o = xproj + ((size_t)m * G + g) * N + n0 + tx * 4;
        if (n0 + tx * 4 + 3 < N) {
            *reinterpret_cast<float4*>(o) = make_float4(acc[i][0] + bias[0], acc[i][1] + bias[1], acc[i][2] + bias[2], acc[i][3] + bias[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + tx * 4 + j < N) o[j] = acc[i][j] + bias[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// recurrence
// ---------------------------------------------------------------------------------------------
constexpr int GRU_NC = 8;  // CTAs per cluster (hidden units split 8 ways)
constexpr int GRU_BS = 8;  // utterances per cluster (= k-slices per hidden unit = lanes per unit)

// KPT = k values per thread = hidden units per CTA; padded hidden size HP = 8*KPT >= H.
template <int KPT>
__global__ void __launch_bounds__(KPT * 8, 1)
gru_seq_kernel(const float* __restrict__ xproj, GroupPtrs w_hh, GroupPtrs b_hh, const float* __restrict__ h0,
               float* __restrict__ y, float* __restrict__ hT, int B, int T, int G, int H, int y_fs, int y_gs) {
    constexpr int HP = 8 * KPT;
    constexpr int MQ = KPT / 4;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int g = blockIdx.y, bslice = blockIdx.z;
    __shared__ __align__(16) float hbuf[2][GRU_BS][HP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = lane & 7, ul = lane >> 3;
    const int unit = rank * KPT + warp * 4 + ul;  // hidden index owned by this lane group
    const bool uvalid = unit < H;

    // ---- weights -> registers.  register i of gate q holds W_hh[q*H + unit][((i/4)*8 + slice)*4 + i%4]
    float w[3][KPT];
    {
        const float* W = w_hh.p[g];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int m = 0; m < MQ; ++m) {
                const int k0 = (m * 8 + slice) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (uvalid && k0 < H) v = __ldg(reinterpret_cast<const float4*>(W + ((size_t)q * H + unit) * H + k0));
                w[q][m * 4 + 0] = v.x; w[q][m * 4 + 1] = v.y; w[q][m * 4 + 2] = v.z; w[q][m * 4 + 3] = v.w;
            }
    }
    const float b_hn = (uvalid && b_hh.p[g]) ? __ldg(b_hh.p[g] + 2 * H + unit) : 0.f;

    // ---- initial state + the two "h_{t+1} has landed" mbarriers (one per h buffer)
    __shared__ __align__(8) uint64_t hbar[2];
    for (int i = tid; i < 2 * GRU_BS * HP; i += blockDim.x) (&hbuf[0][0][0])[i] = 0.f;
    if (tid == 0) {
        tc::mbar_init(&hbar[0], 1);
        tc::mbar_init(&hbar[1], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    if (h0) {
        for (int i = tid; i < GRU_BS * H; i += blockDim.x) {
            const int bb = i / H, k = i - bb * H;
            const int bg = bslice * GRU_BS + bb;
            if (bg < B) hbuf[0][bb][k] = __ldg(h0 + ((size_t)g * B + bg) * H + k);
        }
    }
    __syncthreads();
    cluster.sync();  // every CTA's buffers + barriers are initialised before any remote write lands

    const int bg = bslice * GRU_BS + slice;  // after the reduce, lane `slice` owns utterance `slice`
    const bool valid = uvalid && bg < B;
    const size_t N3 = (size_t)3 * H;
    const float* xp = xproj + (((size_t)(valid ? bg : 0) * T) * G + g) * N3 + (uvalid ? unit : 0);
    const size_t xstep = (size_t)G * N3;
    float* yp = y + ((size_t)(valid ? bg : 0) * T) * ((size_t)G * H) + (size_t)(uvalid ? unit : 0) * y_fs + (size_t)g * y_gs;
    const size_t ystep = (size_t)G * H;

    // shared::cluster addresses of "my" h element and of the barriers in each of the 8 CTAs
    uint32_t rem_h[GRU_NC], rem_bar[GRU_NC];
    {
        const uint32_t lh = tc::smem_u32(&hbuf[0][slice][unit]), lb = tc::smem_u32(&hbar[0]);
#pragma unroll
        for (int c = 0; c < GRU_NC; ++c) {
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_h[c]) : "r"(lh), "r"(c));
            asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_bar[c]) : "r"(lb), "r"(c));
        }
    }
    constexpr uint32_t STEP_BYTES = GRU_NC * (KPT * 8) * 4;  // every thread of every CTA sends one float per step

    float xr = 0.f, xz = 0.f, xn = 0.f;
    if (valid && T > 0) { xr = __ldg(xp); xz = __ldg(xp + H); xn = __ldg(xp + 2 * H); }
    float hnew = 0.f;
    int p = 0;
    uint32_t ph0 = 0, ph1 = 0;
    for (int t = 0; t < T; ++t) {
        // h_t is in hbuf[p]: locally initialised for t == 0, otherwise delivered by st.async from all 8 CTAs
        if (t > 0) {
            if (p) { tc::mbar_wait(&hbar[1], ph1); ph1 ^= 1; } else { tc::mbar_wait(&hbar[0], ph0); ph0 ^= 1; }
        }
        if (tid == 0) tc::mbar_expect_tx(&hbar[p ^ 1], STEP_BYTES);  // arm the barrier h_{t+1} will complete
        // prefetch step t+1
        float nxr = 0.f, nxz = 0.f, nxn = 0.f;
        if (valid && t + 1 < T) {
            const float* q = xp + (size_t)(t + 1) * xstep;
            nxr = __ldg(q); nxz = __ldg(q + H); nxn = __ldg(q + 2 * H);
        }
        float acc[3][GRU_BS];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int bb = 0; bb < GRU_BS; ++bb) acc[q][bb] = 0.f;
        const float4* hb = reinterpret_cast<const float4*>(&hbuf[p][0][0]);
#pragma unroll
        for (int m = 0; m < MQ; ++m) {
#pragma unroll
            for (int bb = 0; bb < GRU_BS; ++bb) {
                const float4 hv = hb[bb * (HP / 4) + m * 8 + slice];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    float a = acc[q][bb];
                    a = fmaf(w[q][m * 4 + 0], hv.x, a);
                    a = fmaf(w[q][m * 4 + 1], hv.y, a);
                    a = fmaf(w[q][m * 4 + 2], hv.z, a);
                    a = fmaf(w[q][m * 4 + 3], hv.w, a);
                    acc[q][bb] = a;
                }
            }
        }
        // transpose-reduce over the 8 k-slices: lane `slice` ends with the sums of utterance `slice`
        float s3[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float v4[4], v2[2];
            const bool h4 = slice & 4, h2 = slice & 2, h1 = slice & 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = h4 ? acc[q][i] : acc[q][i + 4];
                const float keep = h4 ? acc[q][i + 4] : acc[q][i];
                v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = h2 ? v4[i] : v4[i + 2];
                const float keep = h2 ? v4[i + 2] : v4[i];
                v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            {
                const float send = h1 ? v2[0] : v2[1];
                const float keep = h1 ? v2[1] : v2[0];
                s3[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
        }
        const float hold = hbuf[p][slice][unit];
        const float r = sigmoidf_(xr + s3[0]);
        const float z = sigmoidf_(xz + s3[1]);
        const float n = tanhf(xn + r * (s3[2] + b_hn));
        hnew = valid ? ((1.f - z) * n + z * hold) : 0.f;
        // scatter h_{t+1}[unit] of utterance `slice` into the other buffer of all 8 CTAs; each store also
        // credits 4 bytes on the destination CTA's barrier, so no fence / cluster barrier is needed per step
        const uint32_t poff = (uint32_t)(p ^ 1) * (GRU_BS * HP * 4), boff = (uint32_t)(p ^ 1) * 8;
#pragma unroll
        for (int c = 0; c < GRU_NC; ++c)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rem_h[c] + poff),
                         "r"(__float_as_uint(hnew)), "r"(rem_bar[c] + boff)
                         : "memory");
        if (valid) yp[(size_t)t * ystep] = hnew;
        p ^= 1;
        xr = nxr; xz = nxz; xn = nxn;
    }
    if (hT && valid) hT[((size_t)g * B + bg) * H + unit] = (T > 0) ? hnew : hbuf[0][slice][unit];
    cluster.sync();  // nobody exits while peers may still be writing into its shared memory
}

template <int KPT>
static int launch_gru_seq(const float* xproj, const GroupPtrs& w_hh, const GroupPtrs& b_hh, const float* h0, float* y,
                          float* hT, int B, int T, int G, int H, int y_fs, int y_gs, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(GRU_NC, G, (B + GRU_BS - 1) / GRU_BS);
    cfg.blockDim = dim3(KPT * 8);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = GRU_NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CRUSE_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_seq_kernel<KPT>, xproj, w_hh, b_hh, h0, y, hT, B, T, G, H, y_fs, y_gs));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, const float* __restrict__ res, float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     long long rows, int D, int T, int t_begin, int Tc) {
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const bool vec = (D & 3) == 0;
    for (long long r = (long long)blockIdx.x * nwarps + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * nwarps) {
        // Tc > 0: the rows are the frames [t_begin, t_begin + Tc) of every utterance of a frame-major [B,T,D] tensor
        const long long row = Tc > 0 ? (r / Tc) * T + t_begin + (r % Tc) : r;
        const float* xr = x + row * D;
        float* yr = y + row * D;
        float s = 0.f;
        if (vec) {
            for (int i = lane * 4; i < D; i += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
                s += (v.x + v.y) + (v.z + v.w);
            }
        } else {
            for (int i = lane; i < D; i += 32) s += __ldg(xr + i);
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
        if (vec) {
            for (int i = lane * 4; i < D; i += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
                const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        } else {
            for (int i = lane; i < D; i += 32) { const float a = __ldg(xr + i) - mean; q += a * a; }
        }
        const float var = warp_sum(q) / (float)D;
        const float rstd = 1.0f / sqrtf(var + eps);
        if (vec) {
            for (int i = lane * 4; i < D; i += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
                const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + i)) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 o;
                o.x = (v.x - mean) * rstd * gm.x + bt.x;
                o.y = (v.y - mean) * rstd * gm.y + bt.y;
                o.z = (v.z - mean) * rstd * gm.z + bt.z;
                o.w = (v.w - mean) * rstd * gm.w + bt.w;
                if (res) {
                    const float4 rv = __ldg(reinterpret_cast<const float4*>(res + row * D + i));
                    o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                }
                *reinterpret_cast<float4*>(yr + i) = o;
            }
        } else {
            for (int i = lane; i < D; i += 32)
                yr[i] = (__ldg(xr + i) - mean) * rstd * (gamma ? __ldg(gamma + i) : 1.f) + (beta ? __ldg(beta + i) : 0.f) +
                        (res ? __ldg(res + row * D + i) : 0.f);
        }
        if (lane == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
    }
}

static int fill_ptrs(GroupPtrs& dst, const float* const* src, int G, bool required, const char* what) {
    for (int i = 0; i < CRUSE_MAX_GROUPS; ++i) dst.p[i] = nullptr;
    if (!src) {
        if (required) { set_error("%s: null pointer table", what); return -1; }
        return 0;
    }
    for (int i = 0; i < G; ++i) {
        if (required && !src[i]) { set_error("%s: null pointer for group %d", what, i); return -1; }
        dst.p[i] = src[i];
    }
    return 0;
}

}  // namespace cruse

using namespace cruse;

extern "C" int cruse_gru_ih_gemm(const float* x, const float* const* w_ih, const float* const* b_ih,
                                 const float* const* b_hh, float* xproj, int M, int G, int H, void* stream) {
    CRUSE_CHECK_ARG(x && xproj, "gru_ih_gemm: null pointer");
    CRUSE_CHECK_ARG(M > 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && (H % 4) == 0, "gru_ih_gemm: bad sizes M=%d G=%d H=%d (H%%4==0, G<=%d)", M, G, H, CRUSE_MAX_GROUPS);
    GroupPtrs pw, pbi, pbh;
    if (fill_ptrs(pw, w_ih, G, true, "gru_ih_gemm w_ih")) return -1;
    if (fill_ptrs(pbi, b_ih, G, false, "gru_ih_gemm b_ih")) return -1;
    if (fill_ptrs(pbh, b_hh, G, false, "gru_ih_gemm b_hh")) return -1;
    dim3 grid((M + GEMM_BM - 1) / GEMM_BM, (3 * H + GEMM_BN - 1) / GEMM_BN, G);
    gru_ih_gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, pw, pbi, pbh, xproj, M, G, H);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_gru_seq_fwd(const float* xproj, const float* const* w_hh, const float* const* b_hh,
                                 const float* h0, float* y, float* hT, int B, int T, int G, int H, int y_fs,
                                 int y_gs, void* stream) {
    CRUSE_CHECK_ARG(xproj && y, "gru_seq_fwd: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T >= 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && (H % 4) == 0, "gru_seq_fwd: bad sizes B=%d T=%d G=%d H=%d", B, T, G, H);
    CRUSE_CHECK_ARG(H <= 256, "gru_seq_fwd: hidden size per group %d > 256 not supported by the register-resident kernel", H);
    GroupPtrs pw, pb;
    if (fill_ptrs(pw, w_hh, G, true, "gru_seq_fwd w_hh")) return -1;
    if (fill_ptrs(pb, b_hh, G, false, "gru_seq_fwd b_hh")) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const int kpt = ((H + 7) / 8 + 3) / 4 * 4;
    switch (kpt) {
        case 4: return launch_gru_seq<4>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 8: return launch_gru_seq<8>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 12: return launch_gru_seq<12>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 16: return launch_gru_seq<16>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 20: return launch_gru_seq<20>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 24: return launch_gru_seq<24>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 28: return launch_gru_seq<28>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        case 32: return launch_gru_seq<32>(xproj, pw, pb, h0, y, hT, B, T, G, H, y_fs, y_gs, st);
        default: set_error("gru_seq_fwd: unsupported H=%d", H); return -1;
    }
}

namespace cruse {
// LayerNorm over a row whose INPUT is the concatenation [G][H] of the grouped GRU outputs and whose OUTPUT is the
// stack(dim=-1)+flatten interleave of model/cruse_net.py:43-45 (feature j = h*G + g), fused: the recurrence kernel
// then stores 16 contiguous bytes per thread and step instead of four 4-byte stores 16 bytes apart.  G == 4.
template <int HPL>      // hidden units per lane = H / 32 (H = 256 -> 8): the whole row lives in registers, one read pass
__global__ void __launch_bounds__(256)
layernorm_interleave4_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, float* __restrict__ y, long long rows, int H) {
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int D = 4 * H;
    for (long long row = (long long)blockIdx.x * nwarps + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * nwarps) {
        const float* xr = x + row * D;
        float* yr = y + row * D;
        float v[HPL][4];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < HPL; ++i) {
            const int h = lane + 32 * i;
#pragma unroll
            for (int g = 0; g < 4; ++g) v[i][g] = (h < H) ? __ldcs(xr + g * H + h) : 0.f;     // read once, streaming
            s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < HPL; ++i) {
            if (lane + 32 * i < H) {
                const float a = v[i][0] - mean, b = v[i][1] - mean, c = v[i][2] - mean, d = v[i][3] - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
        for (int i = 0; i < HPL; ++i) {
            const int h = lane + 32 * i;
            if (h < H) {
                const float4 gm = gamma ? __ldg(reinterpret_cast<const float4*>(gamma + 4 * h)) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 bt = beta ? __ldg(reinterpret_cast<const float4*>(beta + 4 * h)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 o;
                o.x = (v[i][0] - mean) * rstd * gm.x + bt.x;
                o.y = (v[i][1] - mean) * rstd * gm.y + bt.y;
                o.z = (v[i][2] - mean) * rstd * gm.z + bt.z;
                o.w = (v[i][3] - mean) * rstd * gm.w + bt.w;
                *reinterpret_cast<float4*>(yr + 4 * h) = o;
            }
        }
    }
}
}  // namespace cruse

extern "C" int cruse_layernorm_interleave_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y,
                                              long long rows, int D, int G, void* stream) {
    CRUSE_CHECK_ARG(x && y, "layernorm_interleave_fwd: null pointer");
    CRUSE_CHECK_ARG(rows > 0 && D > 0 && G == 4 && D % 4 == 0, "layernorm_interleave_fwd: needs G == 4 and D %% 4 == 0 (rows=%lld D=%d G=%d)", rows, D, G);
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)cruse::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const int H = D / 4;
    CRUSE_CHECK_ARG(H <= 512, "layernorm_interleave_fwd: D = %d too large (D/4 <= 512)", D);
    if (H <= 256)
        cruse::layernorm_interleave4_fwd_kernel<8><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, eps, y, rows, H);
    else
        cruse::layernorm_interleave4_fwd_kernel<16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, eps, y, rows, H);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps,
                                   const float* residual, float* y, float* mean, float* rstd, long long rows, int D,
                                   void* stream) {
    CRUSE_CHECK_ARG(x && y, "layernorm_fwd: null pointer");
    CRUSE_CHECK_ARG(rows > 0 && D > 0, "layernorm_fwd: bad sizes");
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    layernorm_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, eps, residual, y, mean, rstd, rows, D, 0, 0, 0);
    CRUSE_LAUNCH_OK();
    return 0;
}

// the same over the frames [t_begin, t_end) of every utterance of frame-major x / residual / y [B,T,D]
extern "C" int cruse_layernorm_fwd_range(const float* x, const float* gamma, const float* beta, float eps, const float* residual,
                                         float* y, int B, int T, int D, int t_begin, int t_end, void* stream) {
    CRUSE_CHECK_ARG(x && y, "layernorm_fwd_range: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && D > 0 && t_begin >= 0 && t_begin < t_end && t_end <= T, "layernorm_fwd_range: bad sizes B=%d T=%d D=%d range [%d,%d)",
                    B, T, D, t_begin, t_end);
    const long long rows = (long long)B * (t_end - t_begin);
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    layernorm_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, eps, residual, y, nullptr, nullptr, rows, D, T, t_begin,
                                                                             t_end - t_begin);
    CRUSE_LAUNCH_OK();
    return 0;
}

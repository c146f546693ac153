// gru_exact.cu -- exact-fp32 twins of the three tensor-core kernels of the grouped GRU's TRAINING path
// (model/cruse_net.py:23-31,42-50 of the reference; autograd / cuDNN RNN backward on the reference path):
//
//   cruse_gru_seq_fwd_exact  <->  cruse_gru_seq_fwd_tc   (recurrence, saves r, z, n, W_hn.h+b_hn per step)
//   cruse_gru_seq_bwd_exact  <->  cruse_gru_seq_bwd_tc   (backpropagation through time)
//   cruse_gemm_tn_fp32       <->  cruse_gemm_tn_tc       (dx = dxproj . W_ih, dW = dpre^T . h, split-K planes)
//
// Same contracts and buffer layouts, but every product is an fp32 FMA on the CUDA cores (no tf32 operand rounding),
// so that the whole training step can be run in an exact mode (CRUSE_CONV=fp32 CRUSE_GRU_IH=fp32 CRUSE_GRU_SEQ=fp32)
// and compared with the CPU oracle's autograd at fp32 tolerances (SURVEY.md section 8d gradient gate).  These are the
// parity / debugging kernels, not the fast path: one CTA owns (group, 8 utterances) for all T steps and streams
// W_hh from L2 every step; no cluster, no tensor memory.  Fixed summation order: bit-reproducible.
#include "common.cuh"

namespace cruse {

struct ExPtrs {
    const float* p[CRUSE_MAX_GROUPS];
};

constexpr int EX_UB = 8;        // utterances per CTA
constexpr int EX_THREADS = 256; // one thread per hidden unit (H <= 256)

// wt[g][k][r] = w_hh[g][r][k]   (r < 3H, k < H): the forward reads W_hh "k-outer" with coalesced rows
__global__ void __launch_bounds__(256)
gru_exact_transpose_kernel(ExPtrs w, float* __restrict__ wt, int H) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z;
    const float* W = w.p[g];
    float* O = wt + (size_t)g * 3 * H * H;
    const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, k = k0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < 3 * H && k < H) ? W[(size_t)r * H + k] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, r = r0 + threadIdx.x;
        if (k < H && r < 3 * H) O[(size_t)k * 3 * H + r] = tile[threadIdx.x][i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: h_t = (1-z).n + z.h_{t-1};  r = sig(a_r), z = sig(a_z), n = tanh(x_n + r.(W_hn.h + b_hn))
// xproj carries b_ih (all gates) and b_hh (r, z rows) -- the contract of cruse_gru_ih_gemm.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EX_THREADS)
gru_seq_fwd_exact_kernel(const float* __restrict__ xproj, const float* __restrict__ wt, ExPtrs b_hh, const float* __restrict__ h0,
                         float* __restrict__ y, float* __restrict__ hT, float* __restrict__ gates,
                         int B, int T, int G, int H, int y_fs, int y_gs) {
    extern __shared__ __align__(16) float sm[];
    float* hs = sm;                                   // [H][EX_UB]  (k-major: one float4 pair per k)
    const int g = blockIdx.y, b0 = blockIdx.x * EX_UB;
    const int j = threadIdx.x;
    const bool jv = j < H;
    const float* WT = wt + (size_t)g * 3 * H * H;
    const float bn = jv ? __ldg(b_hh.p[g] + 2 * H + j) : 0.f;
    float hreg[EX_UB];
#pragma unroll
    for (int u = 0; u < EX_UB; ++u) {
        const int b = b0 + u;
        hreg[u] = (jv && b < B && h0) ? h0[((size_t)g * B + b) * H + j] : 0.f;
        if (jv) hs[j * EX_UB + u] = hreg[u];
    }
    __syncthreads();
    const size_t g3 = (size_t)G * 3 * H;
    for (int t = 0; t < T; ++t) {
        float ar[EX_UB], az[EX_UB], an[EX_UB];
#pragma unroll
        for (int u = 0; u < EX_UB; ++u) ar[u] = az[u] = an[u] = 0.f;
        if (jv) {
            for (int k = 0; k < H; ++k) {
                const float wr = __ldg(WT + (size_t)k * 3 * H + j);
                const float wz = __ldg(WT + (size_t)k * 3 * H + H + j);
                const float wn = __ldg(WT + (size_t)k * 3 * H + 2 * H + j);
                const float4 ha = *reinterpret_cast<const float4*>(hs + k * EX_UB);
                const float4 hb = *reinterpret_cast<const float4*>(hs + k * EX_UB + 4);
                const float hv[EX_UB] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                for (int u = 0; u < EX_UB; ++u) {
                    ar[u] = fmaf(wr, hv[u], ar[u]);
                    az[u] = fmaf(wz, hv[u], az[u]);
                    an[u] = fmaf(wn, hv[u], an[u]);
                }
            }
        }
        __syncthreads();                               // everybody has read h_{t-1}
        if (jv) {
#pragma unroll
            for (int u = 0; u < EX_UB; ++u) {
                const int b = b0 + u;
                if (b >= B) continue;
                const float* xp = xproj + ((size_t)b * T + t) * g3 + (size_t)g * 3 * H;
                const float r = 1.0f / (1.0f + expf(-(xp[j] + ar[u])));
                const float z = 1.0f / (1.0f + expf(-(xp[H + j] + az[u])));
                const float hn = an[u] + bn;
                const float n = tanhf(fmaf(r, hn, xp[2 * H + j]));
                const float hnew = fmaf(z, hreg[u] - n, n);     // (1-z).n + z.h
                hreg[u] = hnew;
                hs[j * EX_UB + u] = hnew;
                y[((size_t)b * T + t) * ((size_t)G * H) + (size_t)j * y_fs + (size_t)g * y_gs] = hnew;
                if (gates) {
                    float* gp = gates + (((size_t)b * T + t) * G + g) * 4 * (size_t)H;
                    gp[j] = r; gp[H + j] = z; gp[2 * H + j] = n; gp[3 * H + j] = hn;
                }
            }
        }
        __syncthreads();
    }
    if (hT && jv) {
#pragma unroll
        for (int u = 0; u < EX_UB; ++u)
            if (b0 + u < B) hT[((size_t)g * B + b0 + u) * H + j] = hreg[u];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// BPTT.  Per step (t = T-1 .. 0), for the CTA's 8 utterances:
//   dh = dy_t + carry;  dn = dh.(1-z);  da_n = dn.(1-n^2);  da_r = da_n.hn.r.(1-r);  da_z = dh.(h_{t-1}-n).z.(1-z);  dhn = da_n.r
//   carry = dh.z + W_hh^T . (da_r, da_z, dhn)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EX_THREADS)
gru_seq_bwd_exact_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ gates,
                         const float* __restrict__ h0, ExPtrs w_hh, float* __restrict__ dxproj, float* __restrict__ dpre,
                         float* __restrict__ dh0, float* __restrict__ dbias_part, int B, int T, int G, int H, int y_fs, int y_gs) {
    extern __shared__ __align__(16) float sm[];
    float* ds = sm;                                   // [3H][EX_UB]  dpre of this step, row-major over the gate row
    const int g = blockIdx.y, b0 = blockIdx.x * EX_UB;
    const int j = threadIdx.x;
    const bool jv = j < H;
    const float* W = w_hh.p[g];
    float carry[EX_UB], sb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < EX_UB; ++u) carry[u] = 0.f;
    const size_t g3 = (size_t)G * 3 * H, gh = (size_t)G * H;
    for (int t = T - 1; t >= 0; --t) {
        if (jv) {
#pragma unroll
            for (int u = 0; u < EX_UB; ++u) {
                const int b = b0 + u;
                float dar = 0.f, daz = 0.f, dan = 0.f, dhn = 0.f;
                if (b < B) {
                    const size_t bt = (size_t)b * T + t;
                    const float* gp = gates + (bt * G + g) * 4 * (size_t)H;
                    const float r = gp[j], z = gp[H + j], n = gp[2 * H + j], hn = gp[3 * H + j];
                    const size_t yo = (size_t)j * y_fs + (size_t)g * y_gs;
                    const float hp = t > 0 ? y[(bt - 1) * gh + yo] : (h0 ? h0[((size_t)g * B + b) * H + j] : 0.f);
                    const float dh = dy[bt * gh + yo] + carry[u];
                    const float dn = dh * (1.f - z);
                    dan = dn * (1.f - n * n);
                    dar = dan * hn * r * (1.f - r);
                    daz = dh * (hp - n) * z * (1.f - z);
                    dhn = dan * r;
                    carry[u] = dh * z;
                    float* o = dxproj + bt * g3 + (size_t)g * 3 * H;
                    o[j] = dar; o[H + j] = daz; o[2 * H + j] = dan;
                    float* o2 = dpre + bt * g3 + (size_t)g * 3 * H;
                    o2[j] = dar; o2[H + j] = daz; o2[2 * H + j] = dhn;
                    sb[0] += dar; sb[1] += daz; sb[2] += dan; sb[3] += dhn;
                }
                ds[(size_t)j * EX_UB + u] = dar;
                ds[(size_t)(H + j) * EX_UB + u] = daz;
                ds[(size_t)(2 * H + j) * EX_UB + u] = dhn;
            }
        }
        __syncthreads();
        if (jv) {
            float acc[EX_UB];
#pragma unroll
            for (int u = 0; u < EX_UB; ++u) acc[u] = 0.f;
            for (int r = 0; r < 3 * H; ++r) {
                const float w = __ldg(W + (size_t)r * H + j);
                const float4 da = *reinterpret_cast<const float4*>(ds + (size_t)r * EX_UB);
                const float4 db = *reinterpret_cast<const float4*>(ds + (size_t)r * EX_UB + 4);
                const float dv[EX_UB] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
                for (int u = 0; u < EX_UB; ++u) acc[u] = fmaf(w, dv[u], acc[u]);
            }
#pragma unroll
            for (int u = 0; u < EX_UB; ++u) carry[u] += acc[u];
        }
        __syncthreads();
    }
    if (jv) {
        if (dh0) {
#pragma unroll
            for (int u = 0; u < EX_UB; ++u)
                if (b0 + u < B) dh0[((size_t)g * B + b0 + u) * H + j] = carry[u];
        }
        if (dbias_part) {
            // slices of 16 utterances (the tensor-core kernel's layout); two CTAs share a slice -> two commutative adds onto the
            // caller's zeros: the result does not depend on their order
            const int slice = b0 / 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) atomicAdd(dbias_part + (((size_t)slice * G + g) * 4 + q) * H + j, sb[q]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// C_g[m,n] (plane s) = sum over the s-th K range of A_g[m,k] * B_g[n,k]  (+ bias_g[n]);  64x64 tile, 16-deep, 4x4 per thread
// ---------------------------------------------------------------------------------------------------------------
struct GemmPtrs {
    const float* a[CRUSE_MAX_GROUPS];
    const float* b[CRUSE_MAX_GROUPS];
    const float* bias[CRUSE_MAX_GROUPS];
    float* c[CRUSE_MAX_GROUPS];
};

__global__ void __launch_bounds__(256)
gemm_tn_fp32_kernel(GemmPtrs p, int M, int N, int K, long long lda, long long ldb, long long ldc, int splitk, long long c_plane) {
    __shared__ float As[16][64 + 1];
    __shared__ float Bs[16][64 + 1];
    const int g = blockIdx.z / splitk, s = blockIdx.z % splitk;
    const float* A = p.a[g];
    const float* Bm = p.b[g];
    float* Cg = p.c[g] + (size_t)s * c_plane;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const long long kchunk = ((K + splitk - 1) / splitk + 15) / 16 * 16;
    const long long kb = (long long)s * kchunk, ke = (kb + kchunk < K) ? kb + kchunk : K;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) acc[i][jn] = 0.f;
    for (long long k0 = kb; k0 < ke; k0 += 16) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int idx = tid + l * 256;
            const int row = idx >> 4, kk = idx & 15;
            const long long k = k0 + kk;
            As[kk][row] = (m0 + row < M && k < ke) ? __ldg(A + (size_t)(m0 + row) * lda + k) : 0.f;
            Bs[kk][row] = (n0 + row < N && k < ke) ? __ldg(Bm + (size_t)(n0 + row) * ldb + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) b[jn] = Bs[kk][tx * 4 + jn];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(a[i], b[jn], acc[i][jn]);
        }
        __syncthreads();
    }
    const float* bias = p.bias[g];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
            const int n = n0 + tx * 4 + jn;
            if (n < N) Cg[(size_t)m * ldc + n] = acc[i][jn] + (bias ? __ldg(bias + n) : 0.f);
        }
    }
}

static int fill_ptrs(ExPtrs& out, const float* const* in, int G, const char* what) {
    for (int g = 0; g < CRUSE_MAX_GROUPS; ++g) out.p[g] = nullptr;
    for (int g = 0; g < G; ++g) {
        CRUSE_CHECK_ARG(in[g] != nullptr, "%s: null pointer for group %d", what, g);
        out.p[g] = in[g];
    }
    return 0;
}

}  // namespace cruse

using namespace cruse;

extern "C" size_t cruse_gru_exact_ws_bytes(int G, int H) {
    if (G <= 0 || H <= 0) return 0;
    return (size_t)G * 3 * H * H * sizeof(float);
}

extern "C" int cruse_gru_seq_fwd_exact(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h0,
                                       float* y, float* hT, float* gates, void* ws, int B, int T, int G, int H, int y_fs, int y_gs,
                                       void* stream) {
    CRUSE_CHECK_ARG(xproj && w_hh && b_hh && y && ws, "gru_seq_fwd_exact: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && H <= EX_THREADS,
                    "gru_seq_fwd_exact: bad sizes B=%d T=%d G=%d H=%d (H <= %d)", B, T, G, H, EX_THREADS);
    ExPtrs pw, pb;
    if (fill_ptrs(pw, w_hh, G, "gru_seq_fwd_exact w_hh") || fill_ptrs(pb, b_hh, G, "gru_seq_fwd_exact b_hh")) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    float* wt = static_cast<float*>(ws);
    gru_exact_transpose_kernel<<<dim3((H + 31) / 32, (3 * H + 31) / 32, G), dim3(32, 8), 0, st>>>(pw, wt, H);
    CRUSE_LAUNCH_OK();
    const size_t smem = (size_t)H * EX_UB * sizeof(float);
    gru_seq_fwd_exact_kernel<<<dim3((B + EX_UB - 1) / EX_UB, G), EX_THREADS, smem, st>>>(xproj, wt, pb, h0, y, hT, gates, B, T, G, H,
                                                                                        y_fs, y_gs);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_gru_seq_bwd_exact(const float* dy, const float* y, const float* gates, const float* h0,
                                       const float* const* w_hh, float* dxproj, float* dpre, float* dh0, float* dbias_part,
                                       int B, int T, int G, int H, int y_fs, int y_gs, void* stream) {
    CRUSE_CHECK_ARG(dy && y && gates && w_hh && dxproj && dpre, "gru_seq_bwd_exact: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T > 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && H <= EX_THREADS,
                    "gru_seq_bwd_exact: bad sizes B=%d T=%d G=%d H=%d (H <= %d)", B, T, G, H, EX_THREADS);
    ExPtrs pw;
    if (fill_ptrs(pw, w_hh, G, "gru_seq_bwd_exact w_hh")) return -1;
    const size_t smem = (size_t)3 * H * EX_UB * sizeof(float);
    gru_seq_bwd_exact_kernel<<<dim3((B + EX_UB - 1) / EX_UB, G), EX_THREADS, smem, (cudaStream_t)stream>>>(
        dy, y, gates, h0, pw, dxproj, dpre, dh0, dbias_part, B, T, G, H, y_fs, y_gs);
    CRUSE_LAUNCH_OK();
    return 0;
}

extern "C" int cruse_gemm_tn_fp32(const float* const* A, const float* const* Bm, const float* const* bias, float* const* C,
                                  int G, int M, int N, int K, long long lda, long long ldb, long long ldc,
                                  int splitk, long long c_plane, void* stream) {
    CRUSE_CHECK_ARG(A && Bm && C, "gemm_tn_fp32: null pointer table");
    CRUSE_CHECK_ARG(G > 0 && G <= CRUSE_MAX_GROUPS && M > 0 && N > 0 && K > 0 && splitk >= 1, "gemm_tn_fp32: bad sizes G=%d M=%d N=%d K=%d splitk=%d",
                    G, M, N, K, splitk);
    CRUSE_CHECK_ARG(splitk == 1 || bias == nullptr, "gemm_tn_fp32: bias with split-K planes");
    CRUSE_CHECK_ARG((long long)G * splitk <= 65535, "gemm_tn_fp32: G*splitk too large");
    GemmPtrs p;
    for (int g = 0; g < CRUSE_MAX_GROUPS; ++g) p.a[g] = p.b[g] = p.bias[g] = nullptr, p.c[g] = nullptr;
    for (int g = 0; g < G; ++g) {
        CRUSE_CHECK_ARG(A[g] && Bm[g] && C[g], "gemm_tn_fp32: null operand for group %d", g);
        p.a[g] = A[g]; p.b[g] = Bm[g]; p.c[g] = C[g];
        p.bias[g] = bias ? bias[g] : nullptr;
    }
    gemm_tn_fp32_kernel<<<dim3((M + 63) / 64, (N + 63) / 64, G * splitk), 256, 0, (cudaStream_t)stream>>>(p, M, N, K, lda, ldb, ldc, splitk,
                                                                                                         c_plane);
    CRUSE_LAUNCH_OK();
    return 0;
}

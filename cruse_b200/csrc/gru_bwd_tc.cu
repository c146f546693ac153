// gru_bwd_tc.cu -- backpropagation through time of the grouped GRU (cuDNN's
// RNN backward on the reference path for the 8x nn.GRU of model/cruse_net.py:23-31,43-50), on tcgen05.
//
// Walking t = T-1 .. 0, every step needs  dh_{t-1} += W_hh^T . dpre_t  with dpre_t = (da_r, da_z, dhn) the
// gradient w.r.t. the recurrent pre-activations -- a [H x 3H].[3H x N] product on the critical path.
// Same cluster shape as the forward kernel (gru_seq_tc.cu): H/32 CTAs per (group, 16*NI utterances), CTA c
// owns hidden units [32c, 32c+32).  Here the REDUCTION index (gate row j of the CTA's own units) is what a
// CTA holds locally, so each CTA keeps the 96 columns {r,z,n} x {its units} of W_hh^T for ALL H output
// units resident in tensor memory (A operand: 2 M-tiles x 96 k-columns), multiplies them with its own
// freshly computed dpre slice (B operand, written to shared memory by the gate threads -- no exchange
// needed before the MMA), and the H/32 partial products are reduce-scattered over DSMEM: each warp's
// accumulator quadrant belongs to exactly one owner CTA (unit k -> CTA k/32 == warp index) and is sent
// there with st.async + mbarrier byte counting; the owner's gate threads add the H/32 partials.
//
// Outputs per step: dxproj = (da_r, da_z, da_n) (gradient of the input projections incl. b_ih) and
// dpre = (da_r, da_z, dhn); the weight gradients are dense GEMMs over all (b,t) afterwards
// (cruse_gemm_tn_tc on transposed copies, cruse_transpose_*).
#include "common.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace cruse {
namespace {

constexpr int BW_THREADS = 256;
constexpr int BW_U = 32;
constexpr int BW_TMEM_COLS = 512;
constexpr int BW_D_COL = 256;

struct BwPtrs {
    const float* w_hh[CRUSE_MAX_GROUPS];
};

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t sw128_off(int row, int k, int kb_bytes) {
    const int kb = k >> 5, kk = k & 31;
    return (uint32_t)(kb * kb_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7))) << 4) + ((kk & 3) << 2));
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int NC, int NI>
__global__ void __launch_bounds__(BW_THREADS, 1)
gru_bwd_tc_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ gates,
                  const float* __restrict__ h0, const BwPtrs ptrs, float* __restrict__ dxproj, float* __restrict__ dpre,
                  float* __restrict__ dh0, float* __restrict__ dbias_part, int B, int T, int G, int H, int y_fs, int y_gs) {
    constexpr int NB = 16 * NI;                   // utterances per cluster = MMA N
    constexpr int D_KB = NB * 128;               // bytes of one B k-block tile: NB rows x 32 tf32
    constexpr int NMT = (NC * 32 + 127) / 128;   // M tiles of output units
    constexpr int PART_FLOATS = NC * NB * 32;    // one partial buffer: [src][b][unit]
    extern __shared__ uint8_t smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int g = blockIdx.y, bslice = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sD = smem_raw + (base - tc::smem_u32(smem_raw));          // 3 k-blocks x D_KB  (B operand: dpre of my units)
    float* sPart = reinterpret_cast<float*>(sD + 3 * D_KB);            // [2][NC][NB][32]
    uint64_t* pbar = reinterpret_cast<uint64_t*>(sPart + 2 * PART_FLOATS);
    uint64_t* acc_full = pbar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int b = tid >> 4, jp = tid & 15;
    const int u0 = rank * BW_U + 2 * jp;
    const bool uvalid = u0 < H;
    int bg[NI];
    bool valid[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        bg[i] = bslice * NB + b + 16 * i;
        valid[i] = uvalid && bg[i] < B;
    }
    if (tid == 0) {
        tc::mbar_init(&pbar[0], 1);
        tc::mbar_init(&pbar[1], 1);
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 3) tc::tmem_alloc<BW_TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // ---- A operand -> tensor memory: M-tile mt, lane = output unit k - 128*mt, column mt*128 + q*32 + j holds
    //      W_hh[q*H + 32*rank + j][k]  (tf32).  Warp w loads M-tile w/4, lane quadrant w%4.
    if ((warp >> 2) < NMT) {
        const float* W = ptrs.w_hh[g];
        const int mt = warp >> 2;
        const int k = 128 * mt + 32 * (warp & 3) + lane;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int unit = rank * BW_U + j;
                v[j] = (unit < H && k < H) ? to_tf32(__ldg(W + ((size_t)q * H + unit) * H + k)) : 0.f;
            }
            tmem_st_32x32(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(mt * 128 + q * 32), v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    cluster.sync();

    // ---- owner of my warp's accumulator quadrant: units 32*warp .. 32*warp+31 live in CTA `warp`
    const bool sender = warp < NC;
    uint32_t rem_part = 0, rem_bar = 0;
    if (sender) {
        const uint32_t lp = tc::smem_u32(sPart) + (uint32_t)((rank * NB) * 32 + lane) * 4;   // [src = rank][b = 0][unit = lane]
        const uint32_t lb = tc::smem_u32(&pbar[0]);
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_part) : "r"(lp), "r"(warp));
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rem_bar) : "r"(lb), "r"(warp));
    }
    constexpr uint32_t STEP_BYTES = (uint32_t)PART_FLOATS * 4;

    const size_t ystep = (size_t)G * H;
    const size_t g4 = (size_t)G * 4 * H, g3 = (size_t)G * 3 * H;
    const float* dyp[NI];
    const float* ypp[NI];
    const float* gtp[NI];
    float* dxp[NI];
    float* dpp[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const size_t bt = (size_t)(valid[i] ? bg[i] : 0) * T;
        const size_t yo = bt * ystep + (size_t)(uvalid ? u0 : 0) * y_fs + (size_t)g * y_gs;
        dyp[i] = dy + yo;
        ypp[i] = y + yo;
        gtp[i] = gates + (bt * G + g) * (4 * (size_t)H) + (uvalid ? u0 : 0);
        dxp[i] = dxproj + (bt * G + g) * (3 * (size_t)H) + (uvalid ? u0 : 0);
        dpp[i] = dpre + (bt * G + g) * (3 * (size_t)H) + (uvalid ? u0 : 0);
    }
    auto ld2 = [&](const float* p_, int fs) -> float2 {
        if (fs == 1) return __ldg(reinterpret_cast<const float2*>(p_));
        return make_float2(__ldg(p_), __ldg(p_ + fs));
    };
    // per-step inputs, prefetched one step ahead
    float2 c_dy[NI], c_r[NI], c_z[NI], c_n[NI], c_hn[NI], c_hp[NI];
    auto fetch = [&](int t, float2* a_dy, float2* a_r, float2* a_z, float2* a_n, float2* a_hn, float2* a_hp) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            a_dy[i] = a_r[i] = a_z[i] = a_n[i] = a_hn[i] = a_hp[i] = make_float2(0.f, 0.f);
            if (valid[i] && t >= 0) {
                a_dy[i] = ld2(dyp[i] + (size_t)t * ystep, y_fs);
                const float* gq = gtp[i] + (size_t)t * g4;
                a_r[i] = __ldg(reinterpret_cast<const float2*>(gq));
                a_z[i] = __ldg(reinterpret_cast<const float2*>(gq + H));
                a_n[i] = __ldg(reinterpret_cast<const float2*>(gq + 2 * H));
                a_hn[i] = __ldg(reinterpret_cast<const float2*>(gq + 3 * H));
                if (t > 0) a_hp[i] = ld2(ypp[i] + (size_t)(t - 1) * ystep, y_fs);
                else if (h0) a_hp[i] = __ldg(reinterpret_cast<const float2*>(h0 + ((size_t)g * B + bg[i]) * H + u0));
            }
        }
    };
    fetch(T - 1, c_dy, c_r, c_z, c_n, c_hn, c_hp);

    constexpr uint32_t idesc = tc::instr_desc(2 /*tf32*/, 128, NB);
    const uint64_t bdesc0 = tc::smem_desc_sw128(base);
    float2 dhc[NI];
    float2 sb_r = make_float2(0.f, 0.f), sb_z = sb_r, sb_n = sb_r, sb_hn = sb_r;   // bias-gradient sums over t (and my utterances)
#pragma unroll
    for (int i = 0; i < NI; ++i) dhc[i] = make_float2(0.f, 0.f);
    uint32_t ph[2] = {0, 0};

    auto reduce_partials = [&](int buf, float2* out) {
        const float* pb = sPart + (size_t)buf * PART_FLOATS;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            float2 a = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float2 v = *reinterpret_cast<const float2*>(pb + ((c * NB + b + 16 * i) * 32 + 2 * jp));
                a.x += v.x; a.y += v.y;
            }
            out[i] = a;
        }
    };

    for (int s = 0; s < T; ++s) {
        const int t = T - 1 - s;
        const int p = s & 1;
        if (tid == 0) tc::mbar_expect_tx(&pbar[p], STEP_BYTES);      // arm the barrier this step's partials will complete
        float2 dhm[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) dhm[i] = make_float2(0.f, 0.f);
        if (s > 0) {
            tc::mbar_wait(&pbar[p ^ 1], ph[p ^ 1]);                   // W_hh^T . dpre_{t+1}, reduce-scattered to me
            ph[p ^ 1] ^= 1;
            reduce_partials(p ^ 1, dhm);
        }
        // prefetch step t-1
        float2 n_dy[NI], n_r[NI], n_z[NI], n_n[NI], n_hn[NI], n_hp[NI];
        fetch(t - 1, n_dy, n_r, n_z, n_n, n_hn, n_hp);
        // gate gradients
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            float2 dar, daz, dan, dhn;
            {
                const float dh = c_dy[i].x + dhc[i].x + dhm[i].x;
                const float r = c_r[i].x, z = c_z[i].x, n = c_n[i].x;
                const float dn = dh * (1.f - z), dzg = dh * (c_hp[i].x - n);
                dhc[i].x = dh * z;
                dan.x = dn * (1.f - n * n);
                dar.x = dan.x * c_hn[i].x * r * (1.f - r);
                daz.x = dzg * z * (1.f - z);
                dhn.x = dan.x * r;
            }
            {
                const float dh = c_dy[i].y + dhc[i].y + dhm[i].y;
                const float r = c_r[i].y, z = c_z[i].y, n = c_n[i].y;
                const float dn = dh * (1.f - z), dzg = dh * (c_hp[i].y - n);
                dhc[i].y = dh * z;
                dan.y = dn * (1.f - n * n);
                dar.y = dan.y * c_hn[i].y * r * (1.f - r);
                daz.y = dzg * z * (1.f - z);
                dhn.y = dan.y * r;
            }
            // B operand: row = utterance, k = q*32 + local unit
            const int row = b + 16 * i;
            *reinterpret_cast<float2*>(sD + sw128_off(row, 0 * 32 + 2 * jp, D_KB)) = make_float2(to_tf32(dar.x), to_tf32(dar.y));
            *reinterpret_cast<float2*>(sD + sw128_off(row, 1 * 32 + 2 * jp, D_KB)) = make_float2(to_tf32(daz.x), to_tf32(daz.y));
            *reinterpret_cast<float2*>(sD + sw128_off(row, 2 * 32 + 2 * jp, D_KB)) = make_float2(to_tf32(dhn.x), to_tf32(dhn.y));
            sb_r.x += dar.x; sb_r.y += dar.y; sb_z.x += daz.x; sb_z.y += daz.y;
            sb_n.x += dan.x; sb_n.y += dan.y; sb_hn.x += dhn.x; sb_hn.y += dhn.y;
            if (valid[i]) {
                float* o = dxp[i] + (size_t)t * g3;
                *reinterpret_cast<float2*>(o) = dar;
                *reinterpret_cast<float2*>(o + H) = daz;
                *reinterpret_cast<float2*>(o + 2 * H) = dan;
                float* o2 = dpp[i] + (size_t)t * g3;
                *reinterpret_cast<float2*>(o2) = dar;
                *reinterpret_cast<float2*>(o2 + H) = daz;
                *reinterpret_cast<float2*>(o2 + 2 * H) = dhn;
            }
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        if (warp == 3) {
            tc::tc_fence_after();
            if (tc::elect_one()) {
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_tf32_ts(tmem_base + BW_D_COL + mt * NB, tmem_base + (uint32_t)(mt * 128 + q * 32 + ks * 8),
                                         bdesc0 + (uint64_t)((q * D_KB + ks * 32) >> 4), idesc, (q | ks) ? 1u : 0u);
                tc::umma_commit(acc_full);
            }
            __syncwarp();
        }
        tc::mbar_wait(acc_full, (uint32_t)(s & 1));
        tc::tc_fence_after();
        if (sender) {
            // my warp's quadrant of M-tile warp/4: lane = unit 32*warp + lane, columns = utterances -> owner CTA `warp`
            const uint32_t poff = (uint32_t)p * (PART_FLOATS * 4), boff = (uint32_t)p * 8;
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                float v[16];
                tmem_ld_32x16(tmem_base + BW_D_COL + (warp >> 2) * NB + 16 * i + ((uint32_t)(32 * (warp & 3)) << 16), v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(
                                     rem_part + poff + (uint32_t)((16 * i + c) * 32 * 4)),
                                 "r"(__float_as_uint(v[c])), "r"(rem_bar + boff)
                                 : "memory");
            }
        }
        tc::tc_fence_before();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            c_dy[i] = n_dy[i]; c_r[i] = n_r[i]; c_z[i] = n_z[i]; c_n[i] = n_n[i]; c_hn[i] = n_hn[i]; c_hp[i] = n_hp[i];
        }
    }
    // ---- gradient w.r.t. the initial state (also drains the last reduce-scatter before anybody exits)
    if (T > 0) {
        const int pl = (T - 1) & 1;
        tc::mbar_wait(&pbar[pl], ph[pl]);
        float2 dhm[NI];
        reduce_partials(pl, dhm);
        if (dh0) {
#pragma unroll
            for (int i = 0; i < NI; ++i)
                if (valid[i])
                    *reinterpret_cast<float2*>(dh0 + ((size_t)g * B + bg[i]) * H + u0) = make_float2(dhc[i].x + dhm[i].x, dhc[i].y + dhm[i].y);
        }
    }
    // ---- bias gradients: sum over the cluster's utterances -> one partial row per (slice, group).  Every thread leaves its
    // sums (already over its NI utterances) in shared memory, 128 threads add the 16 utterance rows in order: fixed order, no
    // shared-memory float atomics (threads whose units / utterances do not exist carry zeros)
    __syncthreads();
    // [16 utterance rows][4 gates x 32 units] = 8 KB over the operand tiles and the partial buffers behind them (>= 10 KB in the
    // smallest instantiation): every MMA has completed and no remote write is outstanding any more
    float* sb = reinterpret_cast<float*>(sD);
    {
        float* q = sb + b * 128 + 2 * jp;
        q[0 * 32] = sb_r.x;  q[0 * 32 + 1] = sb_r.y;
        q[1 * 32] = sb_z.x;  q[1 * 32 + 1] = sb_z.y;
        q[2 * 32] = sb_n.x;  q[2 * 32 + 1] = sb_n.y;
        q[3 * 32] = sb_hn.x; q[3 * 32 + 1] = sb_hn.y;
    }
    __syncthreads();
    if (dbias_part && tid < 128) {
        const int q = tid >> 5, unit = rank * BW_U + (tid & 31);
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) a += sb[r * 128 + tid];
        if (unit < H) dbias_part[(((size_t)bslice * G + g) * 4 + q) * H + unit] = a;
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster.sync();
    if (warp == 3) tc::tmem_dealloc<BW_TMEM_COLS>(tmem_base);
}

constexpr size_t bwd_smem_bytes(int NC, int NI) {
    // The kernel holds all 512 TMEM columns of its SM.  Requesting at least 136 KB keeps every other tensor-core CTA of this
    // library (GEMM 99.5 KB, conv stages > 100 KB) off the SM: one that were co-scheduled (the weight-gradient kernels run
    // beside the BPTT on a side stream) would sit in tcgen05.alloc until the whole sequence has been processed.
    const size_t need = 1024 + 3 * (size_t)(16 * NI * 128) + 2 * (size_t)NC * 16 * NI * 32 * 4 + 64;
    return need < 136 * 1024 ? 136 * 1024 : need;
}

template <int NC, int NI>
void bwd_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int G, int nslices, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(NC, G, nslices);
    cfg.blockDim = dim3(BW_THREADS);
    cfg.dynamicSmemBytes = bwd_smem_bytes(NC, NI);
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

template <int NC, int NI>
int bwd_max_clusters() {
    if (cudaFuncSetAttribute(gru_bwd_tc_kernel<NC, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes(NC, NI)) != cudaSuccess) return -2;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    bwd_cfg<NC, NI>(cfg, attr, 1, 1024, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gru_bwd_tc_kernel<NC, NI>, &cfg) != cudaSuccess) return -2;
    return n;
}

template <int NC, int NI>
int launch_bwd(const float* dy, const float* y, const float* gates, const float* h0, const BwPtrs& ptrs, float* dxproj,
               float* dpre, float* dh0, float* dbias_part, int B, int T, int G, int H, int y_fs, int y_gs, cudaStream_t st) {
    CRUSE_CUDA_OK(cudaFuncSetAttribute(gru_bwd_tc_kernel<NC, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes(NC, NI)));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    bwd_cfg<NC, NI>(cfg, attr, G, (B + 16 * NI - 1) / (16 * NI), st);
    CRUSE_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_bwd_tc_kernel<NC, NI>, dy, y, gates, h0, ptrs, dxproj, dpre, dh0, dbias_part, B, T, G, H, y_fs, y_gs));
    return 0;
}

template <int NC>
int launch_bwd_nc(const float* dy, const float* y, const float* gates, const float* h0, const BwPtrs& ptrs, float* dxproj,
                  float* dpre, float* dh0, float* dbias_part, int B, int T, int G, int H, int y_fs, int y_gs, cudaStream_t st) {
    static thread_local int cap16 = -1;
    if (cap16 < 0) cap16 = bwd_max_clusters<NC, 1>();
    const int need16 = G * ((B + 15) / 16);
    if (cap16 > 0 && need16 > cap16) return launch_bwd<NC, 2>(dy, y, gates, h0, ptrs, dxproj, dpre, dh0, dbias_part, B, T, G, H, y_fs, y_gs, st);
    return launch_bwd<NC, 1>(dy, y, gates, h0, ptrs, dxproj, dpre, dh0, dbias_part, B, T, G, H, y_fs, y_gs, st);
}

// ---------------------------------------------------------------------------------------------
// transposes that put the (b,t) reduction index innermost for the weight-gradient GEMMs:
//   out[g][c][m] = in[m * ld + g * gs + c * cs]     m < M (= B*T), c < Cn
// with an optional one-step time shift (h_{t-1} from y: m = b*T + t reads row m-1, t == 0 reads h0 or 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_gcm_kernel(const float* __restrict__ in, const float* __restrict__ h0, float* __restrict__ out, long long M, int Cn,
                     long long ld, long long gs, long long cs, int shift_T, int Bn, long long ldo) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z;
    const long long m0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const long long m = m0 + r;
        const int c = c0 + tx;
        float v = 0.f;
        if (m < M && c < Cn) {
            if (shift_T > 0) {
                const long long bb = m / shift_T;
                const int t = (int)(m - bb * shift_T);
                if (t > 0) v = __ldg(in + (m - 1) * ld + g * gs + c * cs);
                else if (h0) v = __ldg(h0 + ((size_t)g * Bn + bb) * Cn + c);
            } else {
                v = __ldg(in + m * ld + g * gs + c * cs);
            }
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        const long long m = m0 + tx;
        if (c < Cn && m < M) out[((size_t)g * Cn + c) * ldo + m] = tile[tx][r];
    }
}

}  // namespace
}  // namespace cruse

using namespace cruse;

#define BW_DISPATCH(nc, CALL)                     \
    switch (nc) {                                 \
        case 1: return CALL(1);                   \
        case 2: return CALL(2);                   \
        case 3: return CALL(3);                   \
        case 4: return CALL(4);                   \
        case 5: return CALL(5);                   \
        case 6: return CALL(6);                   \
        case 7: return CALL(7);                   \
        case 8: return CALL(8);                   \
        default: break;                           \
    }

extern "C" int cruse_gru_seq_bwd_tc(const float* dy, const float* y, const float* gates, const float* h0,
                                    const float* const* w_hh, float* dxproj, float* dpre, float* dh0, float* dbias_part,
                                    int B, int T, int G, int H, int y_fs, int y_gs, void* stream) {
    CRUSE_CHECK_ARG(dy && y && gates && w_hh && dxproj && dpre, "gru_seq_bwd_tc: null pointer");
    CRUSE_CHECK_ARG(B > 0 && T >= 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && (H % 4) == 0 && H <= 256,
                    "gru_seq_bwd_tc: bad sizes B=%d T=%d G=%d H=%d (H%%4==0, H<=256, G<=%d)", B, T, G, H, CRUSE_MAX_GROUPS);
    BwPtrs ptrs;
    for (int i = 0; i < CRUSE_MAX_GROUPS; ++i) ptrs.w_hh[i] = nullptr;
    for (int i = 0; i < G; ++i) {
        CRUSE_CHECK_ARG(w_hh[i], "gru_seq_bwd_tc: null weight pointer for group %d", i);
        ptrs.w_hh[i] = w_hh[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(N) launch_bwd_nc<N>(dy, y, gates, h0, ptrs, dxproj, dpre, dh0, dbias_part, B, T, G, H, y_fs, y_gs, st)
    BW_DISPATCH((H + BW_U - 1) / BW_U, CALL)
#undef CALL
    set_error("gru_seq_bwd_tc: unsupported H=%d", H);
    return -1;
}

extern "C" int cruse_transpose_gcm(const float* in, const float* h0, float* out, long long M, int G, int Cn, long long ld,
                                   long long gs, long long cs, int shift_T, int Bn, long long ldo, void* stream) {
    CRUSE_CHECK_ARG(in && out, "transpose_gcm: null pointer");
    CRUSE_CHECK_ARG(M > 0 && G > 0 && Cn > 0 && shift_T >= 0 && ldo >= M, "transpose_gcm: bad sizes");
    CRUSE_CHECK_ARG(shift_T == 0 || (M % shift_T) == 0, "transpose_gcm: M must be a multiple of T for the time shift");
    dim3 grid((unsigned)((M + 31) / 32), (Cn + 31) / 32, G);
    transpose_gcm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, h0, out, M, Cn, ld, gs, cs, shift_T, Bn, ldo);
    CRUSE_LAUNCH_OK();
    return 0;
}

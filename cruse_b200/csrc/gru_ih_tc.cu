// gru_ih_tc.cu -- the GRU input projections ("the GRU's three dense matmuls", W_ir|W_iz|W_in stacked
// as weight_ih_l0 [3H, H]) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// accumulator in TMEM, operands staged by TMA.
//
// Replaces the input half of 8x nn.GRU (model/cruse_net.py:23-31,43-50; cuDNN's RNN input GEMM on the
// reference path):   xproj[m, g, n] = sum_k x[m, g*H + k] * w_ih[g][n, k] + b_ih[g][n] (+ b_hh[g][n], n < 2H)
//
// Mapping: one CTA per (n-tile of 256 gate rows, m-tile of 128 frames, group).  A = 128 frames x 32 k
// and B = 256 weight rows x 32 k fp32 boxes arrive by TMA in the 128-byte-swizzled K-major layout the
// UMMA shared-memory descriptors address directly (both operands are K-major in HBM already: x rows and
// PyTorch's [3H, H] weight rows), rounded fp32 -> tf32 by the TMA unit (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32).
// Warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer (4 x m128 n256 k8 per k-block),
// warps 2-5 = epilogue (tcgen05.ld 32 lanes x 32 columns, + bias, 128-bit stores).  Two CTAs fit per SM
// (2 x 96 KB stages, 2 x 256 TMEM columns), so one CTA's epilogue overlaps the other's main loop.
// Out-of-range frames / k / gate rows are zero-filled by TMA and masked in the epilogue, so any M and
// any H % 4 == 0 work (config R: H = 176).
//
// The same kernel is exported as a general "TN" GEMM  C[m,n] = sum_k A[m,k] B[n,k]  (both operands K-major,
// optional split-K into partial planes) for the GRU backward: dx = dxproj . W_ih, dW = dpre^T . h (a9).
#include "common.cuh"
#include "tc_common.cuh"

namespace cruse {

int make_tmap_2d(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows,
                 bool as_tf32, bool atom32) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static thread_local EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
            q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point not available");
            return -2;
        }
        fn = (EncodeFn)p;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_bytes & 15)) {
        set_error("TMA needs a 16-byte aligned base (%p) and row pitch (%llu B)", (const void*)base, (unsigned long long)pitch_bytes);
        return -1;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, as_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu pitch=%llu)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes);
        return -2;
    }
    return 0;
}

namespace {

constexpr int BM = 128, BN = 256, BK = 32;          // tile; BK fp32 = one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int IH_THREADS = 192;
constexpr int IH_SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BN * 4 /*bias*/ + 64 /*barriers*/;

// G independent problems  C_g[m, n] = sum_k A_g[m, k] * B_g[n, k] + bias1_g[n] + (n < bias2_rows ? bias2_g[n] : 0)
struct GemmArgs {
    CUtensorMap a[CRUSE_MAX_GROUPS];
    CUtensorMap b[CRUSE_MAX_GROUPS];
    const float* bias1[CRUSE_MAX_GROUPS];
    const float* bias2[CRUSE_MAX_GROUPS];
    const float* addend[CRUSE_MAX_GROUPS];      // optional [M, N] matrix with C's pitch, added in the epilogue (split-K: plane 0 only)
    float* out[CRUSE_MAX_GROUPS];
};

// A_MN / B_MN: that operand is "MN-major" in HBM -- stored [K rows][M or N contiguous], the layout a weight gradient's
// operands have naturally (dW = dpre^T . h: both factors are [B*T, features] row-major, the reduction index is the ROW).
// tcgen05 reads such a tile through an MN-major shared-memory descriptor, so no transposed copy is made: the tile arrives as
// BM/32 (BN/32) TMA boxes of {32 features, BK rows}, each a column of 4-row x 128-byte swizzle atoms (the one MN-major
// layout tcgen05 has for 32-bit elements, 128-byte swizzle with 32-byte atoms: leading byte offset = one box = BK*128 B,
// stride byte offset = one atom = 512 B, two atoms per UMMA_K = 8 rows).
// b_kshift: B's row for reduction index k is k - b_kshift (rows < 0 read as zero): h_{t-1} paired with frame t.
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(IH_THREADS)
gemm_tn_tc_kernel(const __grid_constant__ GemmArgs args, int M, int N, int K, long long ldc, int bias2_rows, int splitk,
                  long long c_plane, int tm_T, int b_kshift) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B tiles: 1024-byte aligned
    uint8_t* tiles = smem_raw + (base - tc::smem_u32(smem_raw));
    float* s_bias = reinterpret_cast<float*>(tiles + STAGES * STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(s_bias + BN);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int g = blockIdx.z / splitk, split = blockIdx.z - g * splitk;
    const int nkb_all = (K + BK - 1) / BK;
    const int kb_begin = (int)((long long)nkb_all * split / splitk), kb_end = (int)((long long)nkb_all * (split + 1) / splitk);
    const int nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&args.a[g]);
        tc::tma_prefetch_desc(&args.b[g]);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<BN>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                tc::mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
                tc::mbar_expect_tx(&full[s], STAGE_BYTES);
                uint8_t* st = tiles + s * STAGE_BYTES;
                const int k0 = (kb_begin + i) * BK;
                if constexpr (A_MN) {
#pragma unroll
                    for (int c = 0; c < BM / 32; ++c) tc::tma_load_2d(st + c * (BK * 128), &args.a[g], m0 + c * 32, k0, &full[s]);
                } else {
                    tc::tma_load_2d(st, &args.a[g], k0, m0, &full[s]);
                }
                if constexpr (B_MN) {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c)
                        tc::tma_load_2d(st + A_BYTES + c * (BK * 128), &args.b[g], n0 + c * 32, k0 - b_kshift, &full[s]);
                } else {
                    tc::tma_load_2d(st + A_BYTES, &args.b[g], k0, n0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane of a converged warp) =====
        constexpr uint32_t idesc = tc::instr_desc(2 /*tf32*/, BM, BN, A_MN, B_MN);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES;
            tc::mbar_wait(&full[s], (i / STAGES) & 1);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {   // UMMA_K = 8 for tf32: 32 bytes along a K-major row / one 8-row atom of an MN-major box
                    const uint64_t da = A_MN ? tc::smem_desc_mn_sw128_32b(sa + k * 1024, BK * 128) : tc::smem_desc_sw128(sa + k * 32);
                    const uint64_t db = B_MN ? tc::smem_desc_mn_sw128_32b(sb + k * 1024, BK * 128) : tc::smem_desc_sw128(sb + k * 32);
                    tc::umma_tf32(tmem_d, da, db, idesc, (i | k) ? 1u : 0u);
                }
                tc::umma_commit(&empty[s]);           // frees the stage when these MMAs have read it
            }
            __syncwarp();
        }
        if (tc::elect_one()) tc::umma_commit(acc_full);   // accumulator complete
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int quad = warp & 3;
        const float* b1 = (split == 0) ? args.bias1[g] : nullptr;
        const float* b2 = (split == 0) ? args.bias2[g] : nullptr;
        for (int i = threadIdx.x - 64; i < BN; i += 128) {
            const int n = n0 + i;
            float b = 0.f;
            if (n < N) b = (b1 ? __ldg(b1 + n) : 0.f) + ((b2 && n < bias2_rows) ? __ldg(b2 + n) : 0.f);
            s_bias[i] = b;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
        const int m = m0 + quad * 32 + lane;
        // tm_T > 0: A rows are (utterance b, frame t) = b*tm_T + t, C rows are written time-major (t*B + b)
        size_t mrow = (size_t)m;
        if (tm_T > 0) { const int bq = m / tm_T; mrow = (size_t)(m - bq * tm_T) * (size_t)(M / tm_T) + bq; }
        float* orow = args.out[g] + (size_t)split * c_plane + mrow * ldc + n0;
        const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(args.out[g]) & 15) == 0) && ((c_plane & 3) == 0);
        if (nkb > 0) {
            tc::mbar_wait(acc_full, 0);
            tc::tc_fence_after();
        }
        // The accumulator comes out of TMEM with lane = row: storing it directly would make every 16-byte store of a warp hit
        // 32 different rows (ldc floats apart) -- half-sector strided writes.  Instead each warp transposes its 32 x 32 chunk
        // through shared memory (the operand stages are free once the accumulator is complete; pitch 36 floats keeps both the
        // 16-byte stores by row and the 16-byte loads by column group bank-conflict free) and writes 4 full 128-byte row
        // segments per instruction.
        float* stg = reinterpret_cast<float*>(tiles) + quad * (32 * 36);
        const int rr = lane >> 3, cc = (lane & 7) * 4;           // store phase: lane -> (row rr + 4i, columns cc..cc+3)
        float* obase = args.out[g] + (size_t)split * c_plane + n0;
        const float* abase = (split == 0 && args.addend[g]) ? args.addend[g] + n0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            if (n0 + c >= N) break;                      // warp-uniform
            // the addend's eight 16-byte pieces of this chunk are requested BEFORE the accumulator is read and transposed, all at
            // once: loaded inside the store loop they serialise behind each store (the compiler may not move a load over a
            // store through an unrelated float pointer) -- measured 132 instead of 64 us for the whole launch
            float4 ad[8];
            const bool vec_chunk = vec_ok && n0 + c + 32 <= N;
            if (abase && vec_chunk) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int mm = m0 + quad * 32 + rr + 4 * i;
                    size_t mr = (size_t)(mm < M ? mm : 0);
                    if (tm_T > 0) { const int bq = (int)mr / tm_T; mr = (size_t)((int)mr - bq * tm_T) * (size_t)(M / tm_T) + bq; }
                    ad[i] = __ldg(reinterpret_cast<const float4*>(abase + mr * ldc + c + cc));
                }
            }
            float v[32];
            if (nkb > 0) {
                tc::tmem_ld_32x32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
                tc::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (vec_chunk) {
                __syncwarp();                               // the previous chunk's loads from the staging tile are done
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(v[j] + s_bias[c + j], v[j + 1] + s_bias[c + j + 1],
                                                                                   v[j + 2] + s_bias[c + j + 2], v[j + 3] + s_bias[c + j + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rr + 4 * i;
                    const int mm = m0 + quad * 32 + r;
                    if (mm < M) {
                        size_t mr = (size_t)mm;
                        if (tm_T > 0) { const int bq = mm / tm_T; mr = (size_t)(mm - bq * tm_T) * (size_t)(M / tm_T) + bq; }
                        float4 o = *reinterpret_cast<const float4*>(stg + r * 36 + cc);
                        if (abase) { o.x += ad[i].x; o.y += ad[i].y; o.z += ad[i].z; o.w += ad[i].w; }
                        *reinterpret_cast<float4*>(obase + mr * ldc + c + cc) = o;
                    }
                }
            } else if (m < M) {                            // ragged N edge / unaligned C: direct stores
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (vec_ok && !abase && n0 + c + j + 3 < N) {
                        *reinterpret_cast<float4*>(orow + c + j) = make_float4(v[j] + s_bias[c + j], v[j + 1] + s_bias[c + j + 1],
                                                                               v[j + 2] + s_bias[c + j + 2], v[j + 3] + s_bias[c + j + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n0 + c + j + e < N) orow[c + j + e] = v[j + e] + s_bias[c + j + e] + (abase ? __ldg(abase + mrow * ldc + c + j + e) : 0.f);
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<BN>(tmem_d);
}

// ---------------------------------------------------------------------------------------------------------------------
// A-stationary variant for the input projections (K = 256): one CTA per (m-tile of 128 rows, group) computes ALL N <= 768
// gate rows.  The activation tile (128 x 256 fp32 = 128 KB) is brought in ONCE by eight TMA boxes that are all in flight
// together and stays resident; the weights stream through a 4-stage ring of 128 x 32 boxes (they are L2 resident: 3 MB per
// layer), n-tile by n-tile; the accumulator is double buffered in tensor memory (2 x 128 columns), so the epilogue of n-tile j
// (tcgen05.ld -> + bias -> transpose through shared memory -> full 128-byte row segments) runs under the MMAs of n-tile j+1.
// Against the one-tile-per-CTA kernel above this removes the per-tile prologue (barrier init, TMEM allocation, descriptor
// fetch, pipeline fill), re-reads the activations once instead of three times, and occupies M/128 x G SMs (64 for a 63-frame
// chunk of 32 utterances) instead of the whole GPU -- what runs beside the recurrences is capacity bound.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AS_BN = 128, AS_NKB = 8, AS_STAGES = 4;
constexpr int AS_A_KB_BYTES = BM * BK * 4;                       // one k-block of the resident A tile: 16 KB
constexpr int AS_A_BYTES = AS_NKB * AS_A_KB_BYTES;               // 128 KB
constexpr int AS_W_STAGE = AS_BN * BK * 4;                       // 16 KB
constexpr int AS_STG_BYTES = 4 * 32 * 36 * 4;                    // epilogue transpose pads (A stays live, so they have their own space)
constexpr int AS_MAX_N = 768;
constexpr int AS_SMEM = 1024 + AS_A_BYTES + AS_STAGES * AS_W_STAGE + AS_STG_BYTES + AS_MAX_N * 4 + 128;

__global__ void __launch_bounds__(IH_THREADS, 1)
gemm_astat_tc_kernel(const __grid_constant__ GemmArgs args, int M, int N, long long ldc, int bias2_rows, int tm_T) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sA = smem_raw + (base - tc::smem_u32(smem_raw));
    uint8_t* sW = sA + AS_A_BYTES;
    float* stg_all = reinterpret_cast<float*>(sW + AS_STAGES * AS_W_STAGE);
    float* s_bias = stg_all + AS_STG_BYTES / 4;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(s_bias + AS_MAX_N);
    uint64_t* full = a_full + 1;
    uint64_t* empty = full + AS_STAGES;
    uint64_t* tmem_full = empty + AS_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, g = blockIdx.y;
    const int NT = N / AS_BN;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&args.a[g]);
        tc::tma_prefetch_desc(&args.b[g]);
        tc::mbar_init(a_full, 1);
        for (int s = 0; s < AS_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 128); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<2 * AS_BN>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            tc::mbar_expect_tx(a_full, AS_A_BYTES);
            for (int kb = 0; kb < AS_NKB; ++kb) tc::tma_load_2d(sA + kb * AS_A_KB_BYTES, &args.a[g], kb * BK, m0, a_full);
            int i = 0;
            for (int nt = 0; nt < NT; ++nt)
                for (int kb = 0; kb < AS_NKB; ++kb, ++i) {
                    const int s = i % AS_STAGES;
                    tc::mbar_wait(&empty[s], ((i / AS_STAGES) & 1) ^ 1);
                    tc::mbar_expect_tx(&full[s], AS_W_STAGE);
                    tc::tma_load_2d(sW + s * AS_W_STAGE, &args.b[g], kb * BK, nt * AS_BN, &full[s]);
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = tc::instr_desc(2 /*tf32*/, BM, AS_BN);
        tc::mbar_wait(a_full, 0);
        tc::tc_fence_after();
        int i = 0;
        for (int nt = 0; nt < NT; ++nt) {
            const int acc = nt & 1;
            tc::mbar_wait(&tmem_empty[acc], ((nt >> 1) & 1) ^ 1);      // the epilogue has drained this accumulator buffer
            tc::tc_fence_after();
            for (int kb = 0; kb < AS_NKB; ++kb, ++i) {
                const int s = i % AS_STAGES;
                tc::mbar_wait(&full[s], (i / AS_STAGES) & 1);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t sa = base + kb * AS_A_KB_BYTES, sb = base + AS_A_BYTES + s * AS_W_STAGE;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)
                        tc::umma_tf32(tmem_d + (uint32_t)(acc * AS_BN), tc::smem_desc_sw128(sa + k * 32), tc::smem_desc_sw128(sb + k * 32), idesc,
                                      (kb | k) ? 1u : 0u);
                    tc::umma_commit(&empty[s]);
                    if (kb == AS_NKB - 1) tc::umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int quad = warp & 3;
        const float* b1 = args.bias1[g];
        const float* b2 = args.bias2[g];
        for (int i = threadIdx.x - 64; i < N; i += 128) {
            float b = (b1 ? __ldg(b1 + i) : 0.f) + ((b2 && i < bias2_rows) ? __ldg(b2 + i) : 0.f);
            s_bias[i] = b;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float* stg = stg_all + quad * (32 * 36);
        const int rr = lane >> 3, cc = (lane & 7) * 4;
        float* obase = args.out[g];
        // output rows of this warp's four store phases (tm_T > 0: rows (b, t) are written time-major, t*B + b)
        size_t orow[8];
        bool ok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int mm = m0 + quad * 32 + rr + 4 * i;
            ok[i] = mm < M;
            size_t mr = (size_t)(ok[i] ? mm : 0);
            if (tm_T > 0) { const int bq = (int)mr / tm_T; mr = (size_t)((int)mr - bq * tm_T) * (size_t)(M / tm_T) + bq; }
            orow[i] = mr * (size_t)ldc;
        }
        for (int nt = 0; nt < NT; ++nt) {
            const int acc = nt & 1;
            tc::mbar_wait(&tmem_full[acc], (nt >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < AS_BN; c += 32) {
                float v[32];
                tc::tmem_ld_32x32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * AS_BN + c), v);
                tc::tmem_ld_wait();
                if (c == AS_BN - 32) {                       // last read of this buffer: hand it back before the stores
                    tc::tc_fence_before();
                    tc::mbar_arrive_relaxed(&tmem_empty[acc]);
                }
                const int n = nt * AS_BN + c;
                __syncwarp();                                // the previous chunk's loads from the staging tile are done
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * 36 + j) = make_float4(v[j] + s_bias[n + j], v[j + 1] + s_bias[n + j + 1],
                                                                                   v[j + 2] + s_bias[n + j + 2], v[j + 3] + s_bias[n + j + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (ok[i]) *reinterpret_cast<float4*>(obase + orow[i] + n + cc) = *reinterpret_cast<const float4*>(stg + (rr + 4 * i) * 36 + cc);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc<2 * AS_BN>(tmem_d);
}

// process-wide switch (developer A/B): 1 = use the A-stationary kernel where it applies (default), 0 = the tile-per-CTA kernel
static int g_astat = 1;

template <bool A_MN, bool B_MN>
int launch_gemm_tile(const GemmArgs& args, int G, int M, int N, int K, long long ldc, int bias2_rows, int splitk, long long c_plane,
                     cudaStream_t st, int tm_T, int b_kshift) {
    CRUSE_CUDA_OK(cudaFuncSetAttribute(gemm_tn_tc_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, IH_SMEM));
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, G * splitk);
    gemm_tn_tc_kernel<A_MN, B_MN><<<grid, IH_THREADS, IH_SMEM, st>>>(args, M, N, K, ldc, bias2_rows, splitk, c_plane, tm_T, b_kshift);
    CRUSE_LAUNCH_OK();
    return 0;
}

int launch_gemm(GemmArgs& args, int G, int M, int N, int K, long long ldc, int bias2_rows, int splitk, long long c_plane,
                cudaStream_t st, int tm_T = 0, bool astat_ok = false, bool a_mn = false, bool b_mn = false, int b_kshift = 0) {
    for (int g = G; g < CRUSE_MAX_GROUPS; ++g) {
        args.a[g] = args.a[0]; args.b[g] = args.b[0];
        args.bias1[g] = nullptr; args.bias2[g] = nullptr; args.addend[g] = nullptr; args.out[g] = nullptr;
    }
    if (g_astat && astat_ok && splitk == 1 && K == AS_NKB * BK && N % AS_BN == 0 && N <= AS_MAX_N && (ldc & 3) == 0) {
        static bool attr_set = false;
        if (!attr_set) {
            CRUSE_CUDA_OK(cudaFuncSetAttribute(gemm_astat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AS_SMEM));
            attr_set = true;
        }
        gemm_astat_tc_kernel<<<dim3((M + BM - 1) / BM, G), IH_THREADS, AS_SMEM, st>>>(args, M, N, ldc, bias2_rows, tm_T);
        CRUSE_LAUNCH_OK();
        return 0;
    }
    if (a_mn && b_mn) return launch_gemm_tile<true, true>(args, G, M, N, K, ldc, bias2_rows, splitk, c_plane, st, tm_T, b_kshift);
    if (a_mn) return launch_gemm_tile<true, false>(args, G, M, N, K, ldc, bias2_rows, splitk, c_plane, st, tm_T, 0);
    if (b_mn) return launch_gemm_tile<false, true>(args, G, M, N, K, ldc, bias2_rows, splitk, c_plane, st, tm_T, b_kshift);
    return launch_gemm_tile<false, false>(args, G, M, N, K, ldc, bias2_rows, splitk, c_plane, st, tm_T, 0);
}

}  // namespace
}  // namespace cruse

using namespace cruse;

static int gru_ih_gemm_tc_impl(const float* x, const float* const* w_ih, const float* const* b_ih, const float* const* b_hh,
                               float* xproj, int M, int G, int H, int tm_T, void* stream) {
    CRUSE_CHECK_ARG(x && xproj && w_ih, "gru_ih_gemm_tc: null pointer");
    CRUSE_CHECK_ARG(M > 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && (H % 4) == 0,
                    "gru_ih_gemm_tc: bad sizes M=%d G=%d H=%d (H%%4==0, G<=%d)", M, G, H, CRUSE_MAX_GROUPS);
    CRUSE_CHECK_ARG(tm_T == 0 || (tm_T > 0 && M % tm_T == 0), "gru_ih_gemm_tm_tc: M=%d is not a multiple of T=%d", M, tm_T);
    GemmArgs args;
    const bool astat = g_astat && H == AS_NKB * BK;          // K = H = 256: the A-stationary kernel (weight boxes of 128 rows)
    for (int g = 0; g < G; ++g) {
        CRUSE_CHECK_ARG(w_ih[g], "gru_ih_gemm_tc: null weight pointer for group %d", g);
        // A: this group's H columns of x (row pitch G*H floats); B: weight_ih_l0 [3H, H]
        if (int rc = make_tmap_2d(&args.a[g], x + (size_t)g * H, (uint64_t)M, (uint64_t)H, (uint64_t)G * H * 4, BM, true)) return rc;
        if (int rc = make_tmap_2d(&args.b[g], w_ih[g], (uint64_t)3 * H, (uint64_t)H, (uint64_t)H * 4, astat ? AS_BN : BN, true)) return rc;
        args.bias1[g] = b_ih ? b_ih[g] : nullptr;
        args.bias2[g] = b_hh ? b_hh[g] : nullptr;
        args.addend[g] = nullptr;
        args.out[g] = xproj + (size_t)g * 3 * H;
    }
    return launch_gemm(args, G, M, 3 * H, H, (long long)G * 3 * H, 2 * H, 1, 0, (cudaStream_t)stream, tm_T, astat);
}

extern "C" int cruse_gemm_set_astat(int on) {
    g_astat = on ? 1 : 0;
    return 0;
}

extern "C" int cruse_gru_ih_gemm_tc(const float* x, const float* const* w_ih, const float* const* b_ih,
                                    const float* const* b_hh, float* xproj, int M, int G, int H, void* stream) {
    return gru_ih_gemm_tc_impl(x, w_ih, b_ih, b_hh, xproj, M, G, H, 0, stream);
}

extern "C" int cruse_gru_ih_gemm_tm_tc(const float* x, const float* const* w_ih, const float* const* b_ih,
                                       const float* const* b_hh, float* xproj, int B, int T, int G, int H, void* stream) {
    CRUSE_CHECK_ARG(B > 0 && T > 0, "gru_ih_gemm_tm_tc: bad sizes B=%d T=%d", B, T);
    return gru_ih_gemm_tc_impl(x, w_ih, b_ih, b_hh, xproj, B * T, G, H, T, stream);
}

extern "C" int cruse_gemm_tn_tc(const float* const* A, const float* const* Bm, const float* const* bias, float* const* C, int G,
                                int M, int N, int K, long long lda, long long ldb, long long ldc, int splitk,
                                long long c_plane, void* stream) {
    CRUSE_CHECK_ARG(A && Bm && C, "gemm_tn_tc: null pointer table");
    CRUSE_CHECK_ARG(G > 0 && G <= CRUSE_MAX_GROUPS && M > 0 && N > 0 && K > 0 && splitk >= 1 && splitk <= 64,
                    "gemm_tn_tc: bad sizes G=%d M=%d N=%d K=%d splitk=%d", G, M, N, K, splitk);
    CRUSE_CHECK_ARG((lda % 4) == 0 && (ldb % 4) == 0 && lda >= K && ldb >= K && ldc >= N, "gemm_tn_tc: bad pitches lda=%lld ldb=%lld ldc=%lld", lda, ldb, ldc);
    CRUSE_CHECK_ARG(splitk == 1 || (bias == nullptr && c_plane >= (long long)M * ldc), "gemm_tn_tc: split-K needs bias == NULL and c_plane >= M*ldc");
    GemmArgs args;
    for (int g = 0; g < G; ++g) {
        CRUSE_CHECK_ARG(A[g] && Bm[g] && C[g], "gemm_tn_tc: null pointer for problem %d", g);
        if (int rc = make_tmap_2d(&args.a[g], A[g], (uint64_t)M, (uint64_t)K, (uint64_t)lda * 4, BM, true)) return rc;
        if (int rc = make_tmap_2d(&args.b[g], Bm[g], (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 4, BN, true)) return rc;
        args.bias1[g] = bias ? bias[g] : nullptr;
        args.bias2[g] = nullptr;
        args.addend[g] = nullptr;
        args.out[g] = C[g];
    }
    return launch_gemm(args, G, M, N, K, ldc, 0, splitk, c_plane, (cudaStream_t)stream);
}

extern "C" int cruse_gemm_tc(const float* const* A, const float* const* Bm, const float* const* bias, const float* const* addend,
                             float* const* C, int G, int M, int N, int K, long long lda, long long ldb, long long ldc, int splitk,
                             long long c_plane, int a_mn, int b_mn, int b_kshift, void* stream) {
    CRUSE_CHECK_ARG(A && Bm && C, "gemm_tc: null pointer table");
    CRUSE_CHECK_ARG(G > 0 && G <= CRUSE_MAX_GROUPS && M > 0 && N > 0 && K > 0 && splitk >= 1 && splitk <= 64,
                    "gemm_tc: bad sizes G=%d M=%d N=%d K=%d splitk=%d", G, M, N, K, splitk);
    CRUSE_CHECK_ARG((lda % 4) == 0 && (ldb % 4) == 0 && lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K) && ldc >= N,
                    "gemm_tc: bad pitches lda=%lld ldb=%lld ldc=%lld", lda, ldb, ldc);
    CRUSE_CHECK_ARG(splitk == 1 || (bias == nullptr && c_plane >= (long long)M * ldc), "gemm_tc: split-K needs bias == NULL and c_plane >= M*ldc");
    CRUSE_CHECK_ARG(b_kshift >= 0 && (b_kshift == 0 || b_mn), "gemm_tc: b_kshift needs an MN-major B operand");
    GemmArgs args;
    for (int g = 0; g < G; ++g) {
        CRUSE_CHECK_ARG(A[g] && Bm[g] && C[g], "gemm_tc: null pointer for problem %d", g);
        // K-major operand: tensor [rows = M|N][cols = K], box {32 k, BM|BN rows}; MN-major: tensor [rows = K][cols = M|N], box {32 features, BK rows}
        if (int rc = a_mn ? make_tmap_2d(&args.a[g], A[g], (uint64_t)K, (uint64_t)M, (uint64_t)lda * 4, BK, true, true)
                          : make_tmap_2d(&args.a[g], A[g], (uint64_t)M, (uint64_t)K, (uint64_t)lda * 4, BM, true)) return rc;
        if (int rc = b_mn ? make_tmap_2d(&args.b[g], Bm[g], (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 4, BK, true, true)
                          : make_tmap_2d(&args.b[g], Bm[g], (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 4, BN, true)) return rc;
        args.bias1[g] = bias ? bias[g] : nullptr;
        args.bias2[g] = nullptr;
        args.addend[g] = addend ? addend[g] : nullptr;
        CRUSE_CHECK_ARG(!args.addend[g] || (reinterpret_cast<uintptr_t>(args.addend[g]) & 15) == 0, "gemm_tc: addend %d is not 16-byte aligned", g);
        args.out[g] = C[g];
    }
    return launch_gemm(args, G, M, N, K, ldc, 0, splitk, c_plane, (cudaStream_t)stream, 0, false, a_mn != 0, b_mn != 0, b_kshift);
}

// ---------------------------------------------------------------------------------------------
// One streaming GRU step for a large batch of concurrent utterances (BASELINE cfg-5: 2048 utterances x 1 frame).
// With T == 1 there is no recurrence to keep on chip: the hidden half is ONE [B,H] x [H,3H] GEMM per group -- the same
// tcgen05 kernel as the input half -- followed by the gate math as an elementwise pass.
//   hproj[g][b][:] = W_hh[g] . h_prev[g][b]          (tf32 operands, fp32 accumulate)
//   r = s(xr + hr), z = s(xz + hz), n = tanh(xn + r * (hn + b_hn)), h' = (1-z) n + z h     (PyTorch gate order r,z,n;
//   xproj already carries b_ih and the r,z part of b_hh: cruse_gru_ih_gemm)
// ---------------------------------------------------------------------------------------------
namespace cruse {
namespace {
struct StepPtrs { const float* b_hh[CRUSE_MAX_GROUPS]; };

__global__ void __launch_bounds__(256)
gru_step_gates_kernel(const float* __restrict__ xproj, const float* __restrict__ hproj, const float* __restrict__ h_prev, const StepPtrs bp,
                      float* __restrict__ h_new, float* __restrict__ y, int B, int G, int H, int y_fs, int y_gs) {
    const int H4 = H >> 2;
    const long long n = (long long)G * B * H4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % H4) * 4;
        const long long gb = i / H4;
        const int b = (int)(gb % B), g = (int)(gb / B);
        const float* xp = xproj + ((size_t)b * G + g) * 3 * H + j;
        const float* hp = hproj + ((size_t)g * B + b) * 3 * H + j;
        const float4 xr = __ldg(reinterpret_cast<const float4*>(xp)), xz = __ldg(reinterpret_cast<const float4*>(xp + H)),
                     xn = __ldg(reinterpret_cast<const float4*>(xp + 2 * H));
        const float4 hr = __ldg(reinterpret_cast<const float4*>(hp)), hz = __ldg(reinterpret_cast<const float4*>(hp + H)),
                     hn = __ldg(reinterpret_cast<const float4*>(hp + 2 * H));
        const float4 ho = h_prev ? __ldg(reinterpret_cast<const float4*>(h_prev + ((size_t)g * B + b) * H + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 bn = bp.b_hh[g] ? __ldg(reinterpret_cast<const float4*>(bp.b_hh[g] + 2 * H + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float o[4];
        const float xr_[4] = {xr.x, xr.y, xr.z, xr.w}, xz_[4] = {xz.x, xz.y, xz.z, xz.w}, xn_[4] = {xn.x, xn.y, xn.z, xn.w};
        const float hr_[4] = {hr.x, hr.y, hr.z, hr.w}, hz_[4] = {hz.x, hz.y, hz.z, hz.w}, hn_[4] = {hn.x, hn.y, hn.z, hn.w};
        const float ho_[4] = {ho.x, ho.y, ho.z, ho.w}, bn_[4] = {bn.x, bn.y, bn.z, bn.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r = sigmoidf_(xr_[q] + hr_[q]);
            const float z = sigmoidf_(xz_[q] + hz_[q]);
            const float nn = tanhf(xn_[q] + r * (hn_[q] + bn_[q]));
            o[q] = (1.f - z) * nn + z * ho_[q];
        }
        *reinterpret_cast<float4*>(h_new + ((size_t)g * B + b) * H + j) = make_float4(o[0], o[1], o[2], o[3]);
        float* yo = y + (size_t)b * G * H + (size_t)g * y_gs;
        if (y_fs == 1) {
            *reinterpret_cast<float4*>(yo + j) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) yo[(size_t)(j + q) * y_fs] = o[q];
        }
    }
}
}  // namespace
}  // namespace cruse

extern "C" size_t cruse_gru_step_ws_bytes(int B, int G, int H) { return sizeof(float) * (size_t)G * B * 3 * H; }

extern "C" int cruse_gru_step(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h_prev,
                              float* h_new, float* y, void* ws, int B, int G, int H, int y_fs, int y_gs, void* stream) {
    CRUSE_CHECK_ARG(xproj && w_hh && h_new && y && ws, "gru_step: null pointer");
    CRUSE_CHECK_ARG(B > 0 && G > 0 && G <= CRUSE_MAX_GROUPS && H > 0 && (H % 4) == 0, "gru_step: bad sizes B=%d G=%d H=%d", B, G, H);
    CRUSE_CHECK_ARG(h_new != h_prev, "gru_step: h_new must not alias h_prev");
    CRUSE_CHECK_ARG((y_gs % 4) == 0 || y_fs != 1, "gru_step: concatenated output needs a 16-byte aligned group stride");
    float* hproj = static_cast<float*>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    if (h_prev) {
        const float* A[CRUSE_MAX_GROUPS];
        float* C[CRUSE_MAX_GROUPS];
        for (int g = 0; g < G; ++g) {
            CRUSE_CHECK_ARG(w_hh[g], "gru_step: null w_hh[%d]", g);
            A[g] = h_prev + (size_t)g * B * H;
            C[g] = hproj + (size_t)g * B * 3 * H;
        }
        if (int rc = cruse_gemm_tn_tc(A, w_hh, nullptr, C, G, B, 3 * H, H, H, H, 3 * H, 1, 0, stream)) return rc;
    } else {
        CRUSE_CUDA_OK(cudaMemsetAsync(hproj, 0, cruse_gru_step_ws_bytes(B, G, H), st));       // zero state: W_hh . 0
    }
    cruse::StepPtrs bp;
    for (int g = 0; g < CRUSE_MAX_GROUPS; ++g) bp.b_hh[g] = (g < G && b_hh) ? b_hh[g] : nullptr;
    const long long n = (long long)G * B * (H / 4);
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)cruse::sm_count() * 16) blocks = (long long)cruse::sm_count() * 16;
    cruse::gru_step_gates_kernel<<<(int)blocks, 256, 0, st>>>(xproj, hproj, h_prev, bp, h_new, y, B, G, H, y_fs, y_gs);
    CRUSE_LAUNCH_OK();
    return 0;
}

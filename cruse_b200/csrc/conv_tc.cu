// conv_tc.cu -- the conv / skip-conv / transposed-conv stages of the U-Net as implicit GEMMs on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in tensor memory).
//
// Replaces, for the 256-bin pyramid in eval mode, the CUDA-core kernels of conv.cu behind the same C ABI
// (cruse_conv_fwd / cruse_convT_fwd): nn.Conv2d((2,3), stride (1,2), pad (1,1)) + "[..., :-1, :]" +
// folded BatchNorm2d + act (model/cruse_net.py:138,141,149-152), the (1,3) skip convs (:143,153-156) and
// nn.ConvTranspose2d((1,3), stride (1,2)) + "[..., :-1]" + BatchNorm2d + act + skip add (:161-163).
//
// GEMM view.  Activations stay frame-major [B, T, C, F].  One M tile = 128 output positions = TF = 128/FO
// consecutive frames x FO bins of one utterance; N = Cout (conv) or (Cout, even/odd output bin) (convT);
// K = (frequency tap, input channel), the two time taps of the causal encoder convs are NOT materialised:
// the A tile holds TF+1 frames and the kt = 1 tap is the same shared-memory tile addressed FO rows further
// down (a descriptor offset, FO*128 bytes, a multiple of the 1024-byte swizzle atom).
//
// Data movement.  HBM -> registers with coalesced 8-byte (stride-2 stages) / 4-byte loads, every input
// element read ONCE per tile; the frequency taps of a row are the lane's own values plus one
// warp shuffle from the neighbouring bin; registers -> shared memory directly in the K-major
// SWIZZLE_128B layout the UMMA descriptors address (lane = 16 rows x 2 channel pairs, so each 8-byte
// store instruction is bank-conflict free), fp32 -> tf32 rounded (cvt.rna) on the way.  Weights are re-laid
// once per CTA into the same layout and stay resident.  No im2col buffer ever exists in HBM.
//
// Pipeline (persistent CTA, one per SM): 2 x 8 producer warps alternate over the ring of A-tile groups
// (all loads of a group are in flight before the ring slot is waited for), one warp issues the MMAs
// (elect.one) and recycles ring slots with tcgen05.commit, four warps run the epilogue out of a
// double-buffered TMEM accumulator (tcgen05.ld -> + bias, folded BN, act, + skip -> global), so loads,
// tensor math and stores of consecutive tiles overlap.
#include "common.cuh"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

namespace cruse {
namespace {

constexpr int CT_NPW = 8;                                  // producer warps per set
// producer sets (NS, per instantiation): 2 sets of 8 warps alternate over the A-tile groups, or 1 set where the deep
// load bursts of two sets delay the epilogue's stores more than they help (measured per stage, see the dispatch table)
constexpr int CT_EPI_WARPS = 8;                            // two per TMEM lane quadrant, splitting the accumulator columns
constexpr int ct_threads(int ns) { return (CT_NPW * ns + 1 + CT_EPI_WARPS) * 32; }   // 800 / 544
constexpr int CT_SMEM_BUDGET = 222 * 1024;

// optional cap on the persistent grid (0 = one CTA per SM): lets a stage run beside the GRU wavefront on the SMs it leaves free
#ifndef CRUSE_CONV_L2_PREFETCH
#define CRUSE_CONV_L2_PREFETCH 1
#endif
int g_conv_max_ctas = 0;

#ifdef CRUSE_CT_TIMING
// developer instrumentation (never built into the shipped library): clock64 stamps of CTA 0, [tile][slot]
__device__ long long g_ct_timing[64 * 16];
#define CT_STAMP(tile_local, slot) do { if (blockIdx.x == 0 && (tile_local) < 64 && (threadIdx.x & 31) == 0) g_ct_timing[(tile_local) * 16 + (slot)] = clock64(); } while (0)
#else
#define CT_STAMP(tile_local, slot) do {} while (0)
#endif

struct ConvTcArgs {
    const float* in;
    const float* w;
    const float* bias;
    const float* scale;
    const float* shift;
    const float* alpha;
    const float* addend;
    float* out;
    int B, T, act;
    int in_tm, out_tm;      // frame records of in / out are ordered time-major (t*B + b) instead of (b*T + t)
    const float* hist;      // MODE 0, KT == 2: frame -1 of every utterance [B][CIN][FIN] (streaming: the previous chunk's last
                            //    input frame) or NULL = zero padding
    int t_begin, t_end;     // output frames [t_begin, t_end) only (0, 0 = all T): the net is causal, so a time range of a stage needs
                            //    nothing but the same range (and one frame before it) of the stage below -- lets the head of the
                            //    encoder / the tail of the decoder run beside the recurrence instead of before / after it
    int wmode;              // 0: w is this conv's weight; 1 (KT == 1 conv only): data gradient of a (1,3)/stride-1 conv --
                            //    w is THAT conv's weight [Cin_here][Cout_here][1][3], taps flipped
    const float* w2;        // FUSE: weight [CIN][CIN][1][3] of the (1,3) skip conv of this stage's INPUT (model/cruse_net.py:143,153-155)
    float* out2;            // FUSE: its output, frame-major [B,T,CIN,FIN]
};

// MODE 0: conv KT x 3, frequency stride SF, pad 1 (FO = output bins).  MODE 1: convT 1 x 3, stride 2, cropped
// to 2*FO bins (FO = input bins).  MODE 3: conv 1 x 3, stride 2, NO left pad (taps at bins 2fo, 2fo+1, 2fo+2, the last one
// cropped at the right edge) = the data gradient of the transposed conv.  MODE 2: the data gradient of the KT x 3 / stride-2
// conv: a transposed conv over dz (FO = dz bins, output 2*FO bins) whose frequency taps are {dz[i], dz[i+1]} and whose time taps
// look FORWARD (frame t and t+1); CIN = channels of dz (the conv's Cout), COUT = channels of the gradient (the conv's Cin).
// FUSE (MODE 0, KT x 3 / stride 2 only): the (1,3)/stride-1 skip conv of the stage's INPUT rides along.  The stage's A tile already
// holds the input bins 2fo-1, 2fo, 2fo+1 of every row; with a fourth tap (bin 2fo+2) the skip outputs of the bins 2fo and 2fo+1
// are two more groups of accumulator columns (N = COUT + 2*CIN) whose weights are zero outside the current frame's taps -- the
// input tensor is read once instead of twice and the skip conv costs no launch of its own.
template <int MODE, int KT, int SF, int CIN, int COUT, int FO, int GM, int FUSE = 0>
struct ConvTcCfg {
    static constexpr bool CONVLIKE = MODE == 0 || MODE == 3;
    static constexpr int TAPS = CONVLIKE ? (FUSE ? 4 : 3) : 2;     // conv: kf = 0,1,2 (+ bin 2fo+2 when fused); convT: {x[i], x[i-1]}
    static constexpr int N = CONVLIKE ? COUT + (FUSE ? 2 * CIN : 0) : 2 * COUT;
    static constexpr int NPAD = N < 16 ? 16 : N;                   // UMMA M=128 needs N % 16 == 0
    static constexpr int TF = GM * (128 / FO);                     // frames per tile = GM MMA tiles of 128 rows: short-frame
                                                                   // stages (FO >= 32) batch several so that the per-tile
                                                                   // barrier round trips are paid once per ~8 frames
    static constexpr int NFR = TF + KT - 1;                        // frames held by one A slot
    static constexpr int SR = NFR * FO;                            // rows per slot
    static constexpr int SLOT_BYTES = SR * 128;
    static constexpr int CB = CIN < 32 ? CIN : 32;                 // input channels per group
    static constexpr int NG = CIN / CB;                            // groups per tile
    static constexpr int KG = TAPS * CB;                           // K per group
    static constexpr int S = (KG + 31) / 32;                       // slots (32-float K blocks) per group
    static constexpr int GROUP_BYTES = S * SLOT_BYTES;
    static constexpr int NKB = NG * S * KT;                        // weight K blocks
    static constexpr int B_BYTES = NKB * NPAD * 128;
    static constexpr int FIN = CONVLIKE ? SF * FO : FO;
    static constexpr int RD_FIT = (CT_SMEM_BUDGET - B_BYTES) / GROUP_BYTES;
    static constexpr int RD = RD_FIT > 4 ? 4 : RD_FIT;             // ring depth (groups)
    static constexpr int ITEMS = NFR * (CB / 4) * (FO / 16);       // (frame, 4 channels, 16 rows) patches per group
    static constexpr int NIT = ITEMS / CT_NPW;
    static constexpr int ACC_COLS = GM * NPAD;                     // one accumulator buffer: GM tiles side by side
    static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : (2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : 256));
    static constexpr int SMEM = 1024 + RD * GROUP_BYTES + B_BYTES + COUT * 16 + 256;
    static_assert(128 % FO == 0 && FO % 16 == 0, "FO must divide 128 and be a multiple of 16");
    static_assert(CIN % 4 == 0 && CIN % CB == 0, "channel blocking");
    static_assert(ITEMS % CT_NPW == 0, "producer items must split evenly over the warps of a set");
    static_assert(RD >= 2, "ring needs two groups");
    static_assert(NPAD % 16 == 0 && NPAD <= 64 && 2 * ACC_COLS <= 256, "N tile");
    static_assert(NIT <= (MODE == 2 ? 10 : 9), "producer register budget: at most 9 patches per warp and group (10 with one value per patch)");
    static_assert(MODE == 0 || (MODE == 1 && KT == 1 && SF == 1) || (MODE == 3 && KT == 1 && SF == 2) || (MODE == 2 && SF == 1), "instantiation");
    static_assert(!FUSE || (MODE == 0 && SF == 2 && COUT % 16 == 0), "the fused skip conv rides on the stride-2 encoder stages");
};

__device__ __forceinline__ uint32_t f32_to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {   // this warp's 32 lanes x 16 columns
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int MODE, int KT, int SF, int CIN, int COUT, int FO, int GM, int NS, int FUSE = 0>
__global__ void __launch_bounds__(ct_threads(NS), 1) conv_tc_kernel(const ConvTcArgs a) {
    constexpr int CT_SETS = NS, CT_PROD_WARPS = CT_NPW * NS, CT_MMA_WARP = CT_PROD_WARPS, CT_EPI_WARP0 = CT_PROD_WARPS + 1;
    constexpr int CT_THREADS = ct_threads(NS);
    using C = ConvTcCfg<MODE, KT, SF, CIN, COUT, FO, GM, FUSE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;          // swizzle atoms are 1024-byte aligned
    uint8_t* ring = smem_raw + (base - tc::smem_u32(smem_raw));
    uint8_t* sB = ring + C::RD * C::GROUP_BYTES;
    float4* s_par = reinterpret_cast<float4*>(sB + C::B_BYTES);                // per channel: scale, bias*scale + shift, alpha
    uint64_t* full = reinterpret_cast<uint64_t*>(s_par + COUT);
    uint64_t* empty = full + 4;
    uint64_t* acc_full = empty + 4;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) CT_STAMP(63, 14);                               // kernel start
    // programmatic dependent launch: the next kernel of the stream may start its own set-up (weights, TMEM, barriers) on SMs this
    // grid has left; everything that touches activations sits behind griddepcontrol.wait below
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int T = a.T;
    const int chunks = (a.t_end - a.t_begin + C::TF - 1) / C::TF;
    const int ntiles = a.B * chunks;

    // ---- one-time setup: barriers, TMEM, weights in UMMA layout, epilogue parameters
    if (tid == 0) {
        for (int i = 0; i < C::RD; ++i) { tc::mbar_init(&full[i], CT_NPW * 32); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], CT_EPI_WARPS * 32); }
        tc::fence_barrier_init();
    }
    if (warp == CT_MMA_WARP) tc::tmem_alloc<C::TMEM_COLS>(tmem_slot);
    {
        constexpr int WTOT = C::NKB * C::NPAD * 32, WU = 8;
        for (int i0 = tid; i0 < WTOT; i0 += CT_THREADS * WU) {
            float wv[WU];
#pragma unroll
            for (int u = 0; u < WU; ++u) {            // WU independent loads in flight per thread
                const int i = i0 + u * CT_THREADS;
                const int kk = i & 31, n = (i >> 5) % C::NPAD, kbi = (i >> 5) / C::NPAD;
                const int kt = kbi % KT, gs = kbi / KT, s = gs % C::S, g = gs / C::S;
                const int k = s * 32 + kk;                       // K index inside the group: tap * CB + channel
                float v = 0.f;
                if (i < WTOT && k < C::KG && n < C::N) {
                    const int tap = k / C::CB, ci = g * C::CB + (k - tap * C::CB);
                    if (C::CONVLIKE && FUSE && n >= COUT) {
                        // fused skip conv (Conv2d [CIN][CIN][1][3], pad 1): column COUT + 2*cs + ph = output bin 2fo+ph of channel cs,
                        // which reads the CURRENT frame's taps ph, ph+1, ph+2 (tap t = input bin 2fo-1+t)
                        const int cs = (n - COUT) >> 1, ph = (n - COUT) & 1, kf = tap - ph;
                        v = (kt == KT - 1 && kf >= 0 && kf < 3) ? __ldg(a.w2 + ((size_t)cs * CIN + ci) * 3 + kf) : 0.f;
                    } else if (C::CONVLIKE && FUSE && tap == 3) {
                        v = 0.f;                                                                     // the stage itself has three taps
                    } else if (C::CONVLIKE) {
                        v = a.wmode == 1 ? __ldg(a.w + ((size_t)ci * COUT + n) * 3 + (2 - tap))        // dgrad: [Cin][Cout][1][3], flipped
                                         : __ldg(a.w + (((size_t)n * CIN + ci) * KT + kt) * 3 + tap);  // Conv2d [Cout][Cin][KT][3]
                    } else if (MODE == 2) {
                        const int co = n >> 1, ph = n & 1;                                            // Conv2d [CIN][COUT][KT][3] (out, in)
                        // dx[2i] = W[..,1] dz[i];  dx[2i+1] = W[..,2] dz[i] + W[..,0] dz[i+1]
                        const float* wp = a.w + (((size_t)ci * COUT + co) * KT + kt) * 3;
                        if (tap == 0) v = __ldg(wp + 1 + ph);
                        else if (ph == 1) v = __ldg(wp);
                    } else {
                        const int co = n >> 1, ph = n & 1;                                            // ConvTranspose2d [Cin][Cout][1][3]
                        // out[2i] = W[..,0] x[i] + W[..,2] x[i-1];  out[2i+1] = W[..,1] x[i]
                        if (tap == 0) v = __ldg(a.w + ((size_t)ci * COUT + co) * 3 + ph);
                        else if (ph == 0) v = __ldg(a.w + ((size_t)ci * COUT + co) * 3 + 2);
                    }
                }
                wv[u] = v;
            }
#pragma unroll
            for (int u = 0; u < WU; ++u) {
                const int i = i0 + u * CT_THREADS;
                if (i < WTOT) {
                    const int kk = i & 31, n = (i >> 5) % C::NPAD, kbi = (i >> 5) / C::NPAD;
                    const uint32_t off = (uint32_t)kbi * (C::NPAD * 128) + (uint32_t)(n >> 3) * 1024 + (uint32_t)(n & 7) * 128 +
                                         (uint32_t)(((kk >> 2) ^ (n & 7)) << 4) + (uint32_t)(kk & 3) * 4;
                    *reinterpret_cast<uint32_t*>(sB + off) = f32_to_tf32(wv[u]);
                }
            }
        }
    }
    // K padding inside a slot (CIN = 8: 24 of 32) is never written by the producers: zero the ring once
    if (C::KG % 32 != 0)
        for (int i = tid; i < C::RD * C::GROUP_BYTES / 16; i += CT_THREADS) reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < COUT; i += CT_THREADS) {
        const float bi = a.bias ? __ldg(a.bias + i) : 0.f, sc = a.scale ? __ldg(a.scale + i) : 1.f, sh = a.shift ? __ldg(a.shift + i) : 0.f;
        s_par[i] = make_float4(sc, fmaf(bi, sc, sh), a.alpha ? __ldg(a.alpha + i) : 0.f, 0.f);
    }
    tc::fence_proxy_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;
    // the set-up above read parameters only; the upstream grid (which produced this stage's input / skip tensors) must have
    // completed and flushed before any activation is read or written (no-op when launched without the PDL attribute)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (warp == 0) CT_STAMP(63, 13);                               // setup done

    if (warp < CT_PROD_WARPS) {
        // ================= producers: HBM -> registers -> swizzled K-major A tiles =================
        const int set = warp / CT_NPW, wq = warp % CT_NPW;
        // lane -> (row of a 16-row patch, channel pair of a 4-channel chunk): one 8-byte load instruction covers
        // 2 channels x 16 bins (two full 128-byte lines when SF == 2), one 8-byte store instruction is conflict free
        // (per half warp: 8 row phases x 2 channel pairs = 16 distinct 8-byte bank pairs of the swizzled tile)
        const int row16 = lane >> 1, ep = lane & 1;
        // Software pipeline over this set's groups (j = set, set+2, ...): the registers of a group are split in two halves;
        // while half 0 of group j is being converted and stored, the loads of half 0 of the set's NEXT group are already
        // in flight (and likewise for half 1), so HBM requests are outstanding during the whole store phase.
        constexpr int NV = (C::CONVLIKE && SF == 2) ? 2 : 1;
        constexpr int NH0 = (C::NIT + 1) / 2;
        float va[C::NIT][NV], vb[C::NIT][NV], ea[C::NIT], eb[C::NIT];     // channel ci0 / ci0+1: values, edge value
        const int ntl = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA
        const int ngroups = ntl * C::NG;

        auto load_items = [&](int lo, int hi, int j) {
            const int tile = blockIdx.x + (j / C::NG) * gridDim.x, g = j % C::NG;
            const int b = tile / chunks, t0 = a.t_begin + (tile - b * chunks) * C::TF;
#pragma unroll
            for (int it = 0; it < C::NIT; ++it) {
                if (it < lo || it >= hi) continue;
                const int item = it * CT_NPW + wq;
                const int fg = item % (FO / 16), cg = (item / (FO / 16)) % (C::CB / 4), fr = item / ((FO / 16) * (C::CB / 4));
                const int t = MODE == 2 ? t0 + fr : t0 - (KT - 1) + fr;
                bool valid = (t >= 0) && (t < T);
                const int fo = fg * 16 + row16, ci0 = g * C::CB + cg * 4 + 2 * ep;
                const float* src = a.in + ((a.in_tm ? (size_t)t * a.B + b : (size_t)b * T + t) * CIN + ci0) * C::FIN + (C::CONVLIKE ? SF : 1) * fo;
                if (MODE == 0 && KT == 2 && t == -1 && a.hist) {      // streaming: the time tap of frame 0 reads the carried frame
                    valid = true;
                    src = a.hist + ((size_t)b * CIN + ci0) * C::FIN + SF * fo;
                }
#pragma unroll
                for (int q = 0; q < NV; ++q) { va[it][q] = 0.f; vb[it][q] = 0.f; }
                ea[it] = 0.f; eb[it] = 0.f;
                if (valid) {
                    if (C::CONVLIKE && SF == 2) {
                        const float2 p = __ldg(reinterpret_cast<const float2*>(src));
                        const float2 q = __ldg(reinterpret_cast<const float2*>(src + C::FIN));
                        va[it][0] = p.x; va[it][NV - 1] = p.y; vb[it][0] = q.x; vb[it][NV - 1] = q.y;
                    } else {
                        va[it][0] = __ldg(src);
                        vb[it][0] = __ldg(src + C::FIN);
                    }
                    if ((MODE == 0 || MODE == 1) && row16 == 0 && fo > 0) { ea[it] = __ldg(src - 1); eb[it] = __ldg(src + C::FIN - 1); }
                    if ((MODE == 3 || FUSE) && row16 == 15 && fo < FO - 1) { ea[it] = __ldg(src + 2); eb[it] = __ldg(src + C::FIN + 2); }
                    if (((MODE == 0 && SF == 1) || MODE == 2) && row16 == 15 && fo < FO - 1) { ea[it] = __ldg(src + 1); eb[it] = __ldg(src + C::FIN + 1); }
                }
            }
        };
        auto store_items = [&](int lo, int hi, uint8_t* grp) {
#pragma unroll
            for (int it = 0; it < C::NIT; ++it) {
                if (it < lo || it >= hi) continue;
                const int item = it * CT_NPW + wq;
                const int fg = item % (FO / 16), cg = (item / (FO / 16)) % (C::CB / 4), fr = item / ((FO / 16) * (C::CB / 4));
                const int row = fr * FO + fg * 16 + row16;
                float ta[4], tb[4];
                ta[3] = tb[3] = 0.f;
                if (MODE == 3) {                     // taps read bins 2fo, 2fo+1, 2fo+2 (cropped)
                    float ra = __shfl_down_sync(0xffffffffu, va[it][0], 2), rb = __shfl_down_sync(0xffffffffu, vb[it][0], 2);
                    if (row16 == 15) { ra = ea[it]; rb = eb[it]; }
                    ta[0] = va[it][0]; ta[1] = va[it][NV - 1]; ta[2] = ra;
                    tb[0] = vb[it][0]; tb[1] = vb[it][NV - 1]; tb[2] = rb;
                } else if (MODE == 0 && SF == 2) {   // taps read bins 2fo-1, 2fo, 2fo+1
                    float la = __shfl_up_sync(0xffffffffu, va[it][NV - 1], 2), lb = __shfl_up_sync(0xffffffffu, vb[it][NV - 1], 2);
                    if (row16 == 0) { la = ea[it]; lb = eb[it]; }
                    ta[0] = la; ta[1] = va[it][0]; ta[2] = va[it][NV - 1];
                    tb[0] = lb; tb[1] = vb[it][0]; tb[2] = vb[it][NV - 1];
                    if (FUSE) {                      // fourth tap: bin 2fo+2 = the next row's own first value (lane 15: the edge load)
                        float ra = __shfl_down_sync(0xffffffffu, va[it][0], 2), rb = __shfl_down_sync(0xffffffffu, vb[it][0], 2);
                        if (row16 == 15) { ra = ea[it]; rb = eb[it]; }
                        ta[3] = ra; tb[3] = rb;
                    }
                } else if (MODE == 0) {              // taps read bins fo-1, fo, fo+1
                    float la = __shfl_up_sync(0xffffffffu, va[it][0], 2), lb = __shfl_up_sync(0xffffffffu, vb[it][0], 2);
                    float ra = __shfl_down_sync(0xffffffffu, va[it][0], 2), rb = __shfl_down_sync(0xffffffffu, vb[it][0], 2);
                    if (row16 == 0) { la = ea[it]; lb = eb[it]; }
                    if (row16 == 15) { ra = ea[it]; rb = eb[it]; }
                    ta[0] = la; ta[1] = va[it][0]; ta[2] = ra;
                    tb[0] = lb; tb[1] = vb[it][0]; tb[2] = rb;
                } else if (MODE == 2) {              // conv dgrad: K taps = dz[i], dz[i+1]
                    float ra = __shfl_down_sync(0xffffffffu, va[it][0], 2), rb = __shfl_down_sync(0xffffffffu, vb[it][0], 2);
                    if (row16 == 15) { ra = ea[it]; rb = eb[it]; }
                    ta[0] = va[it][0]; ta[1] = ra; ta[2] = 0.f;
                    tb[0] = vb[it][0]; tb[1] = rb; tb[2] = 0.f;
                } else {                             // convT: K taps = x[i], x[i-1]
                    float la = __shfl_up_sync(0xffffffffu, va[it][0], 2), lb = __shfl_up_sync(0xffffffffu, vb[it][0], 2);
                    if (row16 == 0) { la = ea[it]; lb = eb[it]; }
                    ta[0] = va[it][0]; ta[1] = la; ta[2] = 0.f;
                    tb[0] = vb[it][0]; tb[1] = lb; tb[2] = 0.f;
                }
                uint8_t* rowp = grp + (row >> 3) * 1024 + (row16 & 7) * 128 + ep * 8;     // row & 7 == row16 & 7
#pragma unroll
                for (int tap = 0; tap < C::TAPS; ++tap) {
                    const int k = tap * C::CB + cg * 4;                                   // + 2*ep (+1)
                    const int slot = k >> 5, ch = (k & 31) >> 2;
                    *reinterpret_cast<uint2*>(rowp + slot * C::SLOT_BYTES + ((ch ^ (row16 & 7)) << 4)) =
                        make_uint2(f32_to_tf32(ta[tap]), f32_to_tf32(tb[tap]));
                }
            }
        };

        // L2 prefetch of the input block of this set's group after next (one frame record slice per lane, bulk prefetch): the register
        // pipeline above keeps ONE group of loads in flight per set -- 8-16 KB per SM, a quarter of what the HBM latency needs (ncu r2b:
        // dram 25 %, long-scoreboard stalls 45 %) -- so the data of the groups behind it is pulled into L2 ahead of the loads.
        // Measured (32 x 501 frames, alone): 16->32 + skip 67.6 -> 65.5 us, 32->64 49.2 -> 47.1 us, 8->16 + skip unchanged (71.7):
        // the load latency is NOT what bounds these stages (the stalls are the pipeline's mbarrier waits); kept for the 2 us.
        auto prefetch_group = [&](int j) {
#if CRUSE_CONV_L2_PREFETCH
            if (wq != 0 || lane >= C::NFR || j >= ngroups) return;
            const int tile = blockIdx.x + (j / C::NG) * gridDim.x, g = j % C::NG;
            const int b = tile / chunks, t0 = a.t_begin + (tile - b * chunks) * C::TF;
            const int t = (MODE == 2 ? t0 : t0 - (KT - 1)) + lane;
            if (t < 0 || t >= T) return;
            const float* src = a.in + ((a.in_tm ? (size_t)t * a.B + b : (size_t)b * T + t) * CIN + g * C::CB) * C::FIN;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(C::CB * C::FIN * 4)) : "memory");
#endif
        };
        if (set < ngroups) load_items(0, C::NIT, set);
        prefetch_group(set + CT_SETS);
#pragma unroll 1
        for (int j = set; j < ngroups; j += CT_SETS) {
            const int r = j % C::RD;
            const bool more = j + CT_SETS < ngroups;
            prefetch_group(j + 2 * CT_SETS);
            if (wq == 0) CT_STAMP(j, 0 + set * 3);
            tc::mbar_wait_backoff(&empty[r], ((j / C::RD) & 1) ^ 1);
            if (wq == 0) CT_STAMP(j, 1 + set * 3);            // ring slot free
            uint8_t* grp = ring + r * C::GROUP_BYTES;
            store_items(0, NH0, grp);
            if (more) load_items(0, NH0, j + CT_SETS);
            store_items(NH0, C::NIT, grp);
            if (more) load_items(NH0, C::NIT, j + CT_SETS);
            tc::fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core's async-proxy reads
            tc::mbar_arrive(&full[r]);
            if (wq == 0) CT_STAMP(j, 2 + set * 3);            // stored + arrived
        }
    } else if (warp == CT_MMA_WARP) {
        // ================= MMA issuer =================
        constexpr uint32_t idesc = tc::instr_desc(2 /*tf32*/, 128, C::NPAD);
        const uint32_t sB_u32 = base + C::RD * C::GROUP_BYTES;
        int j = 0, lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int ab = lt & 1;
            CT_STAMP(lt, 6);
            tc::mbar_wait(&acc_empty[ab], ((lt >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            CT_STAMP(lt, 7);                                       // accumulator buffer free
            const uint32_t d0 = tmem_d + (uint32_t)(ab * C::ACC_COLS);
#pragma unroll 1
            for (int g = 0; g < C::NG; ++g, ++j) {
                const int r = j % C::RD;
                tc::mbar_wait(&full[r], (j / C::RD) & 1);
                tc::tc_fence_after();
                if (g == C::NG - 1) CT_STAMP(lt, 8);               // last A group of the tile landed
                if (tc::elect_one()) {
                    const uint32_t ga = base + r * C::GROUP_BYTES;
#pragma unroll
                    for (int mt = 0; mt < GM; ++mt) {                // GM MMA tiles of 128 rows share the slot group
#pragma unroll
                        for (int s = 0; s < C::S; ++s) {
                            const int kvalid = (C::KG - 32 * s) < 32 ? (C::KG - 32 * s) : 32;
                            const int ksteps = (kvalid + 7) / 8;
#pragma unroll
                            for (int kt = 0; kt < KT; ++kt) {
                                const uint32_t sa = ga + s * C::SLOT_BYTES + (mt * 128 + (MODE == 2 ? KT - 1 - kt : kt) * FO) * 128;
                                const uint32_t sb = sB_u32 + (uint32_t)(((g * C::S + s) * KT + kt) * (C::NPAD * 128));
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (k < ksteps)
                                        tc::umma_tf32(d0 + mt * C::NPAD, tc::smem_desc_sw128(sa + k * 32), tc::smem_desc_sw128(sb + k * 32), idesc,
                                                      (g | s | kt | k) ? 1u : 0u);
                                }
                            }
                        }
                    }
                    tc::umma_commit(&empty[r]);                     // ring slot free once these MMAs have read it
                    if (g == C::NG - 1) tc::umma_commit(&acc_full[ab]);
                }
                __syncwarp();
                if (g == C::NG - 1) CT_STAMP(lt, 9);               // MMAs issued
            }
        }
    } else {
        // ================= epilogue: TMEM -> registers -> bias, folded BN, act, + skip -> global =================
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int chalf = (warp - CT_EPI_WARP0) >> 2;                // which half of the accumulator columns
        // work units = (MMA tile mt, 16-column chunk): the two warps of a quadrant take alternate units
        constexpr int NCH = C::NPAD / 16;
        constexpr int NU = GM * NCH;
        constexpr int MYU = NU >= 2 ? NU / 2 : 1;
        const bool have = NU >= 2 || chalf == 0;                     // warp-uniform
        const int act = a.act;
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int b = tile / chunks, t0 = a.t_begin + (tile - b * chunks) * C::TF;
            const int ab = lt & 1;
            if (warp == CT_EPI_WARP0) CT_STAMP(lt, 10);
            tc::mbar_wait(&acc_full[ab], (lt >> 1) & 1);
            tc::tc_fence_after();
            if (warp == CT_EPI_WARP0) CT_STAMP(lt, 11);           // accumulator complete
            // pull this warp's share of the accumulator into registers and hand the TMEM buffer straight back to the MMA
            // warp (relaxed arrive: it must not wait for this or the previous tile's global stores to be performed)
            float v[MYU][16];
            if (have) {
#pragma unroll
                for (int i = 0; i < MYU; ++i) {
                    const int u = NU >= 2 ? chalf + 2 * i : 0, mt = u / NCH, ch = u % NCH;
                    tmem_ld16(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * C::ACC_COLS + mt * C::NPAD + ch * 16), v[i]);
                }
                tc::tmem_ld_wait();
            }
            tc::tc_fence_before();
            tc::mbar_arrive_relaxed(&acc_empty[ab]);
            if (have) {
#pragma unroll
                for (int i = 0; i < MYU; ++i) {
                    const int u = NU >= 2 ? chalf + 2 * i : 0, mt = u / NCH, c0 = (u % NCH) * 16;
                    const int row = mt * 128 + quad * 32 + lane;
                    const int tl = row / FO, fo = row % FO;
                    const int t = t0 + tl;
                    if (t >= a.t_end) continue;
                    const size_t rec = a.out_tm ? (size_t)t * a.B + b : (size_t)b * T + t;
                    if (C::CONVLIKE && FUSE && c0 >= COUT) {
                        // the fused skip conv: columns (cs, ph) -> out2[b, t, cs, 2fo + ph], frame-major, no BatchNorm / activation
                        const size_t o2 = ((size_t)b * T + t) * CIN * (2 * FO) + 2 * fo;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int cs = ((c0 - COUT) >> 1) + q;
                            if (cs < CIN) *reinterpret_cast<float2*>(a.out2 + o2 + (size_t)cs * (2 * FO)) = make_float2(v[i][2 * q], v[i][2 * q + 1]);
                        }
                    } else if (C::CONVLIKE) {
                        const size_t o0 = rec * COUT * FO + fo;
                        float ad[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q)      // all skip loads of the chunk in flight before the first store
                            ad[q] = (a.addend && c0 + q < COUT) ? __ldg(a.addend + o0 + (size_t)(c0 + q) * FO) : 0.f;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const int co = c0 + q;
                            if (co < COUT) {
                                const float4 pr = s_par[co];
                                a.out[o0 + (size_t)co * FO] = apply_act(fmaf(v[i][q], pr.x, pr.y), act, pr.z) + ad[q];
                            }
                        }
                    } else {
                        const size_t o0 = rec * COUT * (2 * FO) + 2 * fo;
                        float2 ad[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            ad[q] = (a.addend && (c0 >> 1) + q < COUT) ? __ldg(reinterpret_cast<const float2*>(a.addend + o0 + (size_t)((c0 >> 1) + q) * (2 * FO)))
                                                                      : make_float2(0.f, 0.f);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int co = (c0 >> 1) + q;
                            if (co < COUT) {
                                const float4 pr = s_par[co];
                                *reinterpret_cast<float2*>(a.out + o0 + (size_t)co * (2 * FO)) =
                                    make_float2(apply_act(fmaf(v[i][2 * q], pr.x, pr.y), act, pr.z) + ad[q].x,
                                                apply_act(fmaf(v[i][2 * q + 1], pr.x, pr.y), act, pr.z) + ad[q].y);
                            }
                        }
                    }
                }
            }
            if (warp == CT_EPI_WARP0) CT_STAMP(lt, 12);           // tile stored
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == CT_MMA_WARP) tc::tmem_dealloc<C::TMEM_COLS>(tmem_d);
}

// programmatic dependent launch of the conv stages: OFF by default (CRUSE_CONV_PDL=1 turns it on).  Measured on B200 with the set-up
// of kernel N+1 overlapping the tail of kernel N: inference 1.565 vs 1.48 ms per step, training 6.29 vs 6.07 ms -- the early-resident
// dependent CTAs (200 KB of shared memory each) sit in griddepcontrol.wait on SMs that the other streams of the step would have used.
inline bool conv_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CRUSE_CONV_PDL");
        v = (e && strcmp(e, "1") == 0) ? 1 : 0;
    }
    return v == 1;
}

template <int MODE, int KT, int SF, int CIN, int COUT, int FO, int GM, int NS, int FUSE = 0>
int launch_conv_tc(const ConvTcArgs& a_in, cudaStream_t st) {
    using C = ConvTcCfg<MODE, KT, SF, CIN, COUT, FO, GM, FUSE>;
    ConvTcArgs a = a_in;
    if (a.t_end <= 0) { a.t_begin = 0; a.t_end = a.T; }
    if (a.t_begin < 0 || a.t_begin >= a.t_end || a.t_end > a.T) return -1;
    auto kern = conv_tc_kernel<MODE, KT, SF, CIN, COUT, FO, GM, NS, FUSE>;
    static bool attr_set = false;                                   // per instantiation; benign if raced
    if (!attr_set) {
        CRUSE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_set = true;
    }
    const int chunks = (a.t_end - a.t_begin + C::TF - 1) / C::TF;
    const long long ntiles = (long long)a.B * chunks;
    int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    if (g_conv_max_ctas > 0 && grid > g_conv_max_ctas) grid = g_conv_max_ctas;
    if (conv_pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(ct_threads(NS));
        cfg.dynamicSmemBytes = C::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CRUSE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a));
        return 0;
    }
    kern<<<grid, ct_threads(NS), C::SMEM, st>>>(a);
    CRUSE_LAUNCH_OK();
    return 0;
}

// process-wide numeric mode of the conv stages: 1 = tf32 tensor cores (default), 0 = exact-fp32 CUDA cores.
// Initialised from the environment (CRUSE_CONV=fp32|tf32), switchable through cruse_conv_set_mode().
int g_conv_mode = -1;
int conv_mode() {
    if (g_conv_mode < 0) {
        const char* e = getenv("CRUSE_CONV");
        g_conv_mode = (e && strcmp(e, "fp32") == 0) ? 0 : 1;
    }
    return g_conv_mode;
}
bool conv_tc_enabled() { return conv_mode() == 1; }


}  // namespace

int conv_max_ctas() { return g_conv_max_ctas; }

// Returns 1 when the stage was launched on the tensor cores, 0 when no instantiation matches (the caller then
// runs the CUDA-core kernel), < 0 on error.
int conv_tc_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                int act, const float* addend, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride,
                int in_tm, int out_tm, int wmode, cudaStream_t st, const float* hist, int t_begin, int t_end) {
    if (!conv_tc_enabled()) return 0;
    if (hist && (kt != 2 || in_tm || (reinterpret_cast<uintptr_t>(hist) & 15))) return 0;
    if (wmode != 0 && !(wmode == 1 && kt == 1 && fstride == 1)) return 0;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(addend) & 15)) return 0;
    ConvTcArgs a{in, w, bias, scale, shift, alpha, addend, out, B, T, act, in_tm, out_tm, hist, t_begin, t_end, wmode, nullptr, nullptr};
    int rc = 0;
    // last argument: MMA tiles (128 rows) per pipeline step; stages with long frames (FO >= 32) batch several of them
#define CRUSE_CT_CONV(KT_, SF_, CI_, CO_, FO_, GM_, NS_)                                                    \
    if (kt == KT_ && fstride == SF_ && Cin == CI_ && Cout == CO_ && Fout == FO_ && Fin == SF_ * FO_) {      \
        rc = launch_conv_tc<0, KT_, SF_, CI_, CO_, FO_, GM_, NS_>(a, st);                                   \
        return rc ? rc : 1;                                                                                 \
    }
    CRUSE_CT_CONV(2, 2, 8, 16, 64, 4, 2)
    CRUSE_CT_CONV(2, 2, 16, 32, 32, 1, 2)
    CRUSE_CT_CONV(2, 2, 32, 64, 16, 1, 1)
    CRUSE_CT_CONV(1, 1, 8, 8, 128, 4, 2)
    CRUSE_CT_CONV(1, 1, 16, 16, 64, 1, 2)
    CRUSE_CT_CONV(1, 1, 32, 32, 32, 1, 2)
    CRUSE_CT_CONV(1, 1, 64, 64, 16, 1, 2)
#undef CRUSE_CT_CONV
    return 0;
}

// encoder stage (2,3)/stride-(1,2) + folded BN + act WITH the (1,3) skip conv of its input fused in (eval mode, whole utterances or a
// frame range): out = act(BN(conv(in))), out2 = conv1x3(in; w2).  Returns 1 when launched, 0 when no instantiation matches.
int conv_skip_tc_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                     int act, const float* w2, float* out, float* out2, int B, int T, int Cin, int Fin, int Cout, int Fout, int out_tm,
                     cudaStream_t st, int t_begin, int t_end) {
    if (!conv_tc_enabled()) return 0;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(out2) & 15)) return 0;
    ConvTcArgs a{in, w, bias, scale, shift, alpha, nullptr, out, B, T, act, 0, out_tm, nullptr, t_begin, t_end, 0, w2, out2};
    int rc = 0;
#define CRUSE_CT_FUSED(CI_, CO_, FO_, GM_, NS_)                                      \
    if (Cin == CI_ && Cout == CO_ && Fout == FO_ && Fin == 2 * FO_) {                \
        rc = launch_conv_tc<0, 2, 2, CI_, CO_, FO_, GM_, NS_, 1>(a, st);             \
        return rc ? rc : 1;                                                          \
    }
    CRUSE_CT_FUSED(8, 16, 64, 2, 2)
    CRUSE_CT_FUSED(16, 32, 32, 1, 2)
#undef CRUSE_CT_FUSED
    return 0;
}

// data gradient of the transposed conv: din[ci, i] = sum_co sum_k W[ci,co,0,k] dz[co, 2i+k] -- a (1,3)/stride-2 conv over dz
// without left pad; w is the ConvTranspose2d weight [Cin][Cout][1][3], which is exactly the [out][in][1][3] this conv reads
int convT_dgrad_tc_try(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin, int Fin, int Cout,
                       int Fout, cudaStream_t st) {
    if (!conv_tc_enabled()) return 0;
    if ((reinterpret_cast<uintptr_t>(dz) & 15) || (reinterpret_cast<uintptr_t>(din) & 15) || (reinterpret_cast<uintptr_t>(addend) & 15)) return 0;
    ConvTcArgs a{dz, w, nullptr, nullptr, nullptr, nullptr, addend, din, B, T, CRUSE_ACT_NONE, 0, 0, nullptr, 0, 0, 0, nullptr, nullptr};
    int rc = 0;
#define CRUSE_CT_TD(CO_, CI_, FI_, NS_)                                              \
    if (Cout == CO_ && Cin == CI_ && Fin == FI_ && Fout == 2 * FI_) {                \
        rc = launch_conv_tc<3, 1, 2, CO_, CI_, FI_, 1, NS_>(a, st);                  \
        return rc ? rc : 1;                                                          \
    }
    CRUSE_CT_TD(8, 16, 64, 2)
    CRUSE_CT_TD(16, 32, 32, 2)
    CRUSE_CT_TD(32, 64, 16, 1)
#undef CRUSE_CT_TD
    return 0;
}

// data gradient of the (2,3)/stride-(1,2) encoder conv (MODE 2); w is the Conv2d weight [Cout][Cin][2][3]
int conv_dgrad_tc_try(const float* dz, const float* w, const float* addend, float* din, int B, int T, int Cin, int Fin, int Cout,
                      int Fout, int kt, cudaStream_t st) {
    if (!conv_tc_enabled() || kt != 2) return 0;
    if ((reinterpret_cast<uintptr_t>(dz) & 15) || (reinterpret_cast<uintptr_t>(din) & 15) || (reinterpret_cast<uintptr_t>(addend) & 15)) return 0;
    ConvTcArgs a{dz, w, nullptr, nullptr, nullptr, nullptr, addend, din, B, T, CRUSE_ACT_NONE, 0, 0, nullptr, 0, 0, 0, nullptr, nullptr};
    int rc = 0;
#define CRUSE_CT_CD(CO_, CI_, FO_, GM_, NS_)                                         \
    if (Cout == CO_ && Cin == CI_ && Fout == FO_ && Fin == 2 * FO_) {                \
        rc = launch_conv_tc<2, 2, 1, CO_, CI_, FO_, GM_, NS_>(a, st);                \
        return rc ? rc : 1;                                                          \
    }
    CRUSE_CT_CD(64, 32, 16, 1, 1)
    CRUSE_CT_CD(32, 16, 32, 1, 2)
    CRUSE_CT_CD(16, 8, 64, 1, 2)
#undef CRUSE_CT_CD
    return 0;
}

int convT_tc_try(const float* in, const float* w, const float* bias, const float* scale, const float* shift, const float* alpha,
                 int act, const float* skip, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout, cudaStream_t st,
                 int t_begin, int t_end) {
    if (!conv_tc_enabled()) return 0;
    if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(skip) & 15)) return 0;
    ConvTcArgs a{in, w, bias, scale, shift, alpha, skip, out, B, T, act, 0, 0, nullptr, t_begin, t_end, 0, nullptr, nullptr};
    int rc = 0;
#define CRUSE_CT_CONVT(CI_, CO_, FI_, GM_, NS_)                                      \
    if (Cin == CI_ && Cout == CO_ && Fin == FI_ && Fout == 2 * FI_) {                \
        rc = launch_conv_tc<1, 1, 1, CI_, CO_, FI_, GM_, NS_>(a, st);                \
        return rc ? rc : 1;                                                          \
    }
    CRUSE_CT_CONVT(64, 32, 16, 1, 1)
    CRUSE_CT_CONVT(32, 16, 32, 1, 2)
    CRUSE_CT_CONVT(16, 8, 64, 2, 2)
#undef CRUSE_CT_CONVT
    return 0;
}

}  // namespace cruse

extern "C" int cruse_conv_get_mode(void) { return cruse::conv_mode(); }
extern "C" int cruse_conv_set_mode(int mode) {
    if (mode != 0 && mode != 1) {
        cruse::set_error("conv_set_mode: mode must be 0 (fp32 CUDA cores) or 1 (tf32 tensor cores), got %d", mode);
        return -1;
    }
    cruse::g_conv_mode = mode;
    return 0;
}

extern "C" int cruse_conv_set_max_ctas(int n) {
    if (n < 0) {
        cruse::set_error("conv_set_max_ctas: n must be >= 0, got %d", n);
        return -1;
    }
    cruse::g_conv_max_ctas = n;
    return 0;
}

#ifdef CRUSE_CT_TIMING
extern "C" int cruse_debug_ct_timing(long long* out_host, int n) {
    return (int)cudaMemcpyFromSymbol(out_host, cruse::g_ct_timing, sizeof(long long) * n);
}
#endif

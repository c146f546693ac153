/* cruse_b200.h -- C ABI of libcruse_sm100.so: the CRUSE hot path as sm_100a CUDA kernels.
 *
 * Drop-in boundary.  The reference (Okrio/CRUSE) is pure Python and has
 * no FFI of its own; every arithmetic call on its hot path goes to a PyTorch library op.
 * Each entry point below replaces one of those call sites (cited as reference file:line,
 * paths relative to the reference root) and is what a ctypes stub in the reference would
 * bind.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.  All pointers are DEVICE
 *    pointers (fp32 unless stated) owned by the caller; kernels never allocate, free or
 *    retain them.  `stream` is a cudaStream_t passed as void*.  All launches are
 *    asynchronous on that stream and CUDA-graph capturable.
 *  - return 0 on success, negative on error; cruse_last_error() returns a thread-local
 *    message for the last failing call on this thread.
 *  - activation tensors are FRAME-MAJOR:  [B, T, C, F]  (one STFT frame = C*F contiguous
 *    floats).  For C == 1 this is bit-identical to the reference's [B, 1, T, F]; the GRU
 *    view [B, T, C*F'] (model/cruse_net.py:39-40) is free.  Spectra are complex-interleaved
 *    [B, T, NF, 2] with NF = n_fft/2 + 1.
 *  - parameters are read in PyTorch's native layouts (Conv2d [Cout,Cin,kh,kw],
 *    ConvTranspose2d [Cin,Cout,kh,kw], GRU weight_* [3H, I] rows ordered r,z,n).
 */
#ifndef CRUSE_B200_H
#define CRUSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRUSE_MAX_GROUPS 8

/* activation codes shared by conv / convT / bn_act kernels */
#define CRUSE_ACT_NONE    0
#define CRUSE_ACT_RELU    1   /* model/cruse_net.py:145 (nn.ReLU named elu)          */
#define CRUSE_ACT_PRELU   2   /* optional per-channel PReLU (model/mtfaa.py:170)     */
#define CRUSE_ACT_SIGMOID 3   /* model/cruse_net.py:164                              */

/* addressing of a complex [B,T,F] tensor: element (b,t,f) has its real part at b*sb + t*st + f*sf floats, the imaginary part
 * im_off floats further (see the loss entries below) */
typedef struct { long long sb, st, sf, im_off; } cruse_cplx_layout;

#define CRUSE_PAD_REFLECT  0  /* train_base/acoustics/feature.py:22-30 (torch.stft default) */
#define CRUSE_PAD_CONSTANT 1  /* utils/utils.py:396                                           */

int cruse_version(void);
const char* cruse_last_error(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int cruse_sm_count(void);

/* ---- a1: STFT.  replaces torch.stft at train_base/acoustics/feature.py:22-30 and
 *      utils/utils.py:390-396, fused with the magnitude of utils/utils.py:400.
 *  wav [B,L]; window [n_fft]; spec [B,T,NF,2]; mag [B,T,mag_bins] or NULL
 *  (mag = sqrt(re^2+im^2+mag_eps) of bins [0,mag_bins)).  T must be 1 + L/hop (center=True). */
int cruse_stft_fwd(const float* wav, const float* window, float* spec, float* mag,
                   int B, int L, int n_fft, int hop, int T, int pad_mode,
                   int mag_bins, float mag_eps, void* stream);

/* ---- a6+a7: mask*spectrum + iSTFT.  replaces PreProcess.masking utils/utils.py:417-433
 *      (mag_mapping) and torch.istft at feature.py:53-61 / utils/utils.py:448-454.
 *  spec [B,T,NF,2]; mask [B,T,mask_bins] or NULL (bins >= mask_bins pass through);
 *  est_spec [B,T,NF,2] or NULL receives mask*spec; wav [B,L] or NULL receives the
 *  overlap-add reconstruction (hann-squared envelope normalised, trimmed n_fft/2). */
int cruse_mask_istft_fwd(const float* spec, const float* mask, const float* window,
                         float* est_spec, float* wav,
                         int B, int L, int n_fft, int hop, int T, int mask_bins, void* stream);
/* backward of the iSTFT half (a9 for time-domain losses; the reference gets it from autograd of torch.istft, feature.py:53-61):
 * dwav [B,L] -> dspec [B,T,NF,2] = gradient w.r.t. the (real, imag) spectrum handed to cruse_mask_istft_fwd; feed it to
 * cruse_mask_bwd for the mask gradient.  ws: cruse_istft_bwd_ws_bytes(B, L) bytes of scratch. */
size_t cruse_istft_bwd_ws_bytes(int B, int L);
int cruse_istft_bwd(const float* dwav, const float* window, float* dspec, void* ws, int B, int L, int n_fft, int hop, int T,
                    void* stream);
/* The CTAs [c_begin, c_end) of the same launch: CTA c produces est_spec of the frames [c*FC, (c+1)*FC) and the samples
 * those frames start, reading the mask of the frames [c*FC - (n_fft-1)/hop, (c+1)*FC), FC = cruse_mask_istft_chunk_frames.
 * Disjoint ranges covering [0, ceil(T/FC)) are bit-identical to the whole call. */
int cruse_mask_istft_chunk_frames(int n_fft, int hop);
int cruse_mask_istft_fwd_range(const float* spec, const float* mask, const float* window, float* est_spec, float* wav,
                               int B, int L, int n_fft, int hop, int T, int mask_bins, int c_begin, int c_end, void* stream);

/* backward of the mask apply: dmask[b,t,f] = gscale * (dre*Xre + dim*Xim), f < mask_bins.
 * dest [B,T,NF,2] (row stride NF) ; gscale: device scalar or NULL (=1).  mask [B,T,mask_bins] or NULL: when given
 * the result is multiplied by mask*(1-mask), i.e. it is the gradient BEFORE the sigmoid of model/cruse_net.py:164. */
int cruse_mask_bwd(const float* dest, const float* spec, const float* gscale, const float* mask, float* dmask,
                   int B, int T, int NF, int mask_bins, void* stream);

/* ---- a2/a3: causal strided conv stage.  replaces nn.Conv2d + slice + BatchNorm2d + act at
 *      model/cruse_net.py:138,141,149-152 (kt=2, fstride=2) and the skip convs :143,153-156
 *      (kt=1, fstride=1).  Kernel (kt,3), freq padding 1, time taps look back only.
 *  in [B,T,Cin,Fin]; w [Cout,Cin,kt,3]; bias [Cout]|NULL; out [B,T,Cout,Fout].
 *  hist [B,Cin,Fin]|NULL: the input frame preceding t=0 (streaming chunks; NULL = zero padding, i.e.
 *  the start of an utterance; state-carry convention of model/based_model/cust_conv.py:303-325).
 *  epilogue: v = conv+bias; if scale: v = v*scale[c]+shift[c]; v = act(v) (alpha = PReLU slopes).
 *  stats_ws (optional, train-mode BN): per-CTA partial sums of the pre-affine value,
 *  layout [nparts][2*Cout]; nparts = cruse_conv_nparts(B,T). */
int cruse_conv_fwd(const float* in, const float* hist, const float* w, const float* bias,
                   const float* scale, const float* shift, const float* alpha, int act,
                   float* out, float* stats_ws,
                   int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride, void* stream);
int cruse_conv_nparts(int B, int T);
/* the same stage (eval mode: no hist, no statistics) with the frame records of `in` and / or `out` ordered TIME-MAJOR
 * ([T, B, C, F] instead of [B, T, C, F]): the last encoder stage writes the GRU input time-major so that the input
 * projections of the two-layer wavefront run per chunk of frames on contiguous rows.  Tensor-core instantiations only
 * (returns an error for shapes without one or in fp32 conv mode). */
int cruse_conv_fwd_tm(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                      const float* alpha, int act, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout,
                      int kt, int fstride, int in_time_major, int out_time_major, void* stream);
/* numeric mode of the eval-mode conv / convT stages of the 256-bin pyramid (no `hist`, no `stats_ws`):
 *   1 = implicit GEMM on the tensor cores (tcgen05.mma kind::tf32, fp32 accumulate; default),
 *   0 = exact-fp32 CUDA-core kernels.  Initial value from the environment (CRUSE_CONV=fp32|tf32).
 * Stages without a tensor-core instantiation (Cin == 1, Cout == 1, odd bin counts) always run in fp32. */
int cruse_conv_get_mode(void);
int cruse_conv_set_mode(int mode);
/* cap the persistent grid of the tensor-core conv stages launched from now on (0 = one CTA per SM, the default):
 * used to run the skip convs on the SMs the GRU wavefront leaves free. */
int cruse_conv_set_max_ctas(int n);

/* ---- a5: decoder stage.  replaces nn.ConvTranspose2d((1,3), stride (1,2)) + crop + BN + act
 *      + skip add at model/cruse_net.py:161-164.   w [Cin,Cout,1,3]; output cropped to Fout
 *      (<= 2*Fin+1).  Epilogue as cruse_conv_fwd, then `+ skip` ([B,T,Cout,Fout] or NULL). */
int cruse_convT_fwd(const float* in, const float* w, const float* bias,
                    const float* scale, const float* shift, const float* alpha, int act,
                    const float* skip, float* out, float* stats_ws,
                    int B, int T, int Cin, int Fin, int Cout, int Fout, void* stream);

/* ---- train-mode BatchNorm2d (model/cruse_net.py:141-142): reduce the per-CTA partials,
 *      emit scale/shift (and save mean/invstd), update running stats (momentum, unbiased var). */
int cruse_bn_finalize(const float* stats_ws, int nparts, int C, double count,
                      const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var,
                      float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
/* eval-mode BatchNorm2d folded to an affine: scale = gamma/sqrt(var+eps), shift = beta - mean*scale */
int cruse_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                  float eps, float* scale, float* shift, int C, void* stream);
/* the same for up to 8 BatchNorm2d layers in ONE launch (all eval-mode folds of a forward pass): HOST arrays of n
 * device pointers / sizes; gamma / beta tables (or single entries) may be NULL. */
int cruse_bn_fold_many(const float* const* gamma, const float* const* beta, const float* const* running_mean,
                       const float* const* running_var, const float* eps, float* const* scale, float* const* shift,
                       const int* C, int n, void* stream);
/* y = act(z*scale[c]+shift[c]) (+ skip); z,y [B,T,C,F] */
int cruse_bn_act_fwd(const float* z, const float* scale, const float* shift, const float* alpha, int act,
                     const float* skip, float* y, long long n_frames, int C, int F, void* stream);

/* Encoder stage WITH the skip conv of its input fused in (eval mode; replaces one cruse_conv_fwd + one skip cruse_conv_fwd:
 * model/cruse_net.py:149-152 for stage k+1 and :153-155 for skip_connect_k, which read the same tensor e_k):
 *   out      [B,T,Cout,Fout]  = act(BN(Conv2d((2,3), stride (1,2), pad (1,1))(in)[..., :-1, :]))   (time-major records if out_time_major)
 *   out_skip [B,T,Cin,Fin]    = Conv2d((1,3), pad (0,1), bias=False)(in)        w_skip [Cin][Cin][1][3]
 * One pass over `in`: the stage's shared-memory tile gets a fourth frequency tap and the skip outputs are extra accumulator columns.
 * Tensor-core instantiations only (Cin 8 / 16 of the 256-bin pyramid); t_begin = t_end = 0 means all frames. */
int cruse_conv_skip_fwd(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                        const float* alpha, int act, const float* w_skip, float* out, float* out_skip, int B, int T,
                        int Cin, int Fin, int Cout, int Fout, int out_time_major, int t_begin, int t_end, void* stream);
/* Eval-mode stages restricted to the OUTPUT frames [t_begin, t_end) of every utterance (same argument meaning as
 * cruse_conv_fwd_tm / cruse_convT_fwd / cruse_layernorm_fwd; out / y are the full-size tensors, rows outside the range are
 * not touched).  The net is causal (model/cruse_net.py:138 pads time on the left only), so a frame range of a stage needs
 * the same range -- plus one earlier frame for the (2,3) convs -- of the stage below: the first frames of the encoder
 * can be through before the rest, and the decoder can follow the last GRU layer chunk by chunk.  Tensor-core / streaming
 * kernels only (256-bin pyramid, tf32 conv mode): any other shape is an error, not a fallback. */
int cruse_conv_fwd_range(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                         const float* alpha, int act, float* out, int B, int T, int Cin, int Fin, int Cout, int Fout,
                         int kt, int fstride, int in_time_major, int out_time_major, int t_begin, int t_end, void* stream);
int cruse_convT_fwd_range(const float* in, const float* w, const float* bias, const float* scale, const float* shift,
                          const float* alpha, int act, const float* skip, float* out, int B, int T, int Cin, int Fin,
                          int Cout, int Fout, int t_begin, int t_end, void* stream);
int cruse_layernorm_fwd_range(const float* x, const float* gamma, const float* beta, float eps, const float* residual,
                              float* y, int B, int T, int D, int t_begin, int t_end, void* stream);
/* LayerNorm 2 (+ skip 4) and the whole decoder of the 256-bin pyramid for the frames [t_begin, t_end), ONE launch, eval mode:
 * replaces model/cruse_net.py:51 (ln2), :160 (+ skip_connect_4) and :161-164 repaired (3 x [ConvTranspose2d(1,3)/s(1,2) + folded
 * BatchNorm + ReLU/PReLU + skip], ConvTranspose2d(8->1) + sigmoid).  One warp per frame, activations stay in shared memory,
 * warp-level tf32 tensor-core GEMMs (fp32 accumulate), stage 1 and the LayerNorm in fp32.
 *   cruse_decoder_fused_prep lays the constants of all four stages out once per forward pass (after the BatchNorm fold) as the
 *   kernel's shared-memory image (cruse_decoder_fused_image_floats() floats, 16-byte aligned): w / bias = HOST arrays of 4 device
 *   pointers in decoder order (conv4_t.weight [64,32,1,3], conv3_t [32,16,1,3], conv2_t [16,8,1,3], conv1_t [8,1,1,3]); scale /
 *   shift / alpha = HOST arrays of 3 device pointers (folded BatchNorm and PReLU slopes of stages 4..2; alpha may be NULL for ReLU).
 *   cruse_decoder_fused_range: y2 [B,T,1024] (GRU layer-2 output, feature = channel*16 + bin) -> mask [B,T,256]; skips = HOST array of
 *   4 device pointers: skip4 [B,T,64,16] (added to the LayerNorm output), skip3 [B,T,32,32], skip2 [B,T,16,64], skip1 [B,T,8,128];
 *   max_ctas > 0 caps the grid (running beside the recurrences).  loss_rows (optional, with ref = clean spectrum S and unproc =
 *   noisy spectrum X in any complex layout): the frame's share of wo_male (loss_func/loss.py:121-148) on est = mask * X over the
 *   256 bins, one partial sum per frame, formed where the mask is produced (same per-bin arithmetic as cruse_wo_male_masked_fwd).
 *   skip_convs = 1 (image prepared with wskip4 = skip_connect_4.weight [64,64,1,3] and wskip3 = skip_connect_3.weight [32,32,1,3],
 *   model/cruse_net.py:143): the two deepest skip convs (:154-155) are computed inside the launch from the encoder outputs --
 *   skips[0] = e4 TIME-MAJOR [T,B,64,16] (the GRU input as the last encoder stage writes it), skips[1] = e3 [B,T,32,32]; their
 *   skip tensors never exist in HBM and the two skip-conv launches disappear.
 * Any other geometry is an error, not a fallback (use the per-stage entry points). */
long long cruse_decoder_fused_image_floats(int with_skip_convs);
int cruse_decoder_fused_prep(const float* const* w, const float* const* bias, const float* const* scale, const float* const* shift,
                             const float* const* alpha, int act, const float* wskip4, const float* wskip3, float* image, void* stream);
int cruse_decoder_fused_range(const float* y2, const float* ln_gamma, const float* ln_beta, float ln_eps,
                              const float* const* skips, int skip_convs, const float* image, float* mask, const float* ref,
                              cruse_cplx_layout lref, const float* unproc, cruse_cplx_layout lunp, float* loss_rows, int B, int T,
                              int t_begin, int t_end, int max_ctas, void* stream);
/* the per-frame sums loss_rows[B*T] of cruse_decoder_fused_range -> loss = sum / (B*T*F) (loss.py:147), fixed order */
int cruse_wo_male_finish_rows(const float* rows, int B, int T, int F, float* loss, void* stream);

/* ---- a4: grouped GRU.  replaces nn.GRU(H,H) x groups at model/cruse_net.py:23-31,43-50.
 *  ih GEMM: xproj[m, g, :] = x[m, g*H:(g+1)*H] . w_ih[g]^T + b_ih[g] (+ b_hh[g] for r,z rows)
 *  x [M, G*H] (M = B*T); xproj [M, G, 3H].  w_ih/b_ih/b_hh: HOST arrays of G device pointers. */
int cruse_gru_ih_gemm(const float* x, const float* const* w_ih, const float* const* b_ih,
                      const float* const* b_hh, float* xproj, int M, int G, int H, void* stream);
/* same contract on the tensor cores: tcgen05.mma kind::tf32 (fp32 operands rounded to tf32 by the TMA
 * unit, fp32 accumulation in TMEM).  Needs 16-byte aligned x / w_ih rows (H % 4 == 0). */
int cruse_gru_ih_gemm_tc(const float* x, const float* const* w_ih, const float* const* b_ih,
                         const float* const* b_hh, float* xproj, int M, int G, int H, void* stream);
/* developer A/B switch for cruse_gru_ih_gemm_tc / _tm_tc at H = 256: 1 (default) = the A-stationary kernel (activation tile resident,
 * weights streamed, double-buffered accumulator, one CTA per (128 rows, group)); 0 = one 128 x 256 tile per CTA */
int cruse_gemm_set_astat(int on);
/* recurrence over T with W_hh resident on chip (thread-block cluster per (group, 8-utterance slice)).
 *  y[b,t, j*y_fs + g*y_gs] = h_t[g][b][j]   (layer 1: y_fs=G,y_gs=1 = the stack/flatten interleave of
 *  cruse_net.py:43-45; layer 2: y_fs=1,y_gs=H = cat, :49-50).  h0/hT [G,B,H] or NULL (state carry,
 *  model/based_model/cust_conv.py:303-325). */
int cruse_gru_seq_fwd(const float* xproj, const float* const* w_hh, const float* const* b_hh,
                      const float* h0, float* y, float* hT,
                      int B, int T, int G, int H, int y_fs, int y_gs, void* stream);
/* same contract with the per-step W_hh.h product on the tensor cores (tcgen05.mma kind::tf32, W_hh slice
 * resident in tensor memory, accumulators in TMEM; one cluster of ceil(H/32) CTAs per (group, 2 x 16 utterances)).
 * Only the matmul operands are rounded to tf32; gate math and the z*h carry stay fp32.  H % 4 == 0, H <= 256.
 * gates [B,T,G,4,H] or NULL: saves r, z, n and (W_hn.h + b_hn) of every step for the backward pass. */
int cruse_gru_seq_fwd_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh,
                         const float* h0, float* y, float* hT, float* gates,
                         int B, int T, int G, int H, int y_fs, int y_gs, void* stream);

/* the two entry points of the time-chunked two-layer wavefront (layer 2 of model/cruse_net.py:47-50 starts on chunk k
 * while layer 1, :41-45, is still running chunk k+1; GRU-internal buffers are TIME-MAJOR [T, B, ...] so that a chunk of
 * frames is one contiguous row range):
 *  cruse_gru_ih_gemm_tm_tc: as cruse_gru_ih_gemm_tc for x [B*T, G*H] in frame order, but row (b,t) of the result is
 *    written to row t*B + b of xproj [T, B, G, 3H].
 *  cruse_gru_seq_chunk_tc: Tc steps of the recurrence from state h0 to state hT (both [G,B,H], hT may be NULL);
 *    row (b, t) of xproj / y is b*x_bs + t*x_ts / b*y_bs + t*y_ts rows (a row = G*3H / G*H floats) past the pointers,
 *    which the caller has already advanced to the chunk's first frame. */
int cruse_gru_ih_gemm_tm_tc(const float* x, const float* const* w_ih, const float* const* b_ih,
                            const float* const* b_hh, float* xproj, int B, int T, int G, int H, void* stream);
int cruse_gru_seq_chunk_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh,
                           const float* h0, float* y, float* hT, int B, int Tc, int G, int H, int y_fs, int y_gs,
                           long long x_bs, long long x_ts, long long y_bs, long long y_ts, void* stream);

/* The same wavefront WITHOUT relaunching the recurrence per chunk: one launch per layer over all T frames; before the
 * x-projections of frame bounds[k] are fetched the kernel waits (bounded spin, 2 s, then *err = 1) until
 * wait_flags[k] >= wait_target, and after frame bounds[k+1]-1 has been stored every (CTA, slice) adds 1 to done_flags[k]
 * (release), i.e. done_flags[k] reaches G * ceil(H/32) * ceil(B/16).  cruse_flag_set / cruse_flag_wait are the one-thread
 * kernels the host queues behind a producer / in front of a consumer on its stream.  bounds: HOST array of nchunks+1
 * frame indices (0 .. T, chunks of at least 8 frames); wait_flags / done_flags / err: device, may be NULL. */
int cruse_gru_seq_flagged_tc(const float* xproj, const float* const* w_hh, const float* const* b_hh, float* y,
                             int B, int T, int G, int H, int y_fs, int y_gs,
                             long long x_bs, long long x_ts, long long y_bs, long long y_ts,
                             const int* bounds, int nchunks, const unsigned* wait_flags, unsigned wait_target,
                             unsigned* done_flags, int* err, void* stream);
int cruse_flag_wait(const unsigned* flag, unsigned target, int* err, void* stream);
int cruse_flag_set(unsigned* flag, unsigned value, void* stream);
/* Loud failure of the flag-synchronised wavefront: when *err != 0 (a bounded spin of cruse_gru_seq_flagged_tc / cruse_flag_wait
 * timed out after 2 s -- another process or stream held the SMs the producers needed) every listed output buffer (HOST array of
 * up to 4 DEVICE pointers with their element counts) is filled with NaN, so that a broken dependency can never return a plausible
 * mask / waveform / loss (reference convention: errors are exceptions, loss_func/loss.py:65-68; the host raises when it reads the flag). */
int cruse_poison_on_error(const int* err, float* const* bufs, const long long* counts, int nbufs, void* stream);
/* Developer instrumentation (tools/wavefront_trace.py), not part of the data path: the next cruse_gru_seq_flagged_tc launches write a
 * progress trace of their cluster 0 (%globaltimer of every 8th step [128] and of begin / end of every chunk wait [2*16]) to
 * device_buf + (k%2)*160 (uint64) for launch k = 0, 1, ... (layer 1, layer 2 of a wavefront); NULL switches it off. */
int cruse_debug_seq_trace(void* device_buf);

/* how many clusters of the tcgen05 recurrence kernel the current device can hold at once (each serves two
 * software-pipelined slices of 16 utterances of one group); G*ceil(B/32) above this runs in waves.  <0 on error. */
int cruse_gru_seq_tc_max_clusters(int H);

/* ---- nn.LayerNorm(D) at model/cruse_net.py:32-33,46,51.  x,y [rows, D]; mean/rstd [rows] or NULL.
 *      residual [rows, D] or NULL is added after the affine (fuses `+ skip4`, cruse_net.py:160). */
int cruse_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps,
                        const float* residual, float* y, float* mean, float* rstd,
                        long long rows, int D, void* stream);
/* LayerNorm 1 of the bottleneck fused with the stack(dim=-1)+flatten interleave of model/cruse_net.py:43-46:
 * x [rows, D] holds the G group outputs CONCATENATED ([G][D/G]), y[row, h*G + g] = LN(x)[g*H + h]*gamma[h*G+g]+beta[h*G+g]
 * (statistics are permutation invariant).  G == 4. */
int cruse_layernorm_interleave_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y,
                                   long long rows, int D, int G, void* stream);

/* ---- a8: weighted-magnitude loss wo_male, loss_func/loss.py:121-148, forward + d/d(est).
 *  Complex tensors are addressed as  re = p[b*sb + t*st + f*sf], im = p[... + im_off]  so both the
 *  reference layout [B,2,T,F] and the internal [B,T,NF,2] are accepted without a copy.
 *  loss: device scalar.  dest (optional): gradient of the loss w.r.t. est in est's layout
 *  (same strides), unscaled by any upstream gradient.  ws: >= cruse_wo_male_ws_bytes() bytes. */
int cruse_wo_male_fwd_bwd(const float* ref, cruse_cplx_layout lref, const float* est, cruse_cplx_layout lest,
                          const float* unproc, cruse_cplx_layout lunp, float* dest,
                          float* loss, void* ws, int B, int T, int F, void* stream);
/* forward value only, with the estimate formed on the fly as mask[b,t,f] * unproc[b,t,f] (mask [B,T,F] contiguous):
 * PreProcess.masking (utils/utils.py:417-433, mag_mapping) fused into the loss, so the loss does not wait for the
 * mask*spectrum + iSTFT kernel and runs beside it.  Bit-identical to cruse_wo_male_fwd_bwd on the stored estimate. */
int cruse_wo_male_masked_fwd(const float* ref, cruse_cplx_layout lref, const float* mask, const float* unproc,
                             cruse_cplx_layout lunp, float* loss, void* ws, int B, int T, int F, void* stream);
size_t cruse_wo_male_ws_bytes(void);
/* The other spectral-domain modes of the dispatcher (loss_func/loss.py:31-34) on ref / est in any complex layout:
 * mode 0 = rmse (:59-78): sum_c |est_c - ref_c| / (B*T*F);  mode 1 = c_rmse (:88-118): power-law compressed complex error with
 * c = beta = 0.3, arithmetic kept literally (incl. the tmp1/tmp2 mix at :107-109), plain sum.  dest (NULL or est's layout)
 * receives d loss / d est (0 where |est| = 0, where torch's autograd yields inf / NaN).  ws: cruse_wo_male_ws_bytes(). */
int cruse_spec_loss_fwd_bwd(int mode, const float* ref, cruse_cplx_layout lref, const float* est, cruse_cplx_layout lest,
                            float* dest, float* loss, void* ws, int B, int T, int F, void* stream);
/* SI-SNR of loss_func/loss.py:37-56 on waveforms est / ref [B,L]: value = mean_b 10 log10(|alpha ref|^2 /
 * (|est - alpha ref|^2 + eps) + eps), alpha = <est,ref> / (<ref,ref> + eps); the dispatcher's 'SI-SNR' mode (:25-26) returns
 * its negative.  cruse_sisnr_bwd: dest = gscale * d value / d est (gscale: device scalar or NULL = 1), from the coefficients
 * cruse_sisnr_fwd left in ws (cruse_sisnr_ws_bytes(B) bytes).  Fixed-order reductions: deterministic. */
size_t cruse_sisnr_ws_bytes(int B);
int cruse_sisnr_fwd(const float* est, const float* ref, float* value, void* ws, int B, int L, float eps, void* stream);
int cruse_sisnr_bwd(const float* est, const float* ref, const void* ws, const float* gscale, float* dest, int B, int L,
                    void* stream);
/* The zero-mean SI-SNR loss of the trainer's loss factory (train_base/loss.py:7-25 `si_snr_loss()`, selected by name in
 * tools/train_stand.py:73-75):  -mean_b 20 log10(eps + |t| / (|x_zm - t| + eps)),  t = <x_zm,s_zm> s_zm / (|s_zm|^2 + eps).
 * est = x, ref = s, both [B, L]; value: device scalar; ws: cruse_si_snr_zm_ws_bytes(B) bytes, kept for the backward call.
 * cruse_si_snr_zm_bwd: dest[b,i] = gscale * d loss / d est[b,i] (gscale: device scalar or NULL = 1). */
size_t cruse_si_snr_zm_ws_bytes(int B);
int cruse_si_snr_zm_fwd(const float* est, const float* ref, float* value, void* ws, int B, int L, float eps, void* stream);
int cruse_si_snr_zm_bwd(const float* est, const float* ref, const void* ws, const float* gscale, float* dest, int B, int L,
                        void* stream);
/* The same loss range by range (inference schedule: the loss follows the decoder instead of waiting for the whole mask):
 * _partial_range writes nparts partial sums of the frames [t_begin, t_end) of every utterance into ws[p_off, p_off+nparts),
 * cruse_wo_male_finish adds up the nparts partials of all ranges and divides by B*T*F (loss.py:147).  Same arithmetic per
 * bin as cruse_wo_male_masked_fwd; the summation order differs, so the value agrees to fp32 rounding (1e-6), not bitwise. */
int cruse_wo_male_masked_partial_range(const float* ref, cruse_cplx_layout lref, const float* mask, const float* unproc,
                                       cruse_cplx_layout lunp, void* ws, int p_off, int nparts, int B, int T, int F,
                                       int t_begin, int t_end, void* stream);
int cruse_wo_male_finish(const void* ws, int nparts, int B, int T, int F, float* loss, void* stream);

/* =====================================================================================================
 * a9: backward of a2-a8 (what autograd + cuDNN/ATen compute on the reference path for the modules of
 * model/cruse_net.py:14-55,129-165).  Gradients are written (not accumulated) into caller buffers.
 * ===================================================================================================== */

/* out[j] (+)= sum_p ws[p*n + j]: fixed-order reduction of per-CTA partials */
int cruse_colsum(const float* ws, int nparts, int n, float* out, int accumulate, void* stream);

/* data gradient of cruse_conv_fwd (same B,T,Cin,Fin,Cout,Fout,kt,fstride as the forward call):
 *   din[B,T,Cin,Fin] = conv^T(dz[B,T,Cout,Fout], w) (+ addend[B,T,Cin,Fin] or NULL) */
int cruse_conv_dgrad(const float* dz, const float* w, const float* addend, float* din,
                     int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride, void* stream);
/* weight / bias gradient of cruse_conv_fwd: dw [Cout,Cin,kt,3], dbias [Cout] or NULL.
 * ws: >= cruse_conv_wgrad_ws_bytes(...) bytes of scratch. */
int cruse_conv_wgrad(const float* in, const float* dz, float* dw, float* dbias, void* ws,
                     int B, int T, int Cin, int Fin, int Cout, int Fout, int kt, int fstride, void* stream);
size_t cruse_conv_wgrad_ws_bytes(int B, int T, int Cin, int Fin, int Cout, int Fout, int kt);
/* same for cruse_convT_fwd (w, dw in ConvTranspose2d layout [Cin,Cout,1,3]) */
int cruse_convT_dgrad(const float* dz, const float* w, const float* addend, float* din,
                      int B, int T, int Cin, int Fin, int Cout, int Fout, void* stream);
int cruse_convT_wgrad(const float* in, const float* dz, float* dw, float* dbias, void* ws,
                      int B, int T, int Cin, int Fin, int Cout, int Fout, void* stream);
size_t cruse_convT_wgrad_ws_bytes(int B, int T, int Cin, int Fin, int Cout, int Fout);

/* backward of  y = act(z*scale[c]+shift[c]) (+skip)  with BatchNorm2d in front (z = conv output, pre-BN):
 *  pass 1 (reduce):   partials[nparts][3*C] = per-CTA { sum da, sum da*xhat, sum dy*a*[a<=0] },
 *                     a = z*scale+shift, da = dy*act'(a), xhat = (z-mean[c])*invstd[c]; nparts = cruse_bn_bwd_nparts()
 *  pass 2 (finalize): dgamma = S2, dbeta = S1, dalpha = S3 (PReLU slope); coef[3*C] = {A, M1, M2} with
 *                     training: A = gamma*invstd, M1 = S1/count, M2 = S2/count; eval: A = gamma*invstd, M1 = M2 = 0
 *  pass 3 (apply):    dz = A*(da - M1 - xhat*M2) */
int cruse_bn_bwd_nparts(long long n_frames);
int cruse_bn_act_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, const float* alpha,
                            int act, const float* mean, const float* invstd, float* partials,
                            long long n_frames, int C, int F, void* stream);
int cruse_bn_bwd_finalize(const float* partials, int nparts, int C, double count, const float* gamma, const float* invstd,
                          int training, float* dgamma, float* dbeta, float* dalpha, float* coef, void* stream);
int cruse_bn_act_bwd_apply(const float* dy, const float* z, const float* scale, const float* shift, const float* alpha,
                           int act, const float* mean, const float* invstd, const float* coef, float* dz,
                           long long n_frames, int C, int F, void* stream);

/* dz = dy * y * (1 - y): backward of the mask sigmoid (model/cruse_net.py:164) given its output y */
int cruse_sigmoid_bwd(const float* dy, const float* y, float* dz, long long n, void* stream);

/* nn.LayerNorm backward.  x, dy, dx [rows, D]; mean/rstd [rows] from cruse_layernorm_fwd.
 * partials [cruse_layernorm_bwd_nparts(rows)][2*D] = per-CTA { dgamma, dbeta }; reduce with cruse_colsum. */
int cruse_layernorm_bwd_nparts(long long rows);
int cruse_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                        float* dx, float* partials, long long rows, int D, void* stream);

/* grouped-GRU backpropagation through time (cuDNN RNN backward on the reference path, model/cruse_net.py:23-31,43-50).
 *  dy [B,T,G*H]: gradient w.r.t. the layer output, addressed like y (y_fs, y_gs); y: the forward output (h_t);
 *  gates [B,T,G,4,H] from cruse_gru_seq_fwd_tc; h0 [G,B,H] or NULL.
 *  dxproj [B,T,G,3H] = (da_r, da_z, da_n): gradient of the input projections (and of b_ih; its r,z part also of b_hh);
 *  dpre   [B,T,G,3H] = (da_r, da_z, dhn):  gradient of W_hh.h + b_hh;  dh0 [G,B,H] or NULL.
 *  dbias_part [ceil(B/16), G, 4, H] or NULL, ZERO-INITIALISED by the caller: per-utterance-slice sums over t of
 *  (da_r, da_z, da_n, dhn); column-sum over the first dim gives db_ih = (r,z,n) and db_hh = (r,z,hn). */
int cruse_gru_seq_bwd_tc(const float* dy, const float* y, const float* gates, const float* h0,
                         const float* const* w_hh, float* dxproj, float* dpre, float* dh0, float* dbias_part,
                         int B, int T, int G, int H, int y_fs, int y_gs, void* stream);
/* One streaming step of the grouped GRU for B concurrent utterances (BASELINE cfg-5; the reference's state-carry API is
 * GroupedGRULayer.forward(input, h0) -> (out, h), model/based_model/cust_conv.py:303-325, applied to a 1-frame input):
 * hproj = W_hh . h_prev as ONE tcgen05 GEMM per group (tf32 operands), then the gate math (r,z,n order of nn.GRU,
 * model/cruse_net.py:23-31) elementwise.  xproj [B,G,3H] from cruse_gru_ih_gemm(_tc) with M = B; h_prev [G,B,H] or NULL
 * (zero state); h_new [G,B,H] (must not alias h_prev); y[b, j*y_fs + g*y_gs] = h_new[g][b][j];
 * ws: cruse_gru_step_ws_bytes(B,G,H) bytes of scratch. */
size_t cruse_gru_step_ws_bytes(int B, int G, int H);
int cruse_gru_step(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h_prev,
                   float* h_new, float* y, void* ws, int B, int G, int H, int y_fs, int y_gs, void* stream);
/* G independent GEMMs on tcgen05 (tf32 operands, fp32 accumulate):  C_g[m,n] = sum_k A_g[m,k] * B_g[n,k] (+ bias_g[n]).
 * A_g [M,K] row pitch lda, B_g [N,K] row pitch ldb (both K-major), C_g row pitch ldc (floats).  splitk > 1 writes
 * splitk partial planes C_g + s*c_plane (bias must be NULL); sum them with cruse_colsum.
 * Used for dx = dxproj . W_ih and dW = dpre^T . h of the GRU backward. */
int cruse_gemm_tn_tc(const float* const* A, const float* const* Bm, const float* const* bias, float* const* C,
                     int G, int M, int N, int K, long long lda, long long ldb, long long ldc,
                     int splitk, long long c_plane, void* stream);
/* The same GEMM with either operand read "MN-major", i.e. as it lies in HBM when the reduction index is the ROW of a row-major
 * matrix -- the natural layout of a weight gradient's factors (nn.GRU's dW_ih = dxproj^T . x and dW_hh = dpre^T . h_{t-1},
 * model/cruse_net.py:23-31 through autograd: both factors are [B*T, features]) and of W_ih in dx = dxproj . W_ih, so that no
 * transposed copy is made (tcgen05 MN-major shared-memory descriptors over TMA boxes of {32 features, 32 rows}):
 *   a_mn = 0: A_g[m,k] = A[m*lda + k]     a_mn = 1: A_g[m,k] = A[k*lda + m]
 *   b_mn = 0: B_g[n,k] = Bm[n*ldb + k]    b_mn = 1: B_g[n,k] = Bm[(k - b_kshift)*ldb + n], rows < 0 read as zero
 * (b_kshift = 1 pairs frame t with h_{t-1} when the rows are frames; the caller zeroes the t == 0 rows of A).
 * addend (table or NULL): C_g += addend_g, an [M,N] matrix with C's pitch ldc, 16-byte aligned (split-K: added to plane 0) --
 * where two gradient paths meet (out = g + skip4, model/cruse_net.py:160) the sum costs no extra pass.
 * Everything else as cruse_gemm_tn_tc, which is the a_mn = b_mn = 0, addend = NULL case. */
int cruse_gemm_tc(const float* const* A, const float* const* Bm, const float* const* bias, const float* const* addend,
                  float* const* C, int G, int M, int N, int K, long long lda, long long ldb, long long ldc,
                  int splitk, long long c_plane, int a_mn, int b_mn, int b_kshift, void* stream);
/* Exact-fp32 twins of the three tensor-core kernels of the GRU training path (same contracts and layouts; every product an
 * fp32 FMA on the CUDA cores, fixed summation order).  They make the whole training step runnable without tf32 operand rounding
 * (CRUSE_CONV=fp32 CRUSE_GRU_IH=fp32 CRUSE_GRU_SEQ=fp32), which is how the end-to-end gradients are compared with autograd of
 * the CPU oracle at fp32 tolerances (nn.GRU forward / backward, model/cruse_net.py:23-31,42-50).  Parity kernels, not the fast path.
 *  cruse_gru_seq_fwd_exact: as cruse_gru_seq_fwd_tc (gates [B,T,G,4,H] or NULL); ws: cruse_gru_exact_ws_bytes(G,H) bytes of scratch
 *  (a transposed copy of W_hh).  cruse_gru_seq_bwd_exact: as cruse_gru_seq_bwd_tc.  cruse_gemm_tn_fp32: as cruse_gemm_tn_tc. */
size_t cruse_gru_exact_ws_bytes(int G, int H);
int cruse_gru_seq_fwd_exact(const float* xproj, const float* const* w_hh, const float* const* b_hh, const float* h0,
                            float* y, float* hT, float* gates, void* ws, int B, int T, int G, int H, int y_fs, int y_gs,
                            void* stream);
int cruse_gru_seq_bwd_exact(const float* dy, const float* y, const float* gates, const float* h0,
                            const float* const* w_hh, float* dxproj, float* dpre, float* dh0, float* dbias_part,
                            int B, int T, int G, int H, int y_fs, int y_gs, void* stream);
int cruse_gemm_tn_fp32(const float* const* A, const float* const* Bm, const float* const* bias, float* const* C,
                       int G, int M, int N, int K, long long lda, long long ldb, long long ldc,
                       int splitk, long long c_plane, void* stream);
/* out[(g*Cn + c)*ldo + m] = in[m*ld + g*gs + c*cs] (m < M, c < Cn): puts the (b,t) index innermost for the weight-gradient
 * GEMMs.  shift_T > 0: row m = b*shift_T + t reads row m-1 (h_{t-1} from y) and t == 0 reads h0[g][b][c] (or 0 if NULL; Bn = B). */
int cruse_transpose_gcm(const float* in, const float* h0, float* out, long long M, int G, int Cn,
                        long long ld, long long gs, long long cs, int shift_T, int Bn, long long ldo, void* stream);

/* ---- f4: the step in front of the path, on the device (so that eight GPUs are not fed by a CPU dataloader).
 * cruse_feature_norm: the input feature norms of train_base/model/base_model.py:202-300 on a frame-major magnitude spectrogram
 *   x, y [B,T,F] (the reference's [B,1,F,T]): mode 0 offline_laplace (x / (mean + 1e-5)), 1 cumulative_laplace (x / (running mean over
 *   all bins of frames 0..t + eps)), 2 offline_gaussian ((x - mean) / (std + 1e-5), unbiased std), 3 cumulative_layer (running mean /
 *   variance, formula of :292 kept literally); eps = float32 epsilon (train_base/constant.py:8).
 * cruse_rir_conv: y[b, n] = sum_{k<R} rir[b*rir_stride + k] x[b, n-k], n < L = scipy.signal.fftconvolve(x, rir)[:L]
 *   (dataset/dataset.py:244-247); rir_stride 0 = one impulse response for the whole batch.  y must not alias x.
 * cruse_snr_mix: dataset/dataset.py:236-264 per utterance: clean / (max|clean| + eps), noise / (max|noise| + eps) scaled to
 *   snr_db[b], noisy = clean + noise; level_db (device array or NULL) = the output level in dB FS the reference draws at random
 *   (:262-264, where the reference file ends): noisy and clean are then scaled by 10^(dB/20) / (rms(noisy) + eps).
 *   clean_out may be NULL; ws: cruse_snr_mix_ws_bytes(B). */
int cruse_feature_norm(const float* x, float* y, int B, int T, int F, int mode, void* stream);
int cruse_rir_conv(const float* x, const float* rir, float* y, int B, int L, int R, long long rir_stride, void* stream);
size_t cruse_snr_mix_ws_bytes(int B);
int cruse_snr_mix(const float* clean, const float* noise, const float* snr_db, const float* level_db, float* noisy_out,
                  float* clean_out, void* ws, int B, int L, float eps, void* stream);
/* 16-bit PCM samples -> float32, out = in / 32768: the conversion soundfile / librosa apply when the reference reads a wav file
 * (dataset/dataset.py:20, train_base/acoustics/feature.py:110-114), on the device, so that the HOST-buffer entry points can take the
 * samples in the files' own format (half the bytes per step over PCIe).  16-byte aligned buffers. */
int cruse_pcm16_to_float(const short* in, float* out, long long n, void* stream);

/* =====================================================================================================
 * a10: the gradient exchange of data-parallel training (loss_func/distrib.py:100-116 sync_grad: all_reduce(SUM) of every
 * gradient, then division by the world size; train_base/trainer/base_trainer.py:31 DistributedDataParallel).
 * ONE ncclAllReduce, in place, on the flat fp32 gradient buffer of the step + one scaling pass, on the caller's stream.
 * NCCL is looked up at run time in the libnccl.so.2 of the process (torch's); without it these calls return an error
 * (cruse_nccl_available() == 0) and the library still loads.
 *   cruse_nccl_unique_id: rank 0 fills 128 bytes (ncclUniqueId) that the caller hands to every rank (any side channel);
 *   cruse_nccl_comm_init: collective over all ranks, the rank's device must be current; *comm is an ncclComm_t;
 *   cruse_flat_allreduce: buf[i] <- scale * sum_ranks buf[i], i < n (scale = 1 / world); buf 16-byte aligned;
 *   cruse_nccl_comm_destroy.
 * ===================================================================================================== */
int cruse_nccl_available(void);
int cruse_nccl_unique_id(void* id128);
int cruse_nccl_comm_init(void** comm, int nranks, int rank, const void* id128);
int cruse_flat_allreduce(void* comm, float* buf, long long n, float scale, void* stream);
int cruse_nccl_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* CRUSE_B200_H */

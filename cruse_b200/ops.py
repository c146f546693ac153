"""Thin tensor-level wrappers over the C ABI (include/cruse_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every function
checks its tensors, allocates outputs with the caching allocator, and launches the sm_100a
kernels on ``torch.cuda.current_stream()``.  Activations are frame-major ``[B, T, C, F]``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from ._lib import CplxLayout, check, lib

ACT = {"none": 0, "relu": 1, "prelu": 2, "sigmoid": 3}
PAD = {"reflect": 0, "constant": 1}


# kernels launched per C-ABI call (everything else launches exactly one)
LAUNCHES = {"cruse_wo_male_fwd_bwd": 2, "cruse_wo_male_masked_fwd": 2}


class Profile:
    """Optional per-call instrumentation used by bench.py: counts kernel launches and, when
    ``timing`` is on, brackets every C-ABI call with CUDA events on the launching stream."""

    def __init__(self, timing=False):
        self.timing = timing
        self.launches = 0
        self.events = []      # (name, tag, algorithmic bytes, flops, start_event, end_event)

    def rows(self):
        """[(name, tag, bytes, flops, ms)] in launch order."""
        torch.cuda.synchronize()
        return [(n, tag, by, fl, a.elapsed_time(b)) for n, tag, by, fl, a, b in self.events]


_profile = None


def set_profile(p):
    global _profile
    _profile = p


def _call(name, *args, meta=None):
    """meta = (tag, algorithmic_bytes, flops) of this launch -- the roofline numerators."""
    fn = getattr(lib(), name)
    prof = _profile
    if prof is None:
        check(fn(*args), name)
        return
    prof.launches += LAUNCHES.get(name, 1)
    if prof.timing:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(fn(*args), name)
        b.record()
        tag, by, fl = meta if meta is not None else ("", 0, 0)
        prof.events.append((name, tag, by, fl, a, b))
    else:
        check(fn(*args), name)


def _nb(*tensors):
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


def _p(t):
    return None if t is None else t.data_ptr()


def check_nonneg(rc, what):
    if rc < 0:
        check(rc, what)
    return rc


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name, ndim=None):
    if t is None:
        return
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{name}: expected a CUDA tensor (cruse_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")


def _ptr_table(tensors):
    if tensors is None:
        return None
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


# ------------------------------------------------------------------------------------------
# a1 / a6 / a7 : spectral front and back end
# ------------------------------------------------------------------------------------------
def stft_fwd(wav, window, n_fft, hop, pad_mode="reflect", mag_bins=0, mag_eps=1e-8):
    """wav [B,L] -> (spec [B,T,NF,2], mag [B,T,mag_bins] | None).  feature.py:10-30 + utils.py:400."""
    _req(wav, "wav", 2)
    _req(window, "window", 1)
    B, L = wav.shape
    if window.numel() != n_fft:
        raise RuntimeError(f"stft: window length {window.numel()} != n_fft {n_fft}")
    T = 1 + L // hop
    NF = n_fft // 2 + 1
    spec = torch.empty(B, T, NF, 2, device=wav.device, dtype=torch.float32)
    mag = torch.empty(B, T, mag_bins, device=wav.device, dtype=torch.float32) if mag_bins > 0 else None
    _call("cruse_stft_fwd", _p(wav), _p(window), _p(spec), _p(mag), B, L, n_fft, hop, T, PAD[pad_mode],
                               mag_bins, mag_eps, _stream(),
          meta=(f"stft n{n_fft} h{hop}", _nb(wav, spec, mag), int(B * T * 2.5 * n_fft * 9)))
    return spec, mag


def stft_fwd_into(wav, window, spec, n_fft, hop, pad_mode="reflect"):
    """STFT of wav [B,L] into a preallocated spec [B,T,NF,2] (no allocation: queued on a side stream)."""
    B, L = wav.shape
    T = 1 + L // hop
    if tuple(spec.shape) != (B, T, n_fft // 2 + 1, 2):
        raise RuntimeError(f"stft_fwd_into: spec shape {tuple(spec.shape)} != {(B, T, n_fft // 2 + 1, 2)}")
    _call("cruse_stft_fwd", _p(wav.contiguous()), _p(window), _p(spec), None, B, L, n_fft, hop, T, PAD[pad_mode], 0, 1e-8, _stream(),
          meta=(f"stft n{n_fft} h{hop}", _nb(wav, spec), int(B * T * 2.5 * n_fft * 9)))


def mask_istft_fwd(spec, mask, window, n_fft, hop, length, want_est=True, want_wav=True):
    """spec [B,T,NF,2] (* mask [B,T,Fm]) -> (est_spec | None, wav [B,length] | None).  utils.py:417-454."""
    _req(spec, "spec", 4)
    _req(mask, "mask")
    _req(window, "window", 1)
    B, T, NF, _ = spec.shape
    if NF != n_fft // 2 + 1:
        raise RuntimeError(f"mask_istft: spec has {NF} bins, n_fft={n_fft} needs {n_fft // 2 + 1}")
    mask_bins = 0
    if mask is not None:
        mask_bins = mask.shape[-1]
        if mask.numel() != B * T * mask_bins:
            raise RuntimeError(f"mask_istft: mask shape {tuple(mask.shape)} does not match spec {tuple(spec.shape)}")
    est = torch.empty_like(spec) if want_est else None
    wav = torch.empty(B, length, device=spec.device, dtype=torch.float32) if want_wav else None
    _call("cruse_mask_istft_fwd", _p(spec), _p(mask), _p(window), _p(est), _p(wav), B, length if want_wav else 0,
                                     n_fft, hop, T, mask_bins, _stream(),
          meta=(f"mask_istft n{n_fft} h{hop}", _nb(spec, mask, est, wav), int(B * T * 2.5 * n_fft * 9) if want_wav else 0))
    return est, wav


def mask_istft_chunk_frames(n_fft, hop):
    return check_nonneg(lib().cruse_mask_istft_chunk_frames(n_fft, hop), "cruse_mask_istft_chunk_frames")


def mask_istft_fwd_range(spec, mask, window, n_fft, hop, est, wav, c0, c1):
    """the CTAs [c0, c1) of mask_istft_fwd (each covers mask_istft_chunk_frames frames) into the full-size est / wav."""
    B, T, NF, _ = spec.shape
    mask_bins = mask.shape[-1]
    _call("cruse_mask_istft_fwd_range", _p(spec), _p(mask), _p(window), _p(est), _p(wav), B, wav.shape[-1], n_fft, hop, T, mask_bins,
          c0, c1, _stream(), meta=(f"mask_istft n{n_fft} h{hop} ctas[{c0},{c1})", 0, 0))


def istft_bwd(dwav, window, n_fft, hop):
    """gradient of wav = istft(spec) w.r.t. spec: dwav [B,L] -> dspec [B,T,NF,2] (an STFT of dwav / envelope, scaled)."""
    _req(dwav, "dwav", 2)
    _req(window, "window", 1)
    B, L = dwav.shape
    T = 1 + L // hop
    dspec = torch.empty(B, T, n_fft // 2 + 1, 2, device=dwav.device, dtype=torch.float32)
    ws = _ws(lib().cruse_istft_bwd_ws_bytes(B, L), dwav.device)
    _call("cruse_istft_bwd", _p(dwav), _p(window), _p(dspec), _p(ws), B, L, n_fft, hop, T, _stream(),
          meta=(f"istft_bwd n{n_fft} h{hop}", _nb(dwav, dspec) + 2 * ws.numel() * 4, int(B * T * 2.5 * n_fft * 9)))
    return dspec


def sisnr_fwd(est, ref, eps=1e-8):
    """-> (value 0-dim, ws): mean SI-SNR in dB of est vs ref wav [B,L] (loss.py:37-56); ws feeds sisnr_bwd."""
    _req(est, "est", 2)
    _req(ref, "ref", 2)
    if est.shape != ref.shape:
        raise RuntimeError(f"sisnr: shapes {tuple(est.shape)} vs {tuple(ref.shape)}")
    B, L = est.shape
    value = torch.empty((), device=est.device, dtype=torch.float32)
    ws = _ws(lib().cruse_sisnr_ws_bytes(B), est.device)
    _call("cruse_sisnr_fwd", _p(est), _p(ref), _p(value), _p(ws), B, L, float(eps), _stream(),
          meta=("sisnr", 2 * _nb(est, ref), 8 * B * L))
    return value, ws


def sisnr_bwd(est, ref, ws, gscale=None):
    """gscale * d value / d est (gscale: 0-dim device tensor or None)."""
    B, L = est.shape
    dest = torch.empty_like(est)
    _call("cruse_sisnr_bwd", _p(est), _p(ref), _p(ws), _p(gscale), _p(dest), B, L, _stream(), meta=("sisnr_bwd", _nb(est, ref, dest), 3 * B * L))
    return dest


def si_snr_zm_fwd(est, ref, eps=1e-8):
    """-> (loss 0-dim, ws): the zero-mean SI-SNR loss of train_base/loss.py:7-25 on wav [B,L]; ws feeds si_snr_zm_bwd."""
    _req(est, "est", 2)
    _req(ref, "ref", 2)
    if est.shape != ref.shape:
        raise RuntimeError(f"Dimension mismatch when calculate si_snr, {tuple(est.shape)} vs {tuple(ref.shape)}")
    B, L = est.shape
    value = torch.empty((), device=est.device, dtype=torch.float32)
    ws = _ws(lib().cruse_si_snr_zm_ws_bytes(B), est.device)
    _call("cruse_si_snr_zm_fwd", _p(est), _p(ref), _p(value), _p(ws), B, L, float(eps), _stream(), meta=("si_snr_zm", _nb(est, ref), 8 * B * L))
    return value, ws


def si_snr_zm_bwd(est, ref, ws, gscale=None):
    """gscale * d loss / d est (gscale: 0-dim device tensor or None)."""
    B, L = est.shape
    dest = torch.empty_like(est)
    _call("cruse_si_snr_zm_bwd", _p(est), _p(ref), _p(ws), _p(gscale), _p(dest), B, L, _stream(),
          meta=("si_snr_zm_bwd", _nb(est, ref, dest), 3 * B * L))
    return dest


def mask_bwd(dest, spec, mask_bins, gscale=None, mask=None):
    """dmask[b,t,f] = gscale * Re(conj(X) * dEst), f < mask_bins; with ``mask`` given the result is the gradient
    before the sigmoid (times mask*(1-mask))."""
    _req(dest, "dest", 4)
    _req(spec, "spec", 4)
    _req(mask, "mask")
    B, T, NF, _ = spec.shape
    dmask = torch.empty(B, T, mask_bins, device=spec.device, dtype=torch.float32)
    _call("cruse_mask_bwd", _p(dest), _p(spec), _p(gscale), _p(mask), _p(dmask), B, T, NF, mask_bins, _stream(),
          meta=("mask_bwd", _nb(dmask, mask) + 2 * 8 * B * T * mask_bins, 6 * B * T * mask_bins))
    return dmask


# ------------------------------------------------------------------------------------------
# a2 / a3 / a5 : conv stages
# ------------------------------------------------------------------------------------------
def conv_nparts(B, T):
    return lib().cruse_conv_nparts(B, T)


def conv_fwd(x, w, bias, scale, shift, alpha, act, kt, fstride, want_stats=False, hist=None, exact=False):
    """x [B,T,Cin,Fin] -> out [B,T,Cout,Fout] (+ per-CTA BN partials [nparts, 2*Cout]).
    exact=True pins this call to the exact-fp32 CUDA-core kernel whatever the process-wide conv mode is."""
    if exact and get_conv_mode() == "tf32":
        set_conv_mode("fp32")
        try:
            return conv_fwd(x, w, bias, scale, shift, alpha, act, kt, fstride, want_stats, hist)
        finally:
            set_conv_mode("tf32")
    _req(x, "x", 4)
    _req(w, "w", 4)
    for n, t in (("bias", bias), ("scale", scale), ("shift", shift), ("alpha", alpha)):
        _req(t, n)
    B, T, Cin, Fin = x.shape
    Cout = w.shape[0]
    if tuple(w.shape) != (Cout, Cin, kt, 3):
        raise RuntimeError(f"conv_fwd: weight shape {tuple(w.shape)} != ({Cout},{Cin},{kt},3)")
    Fout = (Fin + 2 - 3) // fstride + 1
    out = torch.empty(B, T, Cout, Fout, device=x.device, dtype=torch.float32)
    stats = torch.empty(conv_nparts(B, T), 2 * Cout, device=x.device, dtype=torch.float32) if want_stats else None
    _req(hist, "hist")
    if hist is not None and (kt != 2 or hist.numel() != B * Cin * Fin):
        raise RuntimeError(f"conv_fwd: hist must be [B,Cin,Fin] and kt == 2, got {tuple(hist.shape)}")
    _call("cruse_conv_fwd", _p(x), _p(hist), _p(w), _p(bias), _p(scale), _p(shift), _p(alpha), ACT[act], _p(out), _p(stats),
                               B, T, Cin, Fin, Cout, Fout, kt, fstride, _stream(),
          meta=(f"conv{kt}x3 {Cin}->{Cout} F{Fin}->{Fout}", _nb(x, out, w, bias), 2 * B * T * Cout * Fout * Cin * kt * 3))
    return (out, stats) if want_stats else out


def conv_fwd_tm(x, w, bias, scale, shift, alpha, act, kt, fstride, B, T, in_tm, out_tm):
    """eval-mode tensor-core stage whose input and / or output frame records are time-major: x [B*T, Cin, Fin] records
    ordered (t, b) when in_tm else (b, t); the result is [T, B, Cout, Fout] when out_tm else [B, T, Cout, Fout]."""
    _req(x, "x")
    _req(w, "w", 4)
    Cout, Cin = w.shape[0], w.shape[1]
    Fin = x.shape[-1]
    if x.numel() != B * T * Cin * Fin:
        raise RuntimeError(f"conv_fwd_tm: x has {x.numel()} elements, expected {B}*{T}*{Cin}*{Fin}")
    Fout = (Fin + 2 - 3) // fstride + 1
    out = torch.empty((T, B, Cout, Fout) if out_tm else (B, T, Cout, Fout), device=x.device, dtype=torch.float32)
    _call("cruse_conv_fwd_tm", _p(x), _p(w), _p(bias), _p(scale), _p(shift), _p(alpha), ACT[act], _p(out), B, T, Cin, Fin, Cout, Fout,
          kt, fstride, 1 if in_tm else 0, 1 if out_tm else 0, _stream(),
          meta=(f"conv{kt}x3 {Cin}->{Cout} F{Fin}->{Fout}{' tm' if (in_tm or out_tm) else ''}", _nb(x, out, w, bias),
                2 * B * T * Cout * Fout * Cin * kt * 3))
    return out


# fuse the skip conv of a stage's input into the stage (inference, tf32 conv mode): CRUSE_FUSE_SKIPS=0 keeps them separate
FUSE_SKIPS = os.environ.get("CRUSE_FUSE_SKIPS", "1") != "0"


def conv_skip_fwd(x, w, bias, scale, shift, alpha, act, w_skip, out=None, out_skip=None, out_tm=False, t0=0, t1=0):
    """eval-mode encoder stage (2,3)/stride (1,2) + folded BN + act WITH the (1,3) skip conv of its input fused in:
    x [B,T,Cin,Fin] -> (out [B,T,Cout,Fin/2] ([T,B,..] with out_tm), out_skip [B,T,Cin,Fin]); one pass over x."""
    _req(x, "x", 4)
    _req(w, "w", 4)
    _req(w_skip, "w_skip", 4)
    B, T, Cin, Fin = x.shape
    Cout = w.shape[0]
    if tuple(w.shape) != (Cout, Cin, 2, 3) or tuple(w_skip.shape) != (Cin, Cin, 1, 3):
        raise RuntimeError(f"conv_skip_fwd: weight shapes {tuple(w.shape)} / {tuple(w_skip.shape)} do not match x {tuple(x.shape)}")
    Fout = (Fin + 2 - 3) // 2 + 1
    if out is None:
        out = torch.empty((T, B, Cout, Fout) if out_tm else (B, T, Cout, Fout), device=x.device, dtype=torch.float32)
    if out_skip is None:
        out_skip = torch.empty(B, T, Cin, Fin, device=x.device, dtype=torch.float32)
    if out.numel() != B * T * Cout * Fout or out_skip.numel() != B * T * Cin * Fin:
        raise RuntimeError(f"conv_skip_fwd: out / out_skip sizes {out.numel()} / {out_skip.numel()} do not match x {tuple(x.shape)}")
    _call("cruse_conv_skip_fwd", _p(x), _p(w), _p(bias), _p(scale), _p(shift), _p(alpha), ACT[act], _p(w_skip), _p(out), _p(out_skip),
          B, T, Cin, Fin, Cout, Fout, 1 if out_tm else 0, t0, t1, _stream(),
          meta=(f"conv2x3 {Cin}->{Cout} F{Fin}->{Fout} + skip1x3 {Cin}->{Cin}", _nb(x, out, out_skip, w, w_skip),
                2 * B * T * (Cout * Fout * Cin * 6 + Cin * Fin * Cin * 3)))
    return out, out_skip


def conv_fwd_range(x, w, bias, scale, shift, alpha, act, kt, fstride, B, T, out, t0, t1, in_tm=False, out_tm=False):
    """eval-mode stage for the output frames [t0, t1) only, written into the full-size ``out`` ([B,T,Cout,Fout], or
    [T,B,Cout,Fout] with ``out_tm``); x is the full-size input of the stage (frames < t1 must be valid)."""
    Cout, Cin = w.shape[0], w.shape[1]
    Fin = x.shape[-1]
    Fout = (Fin + 2 - 3) // fstride + 1
    if x.numel() != B * T * Cin * Fin or out.numel() != B * T * Cout * Fout:
        raise RuntimeError(f"conv_fwd_range: x / out sizes {x.numel()} / {out.numel()} do not match B={B} T={T} {Cin}x{Fin} -> {Cout}x{Fout}")
    _call("cruse_conv_fwd_range", _p(x), _p(w), _p(bias), _p(scale), _p(shift), _p(alpha), ACT[act], _p(out), B, T, Cin, Fin, Cout, Fout,
          kt, fstride, 1 if in_tm else 0, 1 if out_tm else 0, t0, t1, _stream(),
          meta=(f"conv{kt}x3 {Cin}->{Cout} F{Fin}->{Fout} [{t0},{t1})", 4 * B * (t1 - t0) * (Cin * Fin + Cout * Fout),
                2 * B * (t1 - t0) * Cout * Fout * Cin * kt * 3))


def convT_fwd_range(x, w, bias, scale, shift, alpha, act, skip, out, t0, t1):
    """eval-mode decoder stage for the frames [t0, t1) only: x [B,T,Cin,Fin], skip / out [B,T,Cout,Fout] full size."""
    B, T, Cin, Fin = x.shape
    Cout, Fout = out.shape[2], out.shape[3]
    if tuple(w.shape) != (Cin, Cout, 1, 3) or tuple(out.shape[:2]) != (B, T):
        raise RuntimeError(f"convT_fwd_range: weight {tuple(w.shape)} / out {tuple(out.shape)} do not match x {tuple(x.shape)}")
    if skip is not None and tuple(skip.shape) != tuple(out.shape):
        raise RuntimeError(f"convT_fwd_range: skip shape {tuple(skip.shape)} != {tuple(out.shape)}")
    _call("cruse_convT_fwd_range", _p(x), _p(w), _p(bias), _p(scale), _p(shift), _p(alpha), ACT[act], _p(skip), _p(out),
          B, T, Cin, Fin, Cout, Fout, t0, t1, _stream(),
          meta=(f"convT1x3 {Cin}->{Cout} F{Fin}->{Fout} [{t0},{t1})", 4 * B * (t1 - t0) * (Cin * Fin + (2 if skip is not None else 1) * Cout * Fout),
                2 * B * (t1 - t0) * Cin * Fin * Cout * 3))


def layernorm_fwd_range(x, gamma, beta, eps, residual, y, t0, t1):
    """LayerNorm (+ residual) of the frames [t0, t1) of frame-major x / residual / y [B,T,D]."""
    B, T, D = x.shape
    _call("cruse_layernorm_fwd_range", _p(x), _p(gamma), _p(beta), float(eps), _p(residual), _p(y), B, T, D, t0, t1, _stream(),
          meta=(f"layernorm D{D} [{t0},{t1})", 4 * B * (t1 - t0) * D * (3 if residual is not None else 2), 8 * B * (t1 - t0) * D))


# LayerNorm 2 + the four decoder stages of the 256-bin pyramid as one launch per frame range (inference, tf32 conv mode);
# CRUSE_FUSE_DECODER=0 keeps the five per-stage launches
FUSE_DECODER = os.environ.get("CRUSE_FUSE_DECODER", "1") != "0"
# ... and the frames' shares of the wo_male loss in the same launch (CRUSE_FUSE_LOSS=0: separate loss launches per range)
FUSE_LOSS = os.environ.get("CRUSE_FUSE_LOSS", "1") != "0"
# ... and skip convs 4 and 3 from the encoder outputs (CRUSE_FUSE_SKIP34=0: two separate skip-conv launches beside the recurrences)
FUSE_SKIP34 = os.environ.get("CRUSE_FUSE_SKIP34", "1") != "0"


def decoder_fused_prep(ws, biases, scales, shifts, alphas, act, wskip4=None, wskip3=None):
    """the constants of the fused decoder as its shared-memory image (once per forward pass, after the BatchNorm fold): ``ws`` /
    ``biases`` = conv4_t .. conv1_t, ``scales`` / ``shifts`` / ``alphas`` = stages 4..2 -> image tensor for decoder_fused_range.
    ``wskip4`` / ``wskip3`` (skip_connect_4 / _3 weights): the image for ``skip_convs=True`` launches."""
    if (wskip4 is None) != (wskip3 is None):
        raise RuntimeError("decoder_fused_prep: wskip4 and wskip3 come together or not at all")
    if wskip4 is not None and (tuple(wskip4.shape) != (64, 64, 1, 3) or tuple(wskip3.shape) != (32, 32, 1, 3)):
        raise RuntimeError(f"decoder_fused_prep: skip conv weights {tuple(wskip4.shape)} / {tuple(wskip3.shape)} do not match the 256-bin pyramid")
    want = [(64, 32), (32, 16), (16, 8), (8, 1)]
    for k, (w, (ci, co)) in enumerate(zip(ws, want)):
        if tuple(w.shape) != (ci, co, 1, 3):
            raise RuntimeError(f"decoder_fused_prep: stage {4 - k}: weight {tuple(w.shape)} does not match the 256-bin pyramid")
    for t in [*ws, *[b for b in biases if b is not None], *scales, *shifts, *[a for a in (alphas or []) if a is not None]]:
        _req(t, "decoder_fused_prep tensor")
    for t in (wskip4, wskip3):
        _req(t, "decoder_fused_prep skip weight")
    image = torch.empty(int(lib().cruse_decoder_fused_image_floats(int(wskip4 is not None))), device=ws[0].device, dtype=torch.float32)
    tb = (C.c_void_p * 4)(*[b.data_ptr() if b is not None else None for b in biases])
    ta = (C.c_void_p * 3)(*[a.data_ptr() if a is not None else None for a in alphas]) if alphas is not None else None
    _call("cruse_decoder_fused_prep", _ptr_table(ws), tb, _ptr_table(scales), _ptr_table(shifts), ta, ACT[act], _p(wskip4), _p(wskip3),
          _p(image), _stream())
    return image


def decoder_fused_range(y2, ln_gamma, ln_beta, eps, skips, image, mask, t0, t1, max_ctas=0, loss=None, skip_convs=False):
    """model/cruse_net.py:51,160-164 for the frames [t0,t1): y2 [B,T,1024] -> mask [B,T,256] (in place).  ``skips`` = (skip4, skip3,
    skip2, skip1) [B,T,C,F]; ``image`` from decoder_fused_prep.  ``loss`` = (S, layout_S, X, layout_X, rows[B*T]): also leaves the
    frames' shares of wo_male on est = mask * X in ``rows`` (wo_male_finish_rows sums them).  ``skip_convs``: skips[0] / skips[1] are
    the encoder outputs e4 (TIME-MAJOR [T,B,64,16]) / e3 ([B,T,32,32]) and skip convs 4 / 3 run inside the launch (image prepared
    with their weights)."""
    B, T, D = y2.shape
    want = [(64, 16), (32, 32), (16, 64), (8, 128)]
    if D != 1024 or tuple(mask.shape[:2]) != (B, T) or mask[0, 0].numel() != 256:
        raise RuntimeError(f"decoder_fused_range: y2 {tuple(y2.shape)} / mask {tuple(mask.shape)}: the fused decoder is built for the 256-bin pyramid")
    for k, (sk, (c, f)) in enumerate(zip(skips, want)):
        if tuple(sk.shape) != ((T, B, c, f) if (skip_convs and k == 0) else (B, T, c, f)):
            raise RuntimeError(f"decoder_fused_range: stage {4 - k}: skip {tuple(sk.shape)} does not match the pyramid")
    if image.numel() != int(lib().cruse_decoder_fused_image_floats(int(bool(skip_convs)))):
        raise RuntimeError("decoder_fused_range: the image was prepared for the other skip_convs mode")
    for t in [y2, ln_gamma, ln_beta, mask, image, *skips]:
        _req(t, "decoder_fused_range tensor")
    frames = B * (t1 - t0)
    if loss is not None:
        S, lay_s, X, lay_x, rows = loss
        for t in (S, X, rows):
            _req(t, "decoder_fused_range loss tensor")
        if rows.numel() < B * T:
            raise RuntimeError(f"decoder_fused_range: loss rows hold {rows.numel()} values, need B*T = {B * T}")
        largs = (_p(S), lay_s, _p(X), lay_x, _p(rows))
    else:
        zero = CplxLayout(0, 0, 0, 0)
        largs = (None, zero, None, zero, None)
    _call("cruse_decoder_fused_range", _p(y2), _p(ln_gamma), _p(ln_beta), float(eps), _ptr_table(skips), int(bool(skip_convs)), _p(image), _p(mask), *largs,
          B, T, t0, t1, int(max_ctas), _stream(),
          meta=(f"decoder_fused{'+skip34' if skip_convs else ''}{'+loss' if loss is not None else ''} [{t0},{t1})",
                4 * frames * (5 * 1024 + 256 + (4 * 256 if loss is not None else 0)), 2 * frames * (175104 + (294912 if skip_convs else 0))))


def wo_male_finish_rows(rows, B, T, F):
    """sum of the per-frame loss shares of decoder_fused_range / (B*T*F) -> 0-dim loss"""
    loss = torch.empty((), device=rows.device, dtype=torch.float32)
    _call("cruse_wo_male_finish_rows", _p(rows), B, T, F, _p(loss), _stream())
    return loss


def convT_fwd(x, w, bias, scale, shift, alpha, act, skip, Fout, want_stats=False):
    """x [B,T,Cin,Fin] -> out [B,T,Cout,Fout];  w [Cin,Cout,1,3] (ConvTranspose2d layout)."""
    _req(x, "x", 4)
    _req(w, "w", 4)
    for n, t in (("bias", bias), ("scale", scale), ("shift", shift), ("alpha", alpha), ("skip", skip)):
        _req(t, n)
    B, T, Cin, Fin = x.shape
    Cout = w.shape[1]
    if tuple(w.shape) != (Cin, Cout, 1, 3):
        raise RuntimeError(f"convT_fwd: weight shape {tuple(w.shape)} != ({Cin},{Cout},1,3)")
    if skip is not None and tuple(skip.shape) != (B, T, Cout, Fout):
        raise RuntimeError(f"convT_fwd: skip shape {tuple(skip.shape)} != {(B, T, Cout, Fout)}")
    out = torch.empty(B, T, Cout, Fout, device=x.device, dtype=torch.float32)
    stats = torch.empty(conv_nparts(B, T), 2 * Cout, device=x.device, dtype=torch.float32) if want_stats else None
    _call("cruse_convT_fwd", _p(x), _p(w), _p(bias), _p(scale), _p(shift), _p(alpha), ACT[act], _p(skip), _p(out),
                                _p(stats), B, T, Cin, Fin, Cout, Fout, _stream(),
          meta=(f"convT1x3 {Cin}->{Cout} F{Fin}->{Fout}", _nb(x, out, skip, w, bias), 2 * B * T * Cin * Fin * Cout * 3))
    return (out, stats) if want_stats else out


def bn_fold(bn):
    """eval-mode BatchNorm2d -> per-channel (scale, shift)."""
    Cn = bn.num_features
    scale = torch.empty(Cn, device=bn.running_mean.device, dtype=torch.float32)
    shift = torch.empty_like(scale)
    _call("cruse_bn_fold", _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var), float(bn.eps),
                              _p(scale), _p(shift), Cn, _stream())
    return scale, shift


def bn_fold_many(bns):
    """eval-mode fold of up to 8 BatchNorm2d layers in one launch -> [(scale, shift), ...]."""
    n = len(bns)
    dev = bns[0].running_mean.device
    total = sum(b.num_features for b in bns)
    buf = torch.empty(2, total, device=dev, dtype=torch.float32)
    out, off = [], 0
    for b in bns:
        out.append((buf[0, off:off + b.num_features], buf[1, off:off + b.num_features]))
        off += b.num_features
    tab = lambda ts: (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in ts])
    eps = (C.c_float * n)(*[float(b.eps) for b in bns])
    cs = (C.c_int * n)(*[b.num_features for b in bns])
    _call("cruse_bn_fold_many", tab([b.weight for b in bns]), tab([b.bias for b in bns]), tab([b.running_mean for b in bns]),
          tab([b.running_var for b in bns]), C.cast(eps, C.c_void_p), tab([o[0] for o in out]), tab([o[1] for o in out]),
          C.cast(cs, C.c_void_p), n, _stream())
    return out


def bn_finalize(stats, count, bn, update_running=True, counters=None):
    """train-mode BatchNorm2d: per-CTA partials -> (scale, shift, save_mean, save_invstd); updates running stats.
    ``counters`` (list): ``num_batches_tracked`` is appended instead of incremented -- the caller bumps all of a forward pass's
    counters with ONE launch (``bump_counters``) instead of one single-element kernel per stage in the middle of the chain."""
    _req(stats, "stats", 2)
    nparts, C2 = stats.shape
    Cn = C2 // 2
    dev = stats.device
    scale = torch.empty(Cn, device=dev, dtype=torch.float32)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale)
    invstd = torch.empty_like(scale)
    upd = update_running and bn.track_running_stats and bn.running_mean is not None
    mom = bn.momentum if bn.momentum is not None else 0.1
    _call("cruse_bn_finalize", _p(stats), nparts, Cn, float(count), _p(bn.weight), _p(bn.bias), float(bn.eps),
                                  float(mom), _p(bn.running_mean) if upd else None,
                                  _p(bn.running_var) if upd else None, _p(scale), _p(shift), _p(mean), _p(invstd),
                                  _stream())
    if upd and bn.num_batches_tracked is not None:
        if counters is not None:
            counters.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)
    return scale, shift, mean, invstd


def bump_counters(counters):
    if counters:
        torch._foreach_add_(counters, 1)


def bn_act_fwd(z, scale, shift, alpha, act, skip=None):
    _req(z, "z", 4)
    B, T, Cn, F = z.shape
    y = torch.empty_like(z)
    _call("cruse_bn_act_fwd", _p(z), _p(scale), _p(shift), _p(alpha), ACT[act], _p(skip), _p(y), B * T, Cn, F,
          _stream(), meta=(f"bn_act C{Cn} F{F}", _nb(z, y, skip), 2 * z.numel()))
    return y


# ------------------------------------------------------------------------------------------
# a4 : grouped GRU + LayerNorm
# ------------------------------------------------------------------------------------------
# "tf32": tcgen05 tensor cores (default); "fp32": exact-fp32 CUDA-core GEMM (parity debugging)
GRU_IH_MODE = os.environ.get("CRUSE_GRU_IH", "tf32")


def gru_ih_gemm(x, w_ih, b_ih, b_hh, mode=None):
    """x [M, G*H] -> xproj [M, G, 3H] = x_g . w_ih[g]^T + b_ih[g] (+ b_hh[g] on the r,z rows)."""
    mode = mode or GRU_IH_MODE
    _req(x, "x", 2)
    G = len(w_ih)
    H = w_ih[0].shape[1]
    M = x.shape[0]
    if x.shape[1] != G * H:
        raise RuntimeError(f"gru_ih_gemm: x has {x.shape[1]} features, expected {G}*{H}")
    for t in list(w_ih) + list(b_ih or []) + list(b_hh or []):
        _req(t, "gru weight")
    xproj = torch.empty(M, G, 3 * H, device=x.device, dtype=torch.float32)
    tw, tbi, tbh = _ptr_table(w_ih), _ptr_table(b_ih), _ptr_table(b_hh)
    if mode not in ("tf32", "fp32"):
        raise RuntimeError(f"gru_ih_gemm: unknown mode {mode!r}")
    _call("cruse_gru_ih_gemm_tc" if mode == "tf32" else "cruse_gru_ih_gemm", _p(x), tw, tbi, tbh, _p(xproj), M, G, H,
          _stream(), meta=(f"gru_ih[{mode}] G{G} H{H}", _nb(x, xproj, *w_ih), 2 * M * G * H * 3 * H))
    return xproj


# "tf32": tcgen05 recurrence (default); "fp32": exact-fp32 register-resident CUDA-core recurrence
GRU_SEQ_MODE = os.environ.get("CRUSE_GRU_SEQ", "tf32")


def gru_seq_fwd(xproj, w_hh, b_hh, B, T, interleave, h0=None, want_hT=False, mode=None, want_gates=False):
    """recurrence over T; y [B,T,G*H] with y[..., j*G+g] (interleave, cruse_net.py:43-45) or y[..., g*H+j] (cat).
    want_gates: also returns gates [B,T,G,4,H] = r, z, n, W_hn.h+b_hn (saved for backward)."""
    mode = mode or GRU_SEQ_MODE
    _req(xproj, "xproj", 3)
    _req(h0, "h0")
    G = len(w_hh)
    H = w_hh[0].shape[1]
    if tuple(xproj.shape) != (B * T, G, 3 * H):
        raise RuntimeError(f"gru_seq_fwd: xproj shape {tuple(xproj.shape)} != {(B * T, G, 3 * H)}")
    y = torch.empty(B, T, G * H, device=xproj.device, dtype=torch.float32)
    hT = torch.empty(G, B, H, device=xproj.device, dtype=torch.float32) if want_hT else None
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    tw, tb = _ptr_table(w_hh), _ptr_table(b_hh)
    gates = None
    if mode == "tf32":
        if want_gates:
            gates = torch.empty(B, T, G, 4, H, device=xproj.device, dtype=torch.float32)
        _call("cruse_gru_seq_fwd_tc", _p(xproj), tw, tb, _p(h0), _p(y), _p(hT), _p(gates), B, T, G, H, y_fs, y_gs,
              _stream(), meta=(f"gru_seq[tf32] G{G} H{H} T{T}", _nb(xproj, y, gates, *w_hh), 2 * B * T * G * H * 3 * H))
    elif mode == "fp32" and want_gates:
        # exact-fp32 training forward (saves the gates): the parity twin of the tensor-core kernel (csrc/gru_exact.cu)
        gates = torch.empty(B, T, G, 4, H, device=xproj.device, dtype=torch.float32)
        ws = _ws(lib().cruse_gru_exact_ws_bytes(G, H), xproj.device)
        _call("cruse_gru_seq_fwd_exact", _p(xproj), tw, tb, _p(h0), _p(y), _p(hT), _p(gates), _p(ws), B, T, G, H, y_fs, y_gs,
              _stream(), meta=(f"gru_seq[fp32 exact] G{G} H{H} T{T}", _nb(xproj, y, gates, *w_hh), 2 * B * T * G * H * 3 * H))
    elif mode == "fp32":
        _call("cruse_gru_seq_fwd", _p(xproj), tw, tb, _p(h0), _p(y), _p(hT), B, T, G, H, y_fs, y_gs, _stream(),
              meta=(f"gru_seq[fp32] G{G} H{H} T{T}", _nb(xproj, y, *w_hh), 2 * B * T * G * H * 3 * H))
    else:
        raise RuntimeError(f"gru_seq_fwd: unknown mode {mode!r}")
    if want_gates:
        return (y, hT, gates) if want_hT else (y, gates)
    return (y, hT) if want_hT else y


def gru_step(xproj, w_hh, b_hh, h_prev, interleave):
    """ONE streaming step for B concurrent utterances (BASELINE cfg-5): xproj [B,G,3H], h_prev [G,B,H] | None ->
    (y [B,1,G*H] in the layer's output order, h_new [G,B,H]).  The hidden half is one tcgen05 GEMM per group."""
    _req(xproj, "xproj", 3)
    _req(h_prev, "h_prev", 3)
    B, G, H3 = xproj.shape
    H = H3 // 3
    if len(w_hh) != G or tuple(w_hh[0].shape) != (3 * H, H):
        raise RuntimeError(f"gru_step: w_hh {len(w_hh)} x {tuple(w_hh[0].shape)} does not match xproj {tuple(xproj.shape)}")
    if h_prev is not None and tuple(h_prev.shape) != (G, B, H):
        raise RuntimeError(f"gru_step: h_prev shape {tuple(h_prev.shape)} != {(G, B, H)}")
    dev = xproj.device
    y = torch.empty(B, 1, G * H, device=dev, dtype=torch.float32)
    h_new = torch.empty(G, B, H, device=dev, dtype=torch.float32)
    ws = _ws(lib().cruse_gru_step_ws_bytes(B, G, H), dev)
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    _call("cruse_gru_step", _p(xproj), _ptr_table(w_hh), _ptr_table(b_hh), _p(h_prev), _p(h_new), _p(y), _p(ws), B, G, H,
          y_fs, y_gs, _stream(),
          meta=(f"gru_step[tf32] G{G} H{H} B{B}", _nb(xproj, h_prev, h_new, y, *w_hh) + 2 * ws.numel() * 4, 2 * B * G * H * 3 * H))
    return y, h_new


def gru_ih_gemm_tm(x, w_ih, b_ih, b_hh, B, T):
    """x [B*T, G*H] in frame order -> xproj [T, B, G, 3H] TIME-MAJOR (tcgen05, tf32): a chunk of frames of the
    result is one contiguous row range, which is what the two-layer wavefront needs."""
    _req(x, "x", 2)
    G = len(w_ih)
    H = w_ih[0].shape[1]
    if tuple(x.shape) != (B * T, G * H):
        raise RuntimeError(f"gru_ih_gemm_tm: x shape {tuple(x.shape)} != {(B * T, G * H)}")
    xproj = torch.empty(T, B, G, 3 * H, device=x.device, dtype=torch.float32)
    _call("cruse_gru_ih_gemm_tm_tc", _p(x), _ptr_table(w_ih), _ptr_table(b_ih), _ptr_table(b_hh), _p(xproj), B, T, G, H,
          _stream(), meta=(f"gru_ih[tf32,tm] G{G} H{H}", _nb(x, xproj, *w_ih), 2 * B * T * G * H * 3 * H))
    return xproj


def gru_seq_chunk(xproj_tm, w_hh, b_hh, h0, y, hT, t0, t1, interleave, y_time_major):
    """frames [t0, t1) of the recurrence: xproj_tm [T,B,G,3H] time-major; y either time-major [T,B,G*H] or frame
    order [B,T,G*H]; h0 / hT [G,B,H] carry the state from / to the neighbouring chunks (h0 None = zero state)."""
    T, B, G, H3 = xproj_tm.shape
    H = H3 // 3
    if y_time_major:
        y_bs, y_ts, yoff = 1, B, t0 * B * G * H
    else:
        y_bs, y_ts, yoff = T, 1, t0 * G * H
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    _call("cruse_gru_seq_chunk_tc", xproj_tm.data_ptr() + 4 * t0 * B * G * H3, _ptr_table(w_hh), _ptr_table(b_hh), _p(h0),
          y.data_ptr() + 4 * yoff, _p(hT), B, t1 - t0, G, H, y_fs, y_gs, 1, B, y_bs, y_ts, _stream(),
          meta=(f"gru_seq_chunk[tf32] G{G} H{H} T{t1 - t0}", 4 * (t1 - t0) * B * G * (H3 + H), 2 * B * (t1 - t0) * G * H * 3 * H))


# two-layer wavefront of the GGRU bottleneck (cruse_net.GGRU._wavefront); CRUSE_GRU_WAVEFRONT=0 runs the layers back to back
GRU_WAVEFRONT = os.environ.get("CRUSE_GRU_WAVEFRONT", "1") != "0"


_max_clusters_cache = {}


def gru_seq_max_clusters(H):
    """co-resident clusters of the recurrence kernel on the current device (a device property: cached)."""
    key = (torch.cuda.current_device(), H)
    if key not in _max_clusters_cache:
        n = lib().cruse_gru_seq_tc_max_clusters(H)
        if n < 0:
            check(n, "cruse_gru_seq_tc_max_clusters")
        _max_clusters_cache[key] = n
    return _max_clusters_cache[key]


# how the wavefront synchronises its chunks: "flags" = one launch per layer + device-side chunk flags (default);
# "relaunch" = one recurrence launch per chunk, CUDA events between the streams (no spinning kernels: use this under
# tools that serialise kernels, e.g. ncu / compute-sanitizer)
def _under_kernel_serialising_tool():
    # Nsight Compute (ncu) replays / serialises kernels; its injection sets these variables in the profiled process
    return any(k in os.environ for k in ("NV_TPS_LAUNCH_TOKEN", "NV_NSIGHT_INJECTION_PORT_BASE", "NV_COMPUTE_PROFILER_PERFWORKS_DIR"))


GRU_WAVEFRONT_MODE = os.environ.get("CRUSE_GRU_WAVEFRONT_MODE", "relaunch" if _under_kernel_serialising_tool() else "flags")


def gru_seq_flagged(xproj_tm, w_hh, b_hh, y, interleave, y_time_major, bounds, wait_flags, wait_target, done_flags, err):
    """all T frames of one layer in ONE launch; chunk k starts once wait_flags[k] >= wait_target and bumps done_flags[k]
    when it is stored (include/cruse_b200.h: cruse_gru_seq_flagged_tc)."""
    T, B, G, H3 = xproj_tm.shape
    H = H3 // 3
    y_bs, y_ts = (1, B) if y_time_major else (T, 1)
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    nch = len(bounds) - 1
    barr = (C.c_int * (nch + 1))(*bounds)
    _call("cruse_gru_seq_flagged_tc", _p(xproj_tm), _ptr_table(w_hh), _ptr_table(b_hh), _p(y), B, T, G, H, y_fs, y_gs, 1, B, y_bs, y_ts,
          C.cast(barr, C.c_void_p), nch, _p(wait_flags), int(wait_target), _p(done_flags), _p(err), _stream(),
          meta=(f"gru_seq_flagged[tf32] G{G} H{H} T{T}", 4 * T * B * G * (H3 + H), 2 * B * T * G * H * 3 * H))


def flag_set(flag, value=1):
    _call("cruse_flag_set", _p(flag), int(value), _stream())


def flag_wait(flag, target, err):
    _call("cruse_flag_wait", _p(flag), int(target), _p(err), _stream())


class WavefrontTimeout(RuntimeError):
    """a bounded device-side spin of the flag-synchronised GRU wavefront gave up (2 s): the step's outputs are NaN"""


def poison_on_error(err, tensors):
    """device side of the loud failure: fill up to 4 fp32 tensors with NaN if the wavefront error flag ``err`` is set
    (one tiny launch on the current stream; include/cruse_b200.h: cruse_poison_on_error)."""
    tensors = [t for t in tensors if t is not None]
    counts = (C.c_longlong * len(tensors))(*[t.numel() for t in tensors])
    _call("cruse_poison_on_error", _p(err), _ptr_table(tensors), C.cast(counts, C.c_void_p), len(tensors), _stream())


def raise_if_wavefront_failed(err_flags, what="cruse_b200"):
    """host side: read the error flag(s) (synchronises with the device) and raise; later calls fall back to the relaunch
    wavefront, which has no spinning kernels.  Flag mode needs a device this process has to itself: kernels of another
    process / MPS client / user stream that hold SMs can keep a producer off the GPU while its consumer spins."""
    global GRU_WAVEFRONT_MODE
    flags = [f for f in (err_flags if isinstance(err_flags, (list, tuple)) else [err_flags]) if f is not None]
    if any(int(f.item()) != 0 for f in flags):
        GRU_WAVEFRONT_MODE = "relaunch"
        raise WavefrontTimeout(
            f"{what}: the flag-synchronised GRU wavefront timed out (a producer kernel was kept off the GPU for > 2 s while its "
            "consumer was spinning -- is another process or stream using this device?).  The outputs of that step were set to NaN. "
            "Falling back to CRUSE_GRU_WAVEFRONT_MODE=relaunch for the following calls (captured graphs must be re-captured).")


def layernorm_fwd_into(x, gamma, beta, eps, y):
    """LayerNorm over the rows of a contiguous slice, written into a preallocated slice (no allocation: runs on side streams)."""
    D = x.shape[-1]
    rows = x.numel() // D
    _call("cruse_layernorm_fwd", _p(x), _p(gamma), _p(beta), float(eps), None, _p(y), None, None, rows, D, _stream(),
          meta=(f"layernorm D{D}", _nb(x, y), 8 * x.numel()))


def layernorm_interleave_fwd_into(x, gamma, beta, eps, y, G):
    """LayerNorm of rows stored as the concatenation of the G group outputs, written in the interleaved feature order
    of cruse_net.py:43-45 into a preallocated slice."""
    D = x.shape[-1]
    rows = x.numel() // D
    _call("cruse_layernorm_interleave_fwd", _p(x), _p(gamma), _p(beta), float(eps), _p(y), rows, D, G, _stream(),
          meta=(f"layernorm_il D{D}", _nb(x, y), 8 * x.numel()))


def gru_ih_gemm_into(x, w_ih, b_ih, b_hh, xproj, tables=None):
    """tcgen05 input projections of a contiguous row range into a preallocated xproj slice."""
    G = len(w_ih)
    H = w_ih[0].shape[1]
    M = x.shape[0]
    tw, tbi = tables if tables is not None else (_ptr_table(w_ih), _ptr_table(b_ih))
    _call("cruse_gru_ih_gemm_tc", _p(x), tw, tbi, _ptr_table(b_hh), _p(xproj), M, G, H, _stream(),
          meta=(f"gru_ih[tf32] G{G} H{H}", _nb(x, xproj, *w_ih), 2 * M * G * H * 3 * H))


def layernorm_fwd(x, gamma, beta, eps, residual=None, want_stats=False):
    _req(x, "x")
    _req(residual, "residual")
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    _call("cruse_layernorm_fwd", _p(x), _p(gamma), _p(beta), float(eps), _p(residual), _p(y), _p(mean), _p(rstd),
                                    rows, D, _stream(), meta=(f"layernorm D{D}", _nb(x, y, residual), 8 * x.numel()))
    return (y, mean, rstd) if want_stats else y


# ------------------------------------------------------------------------------------------
# a8 : weighted-magnitude loss
# ------------------------------------------------------------------------------------------
def layout_bctf(t):
    """reference layout [B,2,T,F] (loss.py:129-140)."""
    B, two, T, F = t.shape
    return CplxLayout(2 * T * F, F, 1, T * F)


def layout_btf2(t):
    """internal interleaved layout [B,T,NF,2]."""
    B, T, NF, two = t.shape
    return CplxLayout(T * NF * 2, NF * 2, 2, 1)


def _loss_ws(device):
    # per-call workspace from the caching allocator: keeps the op re-entrant across streams
    return torch.empty(lib().cruse_wo_male_ws_bytes() // 4, device=device, dtype=torch.float32)


def wo_male_fwd_bwd(ref, lref, est, lest, unp, lunp, B, T, F, want_grad=False):
    """-> (loss 0-dim tensor, dL/d est in est's layout | None)."""
    _req(ref, "ref")
    _req(est, "est")
    _req(unp, "unproc")
    loss = torch.empty((), device=est.device, dtype=torch.float32)
    dest = torch.zeros_like(est) if want_grad else None
    ws = _loss_ws(est.device)
    _call("cruse_wo_male_fwd_bwd", _p(ref), lref, _p(est), lest, _p(unp), lunp, _p(dest), _p(loss), _p(ws),
                                      B, T, F, _stream(),
          meta=("wo_male", B * T * F * 8 * (4 if want_grad else 3), 30 * B * T * F))
    return loss, dest


def wo_male_masked_fwd(ref, lref, mask, unp, lunp, B, T, F):
    """loss value with the estimate formed on the fly as mask * unproc (mask [B,T,F])."""
    _req(ref, "ref")
    _req(mask, "mask")
    _req(unp, "unproc")
    if mask.numel() != B * T * F:
        raise RuntimeError(f"wo_male_masked: mask has {mask.numel()} elements, expected {B}*{T}*{F}")
    loss = torch.empty((), device=mask.device, dtype=torch.float32)
    ws = _loss_ws(mask.device)
    _call("cruse_wo_male_masked_fwd", _p(ref), lref, _p(mask), _p(unp), lunp, _p(loss), _p(ws), B, T, F, _stream(),
          meta=("wo_male[mask fused]", B * T * F * (8 * 2 + 4), 30 * B * T * F))
    return loss


def spec_loss_fwd_bwd(mode, ref, lref, est, lest, B, T, F, want_grad=False):
    """mode 'MSE' (rmse, loss.py:59-78) or 'C_MSE' (c_rmse, :88-118) -> (loss 0-dim, dL/d est in est's layout | None)."""
    _req(ref, "ref")
    _req(est, "est")
    loss = torch.empty((), device=est.device, dtype=torch.float32)
    dest = torch.zeros_like(est) if want_grad else None
    ws = _loss_ws(est.device)
    _call("cruse_spec_loss_fwd_bwd", {"MSE": 0, "C_MSE": 1}[mode], _p(ref), lref, _p(est), lest, _p(dest), _p(loss), _p(ws), B, T, F,
          _stream(), meta=(f"spec_loss[{mode}]", B * T * F * 8 * (3 if want_grad else 2), 40 * B * T * F))
    return loss, dest


def wo_male_masked_partial_range(ref, lref, mask, unp, lunp, ws, p_off, nparts, B, T, F, t0, t1):
    """partial sums of the masked loss over the frames [t0, t1) into ws[p_off, p_off + nparts)."""
    _call("cruse_wo_male_masked_partial_range", _p(ref), lref, _p(mask), _p(unp), lunp, _p(ws), p_off, nparts, B, T, F, t0, t1, _stream(),
          meta=(f"wo_male[mask fused] [{t0},{t1})", B * (t1 - t0) * F * (8 * 2 + 4), 30 * B * (t1 - t0) * F))


def wo_male_finish(ws, nparts, B, T, F):
    loss = torch.empty((), device=ws.device, dtype=torch.float32)
    _call("cruse_wo_male_finish", _p(ws), nparts, B, T, F, _p(loss), _stream())
    return loss


def loss_workspace(device):
    return _loss_ws(device)


# ------------------------------------------------------------------------------------------
# a9 : backward
# ------------------------------------------------------------------------------------------
def _ws(nbytes, device):
    return torch.empty((nbytes + 3) // 4, device=device, dtype=torch.float32)


def colsum(ws, nparts, n, out, accumulate=False):
    """out[j] (+)= sum_p ws[p*n + j]"""
    _call("cruse_colsum", _p(ws), nparts, n, _p(out), 1 if accumulate else 0, _stream())
    return out


def conv_dgrad(dz, w, in_shape, kt, fstride, addend=None):
    """data gradient of conv_fwd: dz [B,T,Cout,Fout] -> din [B,T,Cin,Fin] (+ addend)."""
    _req(dz, "dz", 4)
    _req(w, "w", 4)
    _req(addend, "addend")
    B, T, Cin, Fin = in_shape
    Cout, Fout = dz.shape[2], dz.shape[3]
    din = torch.empty(B, T, Cin, Fin, device=dz.device, dtype=torch.float32)
    _call("cruse_conv_dgrad", _p(dz), _p(w), _p(addend), _p(din), B, T, Cin, Fin, Cout, Fout, kt, fstride, _stream(),
          meta=(f"conv{kt}x3 dgrad {Cout}->{Cin} F{Fout}->{Fin}", _nb(dz, din, addend, w), 2 * B * T * Cout * Fout * Cin * kt * 3))
    return din


def conv_wgrad(x, dz, kt, fstride, want_bias=True):
    """weight/bias gradient of conv_fwd -> (dw [Cout,Cin,kt,3], dbias [Cout] | None)."""
    _req(x, "x", 4)
    _req(dz, "dz", 4)
    B, T, Cin, Fin = x.shape
    Cout, Fout = dz.shape[2], dz.shape[3]
    dw = torch.empty(Cout, Cin, kt, 3, device=x.device, dtype=torch.float32)
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if want_bias else None
    ws = _ws(lib().cruse_conv_wgrad_ws_bytes(B, T, Cin, Fin, Cout, Fout, kt), x.device)
    _call("cruse_conv_wgrad", _p(x), _p(dz), _p(dw), _p(db), _p(ws), B, T, Cin, Fin, Cout, Fout, kt, fstride, _stream(),
          meta=(f"conv{kt}x3 wgrad {Cin}->{Cout} F{Fin}->{Fout}", _nb(x, dz, dw), 2 * B * T * Cout * Fout * Cin * kt * 3))
    return dw, db


def convT_dgrad(dz, w, in_shape, addend=None):
    """data gradient of convT_fwd: dz [B,T,Cout,Fout] -> din [B,T,Cin,Fin] (+ addend)."""
    _req(dz, "dz", 4)
    _req(w, "w", 4)
    _req(addend, "addend")
    B, T, Cin, Fin = in_shape
    Cout, Fout = dz.shape[2], dz.shape[3]
    din = torch.empty(B, T, Cin, Fin, device=dz.device, dtype=torch.float32)
    _call("cruse_convT_dgrad", _p(dz), _p(w), _p(addend), _p(din), B, T, Cin, Fin, Cout, Fout, _stream(),
          meta=(f"convT1x3 dgrad {Cout}->{Cin} F{Fout}->{Fin}", _nb(dz, din, addend, w), 2 * B * T * Cin * Fin * Cout * 3))
    return din


def convT_wgrad(x, dz, want_bias=True):
    """weight/bias gradient of convT_fwd -> (dw [Cin,Cout,1,3], dbias [Cout] | None)."""
    _req(x, "x", 4)
    _req(dz, "dz", 4)
    B, T, Cin, Fin = x.shape
    Cout, Fout = dz.shape[2], dz.shape[3]
    dw = torch.empty(Cin, Cout, 1, 3, device=x.device, dtype=torch.float32)
    db = torch.empty(Cout, device=x.device, dtype=torch.float32) if want_bias else None
    ws = _ws(lib().cruse_convT_wgrad_ws_bytes(B, T, Cin, Fin, Cout, Fout), x.device)
    _call("cruse_convT_wgrad", _p(x), _p(dz), _p(dw), _p(db), _p(ws), B, T, Cin, Fin, Cout, Fout, _stream(),
          meta=(f"convT1x3 wgrad {Cin}->{Cout} F{Fin}->{Fout}", _nb(x, dz, dw), 2 * B * T * Cin * Fin * Cout * 3))
    return dw, db


def bn_act_bwd(dy, z, scale, shift, alpha, act, mean, invstd, gamma, count, training=True):
    """backward of y = act(BN(z)): -> (dz, dgamma, dbeta, dalpha | None).  Two streaming passes over (dy, z)."""
    _req(dy, "dy", 4)
    _req(z, "z", 4)
    B, T, Cn, F = z.shape
    dev = z.device
    nparts = lib().cruse_bn_bwd_nparts(B * T)
    partials = torch.empty(nparts, 3 * Cn, device=dev, dtype=torch.float32)
    _call("cruse_bn_act_bwd_reduce", _p(dy), _p(z), _p(scale), _p(shift), _p(alpha), ACT[act], _p(mean), _p(invstd),
          _p(partials), B * T, Cn, F, _stream(), meta=(f"bn_act_bwd_reduce C{Cn} F{F}", _nb(dy, z), 6 * z.numel()))
    dgamma = torch.empty(Cn, device=dev, dtype=torch.float32)
    dbeta = torch.empty_like(dgamma)
    dalpha = torch.empty_like(dgamma) if act == "prelu" else None
    coef = torch.empty(3 * Cn, device=dev, dtype=torch.float32)
    _call("cruse_bn_bwd_finalize", _p(partials), nparts, Cn, float(count), _p(gamma), _p(invstd), 1 if training else 0,
          _p(dgamma), _p(dbeta), _p(dalpha), _p(coef), _stream())
    dz = torch.empty_like(z)
    _call("cruse_bn_act_bwd_apply", _p(dy), _p(z), _p(scale), _p(shift), _p(alpha), ACT[act], _p(mean), _p(invstd),
          _p(coef), _p(dz), B * T, Cn, F, _stream(), meta=(f"bn_act_bwd_apply C{Cn} F{F}", _nb(dy, z, dz), 8 * z.numel()))
    return dz, dgamma, dbeta, dalpha


def layernorm_bwd(dy, x, gamma, mean, rstd):
    """-> (dx, dgamma, dbeta)."""
    _req(dy, "dy")
    _req(x, "x")
    D = x.shape[-1]
    rows = x.numel() // D
    nparts = lib().cruse_layernorm_bwd_nparts(rows)
    partials = torch.empty(nparts, 2 * D, device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    _call("cruse_layernorm_bwd", _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(partials), rows, D, _stream(),
          meta=(f"layernorm_bwd D{D}", _nb(dy, x, dx), 12 * x.numel()))
    dgb = torch.empty(2 * D, device=x.device, dtype=torch.float32)
    _call("cruse_colsum", _p(partials), nparts, 2 * D, _p(dgb), 0, _stream())
    return dx, dgb[:D], dgb[D:]


def gru_seq_bwd(dy, y, gates, w_hh, B, T, interleave, h0=None, want_dh0=False, mode=None):
    """BPTT of one grouped-GRU layer -> (dxproj [B*T,G,3H], dpre [B*T,G,3H], dbias [G,4,H] = sums of (da_r,da_z,da_n,dhn),
    dh0 [G,B,H] | None).  mode (default GRU_SEQ_MODE): 'tf32' = tcgen05 kernel, 'fp32' = exact CUDA-core twin."""
    mode = mode or GRU_SEQ_MODE
    _req(dy, "dy", 3)
    _req(y, "y", 3)
    _req(gates, "gates", 5)
    _req(h0, "h0")
    G = len(w_hh)
    H = w_hh[0].shape[1]
    if tuple(gates.shape) != (B, T, G, 4, H) or tuple(y.shape) != (B, T, G * H) or tuple(dy.shape) != (B, T, G * H):
        raise RuntimeError(f"gru_seq_bwd: shapes dy {tuple(dy.shape)} y {tuple(y.shape)} gates {tuple(gates.shape)} vs B={B} T={T} G={G} H={H}")
    dxproj = torch.empty(B * T, G, 3 * H, device=y.device, dtype=torch.float32)
    dpre = torch.empty_like(dxproj)
    dh0 = torch.empty(G, B, H, device=y.device, dtype=torch.float32) if want_dh0 else None
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    tw = _ptr_table(w_hh)
    nsl = (B + 15) // 16
    dbp = torch.zeros(nsl, G * 4 * H, device=y.device, dtype=torch.float32)
    if mode not in ("tf32", "fp32"):
        raise RuntimeError(f"gru_seq_bwd: unknown mode {mode!r}")
    _call("cruse_gru_seq_bwd_tc" if mode == "tf32" else "cruse_gru_seq_bwd_exact", _p(dy), _p(y), _p(gates), _p(h0), tw, _p(dxproj),
          _p(dpre), _p(dh0), _p(dbp), B, T, G, H, y_fs, y_gs, _stream(),
          meta=(f"gru_seq_bwd[{mode}] G{G} H{H} T{T}", _nb(dy, y, gates, dxproj, dpre, *w_hh), 2 * B * T * G * H * 3 * H))
    dbias = torch.empty(G * 4 * H, device=y.device, dtype=torch.float32)
    _call("cruse_colsum", _p(dbp), nsl, G * 4 * H, _p(dbias), 0, _stream())
    return dxproj, dpre, dbias.view(G, 4, H), dh0


def transpose_gcm(x, M, G, Cn, ld, gs, cs, shift_T=0, h0=None, Bn=0):
    """out[g][c][m] = x[m*ld + g*gs + c*cs] (row pitch = M rounded up to 4 floats for TMA); with shift_T the source
    row is m-1 (t == 0 -> h0 or 0)."""
    M4 = (M + 3) // 4 * 4
    out = torch.empty(G, Cn, M4, device=x.device, dtype=torch.float32)
    _call("cruse_transpose_gcm", _p(x), _p(h0), _p(out), M, G, Cn, ld, gs, cs, shift_T, Bn, M4, _stream(),
          meta=(f"transpose G{G} C{Cn}", 2 * G * Cn * M * 4, 0))
    return out


def gemm_tn_tc(A, Bm, C, M, N, K, lda, ldb, ldc, bias=None, splitk=1, c_plane=0, mode=None):
    """G GEMMs C_g[m,n] = sum_k A_g[m,k] B_g[n,k] (+bias_g[n]); A/Bm/C/bias: lists of tensors (views allowed: only
    data_ptr and the given pitches are used).  mode (default GRU_IH_MODE): 'tf32' = tcgen05, 'fp32' = exact CUDA-core twin."""
    mode = mode or GRU_IH_MODE
    if mode not in ("tf32", "fp32"):
        raise RuntimeError(f"gemm_tn_tc: unknown mode {mode!r}")
    G = len(A)
    ta, tb, tcs, tbias = _ptr_table(A), _ptr_table(Bm), _ptr_table(C), _ptr_table(bias)
    _call("cruse_gemm_tn_tc" if mode == "tf32" else "cruse_gemm_tn_fp32", ta, tb, tbias, tcs, G, M, N, K, lda, ldb, ldc, splitk, c_plane,
          _stream(), meta=(f"gemm_tn[{mode}] G{G} {M}x{N}x{K} splitk{splitk}", 4 * G * (M * K + N * K + M * N * splitk), 2 * G * M * N * K))


def gemm_tc(A, Bm, C, M, N, K, lda, ldb, ldc, a_mn=False, b_mn=False, b_kshift=0, bias=None, addend=None, splitk=1, c_plane=0):
    """gemm_tn_tc with either operand MN-major (stored [K rows][M | N contiguous], row pitch lda / ldb): C_g[m,n] =
    sum_k A_g[m,k] B_g[n,k] (+ addend_g[m,n], pitch ldc); b_kshift pairs A's reduction index k with B's row k - b_kshift
    (rows < 0 are zero).  tf32 only."""
    G = len(A)
    ta, tb, tcs, tbias, tadd = _ptr_table(A), _ptr_table(Bm), _ptr_table(C), _ptr_table(bias), _ptr_table(addend)
    _call("cruse_gemm_tc", ta, tb, tbias, tadd, tcs, G, M, N, K, lda, ldb, ldc, splitk, c_plane, 1 if a_mn else 0, 1 if b_mn else 0,
          b_kshift, _stream(), meta=(f"gemm[tf32,{'mn' if a_mn else 'k'}/{'mn' if b_mn else 'k'}] G{G} {M}x{N}x{K} splitk{splitk}",
                                     4 * G * (M * K + N * K + M * N * splitk), 2 * G * M * N * K))


def sigmoid_bwd(dy, y):
    _req(dy, "dy")
    _req(y, "y")
    dz = torch.empty_like(y)
    _call("cruse_sigmoid_bwd", _p(dy), _p(y), _p(dz), y.numel(), _stream(), meta=("sigmoid_bwd", _nb(dy, y, dz), 3 * y.numel()))
    return dz


# ------------------------------------------------------------------------------------------
# numeric mode of the eval-mode conv stages (include/cruse_b200.h: cruse_conv_set_mode)
# ------------------------------------------------------------------------------------------
def set_conv_mode(mode: str):
    """'tf32' = tcgen05 implicit GEMM (default), 'fp32' = exact-fp32 CUDA-core kernels."""
    if mode not in ("tf32", "fp32"):
        raise RuntimeError(f"set_conv_mode: unknown mode {mode!r}")
    check(lib().cruse_conv_set_mode(1 if mode == "tf32" else 0), "cruse_conv_set_mode")


def get_conv_mode() -> str:
    return "tf32" if lib().cruse_conv_get_mode() == 1 else "fp32"


def set_conv_max_ctas(n: int):
    """cap (0 = none) on the persistent grid of the tensor-core conv stages launched from now on."""
    check(lib().cruse_conv_set_max_ctas(int(n)), "cruse_conv_set_max_ctas")


# training backward: weight-gradient kernels on a side stream beside the BPTT launches (autograd._SideWork)
OVERLAP_BWD = os.environ.get("CRUSE_OVERLAP_BWD", "1") != "0"
BWD_SIDE_CAP = os.environ.get("CRUSE_BWD_SIDE_CAP", "1") != "0"
BWD_PRIORITY = os.environ.get("CRUSE_BWD_PRIORITY", "1") != "0"         # backward: dependent chain on a priority stream, the rest early on the side
BWD_TAIL_CAP = int(os.environ.get("CRUSE_BWD_TAIL_CAP", "0"))          # cap (CTAs) on the encoder weight gradients that run beside the encoder chain; 0 = none
GEMM_MN_MAJOR = os.environ.get("CRUSE_GEMM_MN_MAJOR", "1") != "0"       # GRU weight-gradient / dx GEMMs read their factors in place
FWD_SIDE_SKIPS = os.environ.get("CRUSE_FWD_SIDE_SKIPS", "1") != "0"   # training forward: skip convs beside the layer-1 recurrence
BWD_SIDE_L1 = os.environ.get("CRUSE_BWD_SIDE_L1", "1") != "0"      # layer-1 GRU weight gradients beside the encoder backward

# inference: pipeline the head of the encoder and the tail of the decoder with the GRU wavefront (cruse_net.GGRU._wavefront)
PIPELINE_EDGES = os.environ.get("CRUSE_PIPELINE_EDGES", "1") != "0"

# run the skip convs (and the clean-speech STFT) on a low-priority side stream beside the GRU wavefront
OVERLAP_SKIPS = os.environ.get("CRUSE_OVERLAP_SKIPS", "1") != "0"
SKIP_MAX_CTAS = int(os.environ.get("CRUSE_SKIP_MAX_CTAS", "0"))     # 0 = no cap (measured best on B200: 1.90 vs 1.94 ms at 80)


def pcm16_to_float(src, dst):
    """int16 PCM samples (device) -> float32 = src / 32768 (what soundfile / librosa return for a 16-bit wav file), in ``dst``"""
    if src.dtype != torch.int16 or dst.dtype != torch.float32 or src.numel() != dst.numel() or not src.is_cuda or not dst.is_cuda:
        raise RuntimeError("pcm16_to_float: need an int16 and a float32 CUDA tensor of the same size")
    if not (src.is_contiguous() and dst.is_contiguous()):
        raise RuntimeError("pcm16_to_float: contiguous tensors only")
    _call("cruse_pcm16_to_float", src.data_ptr(), _p(dst), src.numel(), _stream())
    return dst

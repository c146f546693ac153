"""Loss surface of the reference on the fused sm_100a kernel.

``loss_func(mode).loss(inputs, labels, noisy)`` mirrors loss_func/loss.py:16-34 and
``wo_male(ref, est, unproc)`` mirrors :121-148 (tensors ``[B,2,T,F]``, channel 0 = real,
1 = imag).  ``wo_male`` is differentiable w.r.t. ``est``: forward value and gradient come
out of the same streaming pass.  Factory ``wo_male_loss()`` follows the name -> callable
convention of train_base/loss.py that tools/train_stand.py:73-75 uses.
"""
from __future__ import annotations

import torch

from . import ops


class _WoMale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, est, unproc, layout, nbins=None):
        lay = ops.layout_bctf if layout == "bctf" else ops.layout_btf2
        if layout == "bctf":
            B, _, T, F = est.shape
        else:
            B, T, F, _ = est.shape
        if nbins is not None:
            F = nbins                      # loss over bins [0, nbins) only; the gradient of the other bins is 0
        need = est.requires_grad
        ref_c, est_c, unp_c = ref.contiguous(), est.contiguous(), unproc.contiguous()
        loss, dest = ops.wo_male_fwd_bwd(ref_c, lay(ref_c), est_c, lay(est_c), unp_c, lay(unp_c), B, T, F,
                                         want_grad=need)
        ctx.save_for_backward(dest)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dest,) = ctx.saved_tensors
        return None, (dest * g if dest is not None else None), None, None, None


class _WoMaleOfMask(torch.autograd.Function):
    """wo_male(S, mask * X, X) on interleaved spectra [B,T,NF,2] over the masked bins, differentiable w.r.t. the mask: what the
    training step composes from utils/utils.py:417-420 (est = mask * X) and loss_func/loss.py:121-148 as ONE autograd node -- the
    upstream scalar gradient scales the mask gradient inside the mask_bwd kernel instead of a pass of its own over dL/d est."""

    @staticmethod
    def forward(ctx, ref, mask, unproc, n_fft, hop):
        from .acoustics import hann_window
        B, T, F = mask.shape
        est, _ = ops.mask_istft_fwd(unproc, mask, hann_window(n_fft, n_fft, unproc.device), n_fft, hop, 0, want_est=True, want_wav=False)
        loss, dest = ops.wo_male_fwd_bwd(ref, ops.layout_btf2(ref), est, ops.layout_btf2(est), unproc, ops.layout_btf2(unproc), B, T, F,
                                         want_grad=True)
        ctx.save_for_backward(dest, unproc)
        ctx.F = F
        return loss

    @staticmethod
    def backward(ctx, g):
        dest, unproc = ctx.saved_tensors
        return None, ops.mask_bwd(dest, unproc, ctx.F, gscale=g.reshape(1).contiguous().float()), None, None, None


def wo_male_of_mask(ref, mask, unproc, n_fft, hop):
    """loss of the masked noisy spectrum against the clean one, mask [B,T,F] (requires grad), spectra [B,T,NF,2]"""
    return _WoMaleOfMask.apply(ref.contiguous(), mask.contiguous(), unproc.contiguous(), n_fft, hop)


def wo_male(ref, est, unproc, norm=False, eps=1e-8):
    """loss_func/loss.py:121-148 (repairs: App. A.4).  ref/est/unproc [B,2,T,F] -> 0-dim loss."""
    if ref.shape != est.shape:
        raise RuntimeError(f"Dimension mismatch when calculate wo-male, {ref.shape} vs {est.shape}")   # :122-125
    if unproc.shape != ref.shape:
        raise RuntimeError(f"Dimension mismatch when calculate wo-male, {ref.shape} vs {unproc.shape}")
    return _WoMale.apply(ref, est, unproc, "bctf")


def wo_male_frames(ref, est, unproc, F):
    """same loss on internal interleaved spectra [B,T,NF,2], restricted to bins [0,F) -- zero-copy."""
    B, T, NF, _ = est.shape
    loss, _ = ops.wo_male_fwd_bwd(ref, ops.layout_btf2(ref), est, ops.layout_btf2(est), unproc,
                                  ops.layout_btf2(unproc), B, T, F, want_grad=False)
    return loss


def wo_male_frames_masked(ref, mask, unproc, F):
    """the same value with est = mask * unproc formed inside the loss kernel (mask [B,T,F]): no dependence on the stored
    estimate, so it can run beside the mask*spectrum + iSTFT kernel."""
    B, T, NF, _ = unproc.shape
    return ops.wo_male_masked_fwd(ref, ops.layout_btf2(ref), mask, unproc, ops.layout_btf2(unproc), B, T, F)


def wo_male_frames_autograd(ref, est, unproc, F):
    """wo_male on interleaved spectra [B,T,NF,2] over bins [0,F), differentiable w.r.t. est (training path)."""
    return _WoMale.apply(ref, est, unproc, "btf2", F)


class _SpecLoss(torch.autograd.Function):
    """rmse / c_rmse on [B,2,T,F] spectra, differentiable w.r.t. est."""

    @staticmethod
    def forward(ctx, ref, est, mode):
        B, _, T, F = est.shape
        ref_c, est_c = ref.contiguous().float(), est.contiguous().float()
        loss, dest = ops.spec_loss_fwd_bwd(mode, ref_c, ops.layout_bctf(ref_c), est_c, ops.layout_bctf(est_c), B, T, F,
                                           want_grad=est.requires_grad)
        ctx.save_for_backward(dest)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dest,) = ctx.saved_tensors
        return None, (dest * g if dest is not None else None), None


def _check_bctf(ref, est, what):
    if ref.shape != est.shape:
        raise RuntimeError(f"Dimension mismatch when calculate {what}, {tuple(ref.shape)} vs {tuple(est.shape)}")   # loss.py:65-68, 89-92
    if est.dim() != 4 or est.shape[1] != 2:
        raise RuntimeError(f"{what}: expected [B,2,T,F], got {tuple(est.shape)}")
    if not est.is_cuda:
        raise RuntimeError(f"{what}: cruse_b200 runs on sm_100a only (no CPU fallback)")


def rmse(ref, est, eps=1e-8):
    """loss_func/loss.py:59-78 ('MSE' mode): sum sqrt((est-ref)^2) / (B*T*F) on [B,2,T,F]."""
    _check_bctf(ref, est, "rmse")
    return _SpecLoss.apply(ref, est, "MSE")


def c_rmse(ref, est, unproc=None, norm=False, eps=1e-8):
    """loss_func/loss.py:88-118 ('C_MSE' mode): power-law compressed complex error, arithmetic kept literally."""
    _check_bctf(ref, est, "c_mse")
    return _SpecLoss.apply(ref, est, "C_MSE")


class _SiSnr(torch.autograd.Function):
    """mean SI-SNR (dB) of est vs ref waveforms, differentiable w.r.t. est (loss_func/loss.py:37-56)."""

    @staticmethod
    def forward(ctx, est, ref, eps):
        est, ref = est.contiguous(), ref.contiguous()
        value, ws = ops.sisnr_fwd(est, ref, eps)
        ctx.save_for_backward(est, ref, ws)
        return value

    @staticmethod
    def backward(ctx, g):
        est, ref, ws = ctx.saved_tensors
        return ops.sisnr_bwd(est, ref, ws, g.contiguous().float()), None, None


def sisnr(s1, s2, eps=1e-8):
    """loss_func/loss.py:47-56: s1 = estimate, s2 = target, [B, L] (any leading dims are flattened) -> mean SI-SNR in dB."""
    if s1.shape != s2.shape:
        raise RuntimeError(f"Dimension mismatch when calculate si-snr, {tuple(s1.shape)} vs {tuple(s2.shape)}")
    if not s1.is_cuda:
        raise RuntimeError("sisnr: cruse_b200 runs on sm_100a only (no CPU fallback)")
    L = s1.shape[-1]
    return _SiSnr.apply(s1.reshape(-1, L).float(), s2.reshape(-1, L).float(), eps)


class _SiSnrZm(torch.autograd.Function):
    """zero-mean SI-SNR loss (train_base/loss.py:7-25), differentiable w.r.t. the estimate."""

    @staticmethod
    def forward(ctx, x, s, eps):
        x, s = x.contiguous(), s.contiguous()
        value, ws = ops.si_snr_zm_fwd(x, s, eps)
        ctx.save_for_backward(x, s, ws)
        return value

    @staticmethod
    def backward(ctx, g):
        x, s, ws = ctx.saved_tensors
        return ops.si_snr_zm_bwd(x, s, ws, g.contiguous().float()), None, None


def si_snr_loss():
    """train_base/loss.py:7-25: the factory tools/train_stand.py:73-75 resolves by name (``getattr(train_base.loss, name)(**args)``);
    returns ``si_snr(x, s, eps=1e-8)`` on waveforms [..., L] -> 0-dim loss (negative mean SI-SNR of the zero-mean signals, in dB)."""
    def si_snr(x, s, eps=1e-8):
        if x.shape != s.shape:
            raise RuntimeError(f"Dimension mismatch when calculate si_snr, {x.shape} vs {s.shape}")        # :13-16
        if not x.is_cuda:
            raise RuntimeError("si_snr: cruse_b200 runs on sm_100a only (no CPU fallback)")
        L = x.shape[-1]
        return _SiSnrZm.apply(x.reshape(-1, L).float(), s.reshape(-1, L).float(), eps)
    return si_snr


def wo_male_loss():
    """factory in the style of train_base/loss.py:7-25: returns loss(est, ref, noisy)."""
    def loss(est, ref, noisy):
        return wo_male(ref, est, noisy)
    loss.cruse_kind = "wo_male"          # cruse_b200.trainer.Trainer runs the captured STFT + forward + wo_male + backward step for it
    return loss


class loss_func:
    """loss_func/loss.py:16-34 dispatcher; WO_MALE (the hot path), SI-SNR, C_MSE and MSE are built -- every mode the
    reference's dispatcher itself implements (:24-34; the remaining names fall through to None there)."""

    MODES = ['SI-SNR', 'SS-SNR', 'MSE', 'Normal_MSE', 'CN_MSE', 'D_MSE', 'WO_MALE', 'C_MSE']

    def __init__(self, loss_mode):
        assert loss_mode in self.MODES, "Loss mode must be one of ***"      # :19-21
        self.loss_mode = loss_mode

    def loss(self, inputs, labels, noisy=None):
        if self.loss_mode == 'WO_MALE':
            return wo_male(labels, inputs, noisy)                            # :29-30 (arg order)
        if self.loss_mode == 'SI-SNR':
            return -(sisnr(inputs, labels))                                  # :25-26 (time-domain estimate / target)
        if self.loss_mode == 'SS-SNR':
            return 0                                                         # :27-28
        if self.loss_mode == 'C_MSE':
            return c_rmse(labels, inputs)                                    # :31-32
        if self.loss_mode == 'MSE':
            return rmse(labels, inputs)                                      # :33-34
        raise NotImplementedError(f"loss mode {self.loss_mode!r} is outside the built hot path (SURVEY.md 8f2)")

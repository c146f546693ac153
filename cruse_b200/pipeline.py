"""The hot path end to end: wav -> STFT -> U-Net mask ->
mask*spectrum -> iSTFT (+ weighted-magnitude loss against the clean spectrum).

Everything between the input waveform and the outputs stays in HBM in the kernels' own
frame-major layouts; this file only sequences launches on the current stream.
"""
from __future__ import annotations

import torch

from . import ops
from .acoustics import hann_window, stft_frames
from .loss import wo_male_frames

EPS_MAG = 1e-8   # utils/utils.py:400


def enhance(model, noisy, n_fft=512, hop=320, pad_mode="reflect"):
    """noisy wav [B,L] -> (enhanced wav [B,L], est spec [B,T,NF,2], mask [B,T,F], noisy spec [B,T,NF,2])."""
    F = model.in_feat
    X, mag = stft_frames(noisy, n_fft, hop, n_fft, pad_mode, mag_bins=F, mag_eps=EPS_MAG)   # feature.py:10-30, utils.py:400
    mask = model.forward_frames(mag)                                                        # cruse_net.py:147-165
    est, wav = ops.mask_istft_fwd(X, mask, hann_window(n_fft, n_fft, noisy.device), n_fft, hop, noisy.shape[-1])
    return wav, est, mask, X


def forward_loss(model, noisy, clean, n_fft=512, hop=320, pad_mode="reflect"):
    """STFT + forward + mask + iSTFT + wo_male over the F bins the net sees -> (loss, wav, est, mask)."""
    wav, est, mask, X = enhance(model, noisy, n_fft, hop, pad_mode)
    S, _ = stft_frames(clean, n_fft, hop, n_fft, pad_mode)
    loss = wo_male_frames(S, est, X, model.in_feat)                                          # loss.py:121-148
    return loss, wav, est, mask


def forward_loss_host(model, noisy_host, clean_host, n_fft=512, hop=320, pad_mode="reflect", want_wav=False):
    """Reference-facing call with HOST buffers (pinned or pageable CPU tensors): copies in, runs the
    path, copies the loss (and optionally the enhanced waveform) back.  Used for the e2e bench."""
    dev = next(model.parameters()).device
    noisy = noisy_host.to(dev, non_blocking=True)
    clean = clean_host.to(dev, non_blocking=True)
    loss, wav, _, _ = forward_loss(model, noisy, clean, n_fft, hop, pad_mode)
    out = loss.to("cpu", non_blocking=False)
    if want_wav:
        return out, wav.to("cpu")
    return out


def train_forward_loss(model, noisy, clean, n_fft=512, hop=320, pad_mode="reflect"):
    """Training step forward: STFT + U-Net (saving what backward needs) + mask*X + wo_male -> loss with autograd
    history; ``loss.backward()`` runs the sm_100a backward kernels and fills ``param.grad``."""
    from .autograd import mask_apply, unet2_frames_autograd
    from .loss import wo_male_frames_autograd
    F = model.in_feat
    X, mag = stft_frames(noisy, n_fft, hop, n_fft, pad_mode, mag_bins=F, mag_eps=EPS_MAG)
    S, _ = stft_frames(clean, n_fft, hop, n_fft, pad_mode)
    mask = unet2_frames_autograd(model, mag)
    est = mask_apply(mask, X, n_fft, hop)
    return wo_male_frames_autograd(S, est, X, F)

"""The step in front of the hot path, on the device: the input feature norms of
train_base/model/base_model.py:202-300 and the on-the-fly mixing of dataset/dataset.py:236-264 (snr_mix with room impulse responses),
so that an 8-GPU job is not fed by a CPU dataloader.  Thin wrappers over csrc/frontend.cu; CUDA tensors only."""
from __future__ import annotations

import torch

from . import ops
from ._lib import lib

_MODES = {"offline_laplace_norm": 0, "cumulative_laplace_norm": 1, "offline_gaussian_norm": 2, "cumulative_layer_norm": 3}


def feature_norm(x, norm_type: str):
    """x: frame-major magnitudes [B,T,F] (what pipeline / acoustics.stft_frames produce) or the reference's [B,1,F,T]
    (base_model.py: ``input: [B, C, F, T]``); returns the same layout.  ``norm_type`` as in ``BaseModel.norm_wrapper`` (:302-315)."""
    if norm_type not in _MODES:
        raise NotImplementedError("You must set up a type of Norm. e.g. offline_laplace_norm, cumulative_laplace_norm, forgetting_norm, etc.")
    if not x.is_cuda:
        raise RuntimeError("feature_norm: cruse_b200 runs on sm_100a only (no CPU fallback)")
    ref_layout = x.dim() == 4
    if ref_layout:
        if x.shape[1] != 1:
            raise RuntimeError(f"feature_norm: [B,C,F,T] input needs C == 1 (the path is single channel), got {tuple(x.shape)}")
        frames = x[:, 0].transpose(1, 2).contiguous()
    elif x.dim() == 3:
        frames = x.contiguous()
    else:
        raise RuntimeError(f"feature_norm: expected [B,T,F] or [B,1,F,T], got {tuple(x.shape)}")
    B, T, F = frames.shape
    y = torch.empty_like(frames)
    ops._call("cruse_feature_norm", ops._p(frames), ops._p(y), B, T, F, _MODES[norm_type], ops._stream(),
              meta=(f"feature_norm[{norm_type}]", 2 * 4 * frames.numel(), 4 * frames.numel()))
    return y.transpose(1, 2).unsqueeze(1).contiguous() if ref_layout else y


def cumulative_laplace_norm(x):
    return feature_norm(x, "cumulative_laplace_norm")


def cumulative_layer_norm(x):
    return feature_norm(x, "cumulative_layer_norm")


def offline_laplace_norm(x):
    return feature_norm(x, "offline_laplace_norm")


def offline_gaussian_norm(x):
    return feature_norm(x, "offline_gaussian_norm")


def rir_conv(y, rir):
    """scipy.signal.fftconvolve(y, rir)[:len(y)] per utterance (dataset.py:244-247): y [B,L], rir [R] (shared) or [B,R]."""
    y = y.contiguous().float()
    rir = rir.contiguous().float()
    B, L = y.shape
    R = rir.shape[-1]
    stride = 0 if rir.dim() == 1 else R
    if rir.dim() == 2 and rir.shape[0] != B:
        raise RuntimeError(f"rir_conv: rir batch {rir.shape[0]} != {B}")
    out = torch.empty_like(y)
    ops._call("cruse_rir_conv", ops._p(y), ops._p(rir), ops._p(out), B, L, R, stride, ops._stream(),
              meta=(f"rir_conv R{R}", 2 * 4 * y.numel(), 2 * y.numel() * R))
    return out


def snr_mix(clean_y, noise_y, snr, target_dB_FS=None, rir=None, rir_noise=None, eps=1e-7):
    """dataset/dataset.py:236-264 for a batch on the device: clean_y / noise_y [B,L]; ``snr`` and ``target_dB_FS`` scalars or [B]
    tensors (the reference draws both per item on the host, :224-230,262-264); returns (noisy [B,L], clean [B,L])."""
    if not clean_y.is_cuda:
        raise RuntimeError("snr_mix: cruse_b200 runs on sm_100a only (no CPU fallback)")
    if clean_y.shape != noise_y.shape or clean_y.dim() != 2:
        raise RuntimeError(f"snr_mix: clean / noise must both be [B,L], got {tuple(clean_y.shape)} / {tuple(noise_y.shape)}")
    clean_y, noise_y = clean_y.contiguous().float(), noise_y.contiguous().float()
    if rir is not None:
        clean_y = rir_conv(clean_y, rir.to(clean_y.device))
    if rir_noise is not None:
        noise_y = rir_conv(noise_y, rir_noise.to(clean_y.device))
    B, L = clean_y.shape
    dev = clean_y.device
    snr_t = torch.as_tensor(snr, dtype=torch.float32, device=dev).expand(B).contiguous()
    lvl_t = None if target_dB_FS is None else torch.as_tensor(target_dB_FS, dtype=torch.float32, device=dev).expand(B).contiguous()
    noisy, clean = torch.empty_like(clean_y), torch.empty_like(clean_y)
    ws = torch.empty(lib().cruse_snr_mix_ws_bytes(B) // 4, device=dev, dtype=torch.float32)
    ops._call("cruse_snr_mix", ops._p(clean_y), ops._p(noise_y), ops._p(snr_t), ops._p(lvl_t), ops._p(noisy), ops._p(clean), ops._p(ws),
              B, L, float(eps), ops._stream(), meta=("snr_mix", 4 * 4 * clean_y.numel(), 8 * clean_y.numel()))
    return noisy, clean

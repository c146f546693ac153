"""STFT front-end / iSTFT back-end with the reference's call surface.

Mirrors ``stft`` / ``istft`` of train_base/acoustics/feature.py:10-61 and ``PreProcess`` of
utils/utils.py:365-455 (pre_stft / masking / reconstruction), running on the framed-FFT
kernels of libcruse_sm100.so.  Internally spectra are ``[B, T, NF, 2]`` (frame-major,
complex-interleaved); the reference layouts are returned as zero-copy views of that buffer.
"""
from __future__ import annotations

import torch

from . import ops

_windows = {}


def hann_window(n_fft, win_length, device):
    """periodic hann(win_length), zero-padded (centred) to n_fft like torch.stft does (feature.py:22-30)."""
    key = (n_fft, win_length, device.type, device.index)
    w = _windows.get(key)
    if w is None:
        if win_length > n_fft:
            raise RuntimeError(f"win_length {win_length} > n_fft {n_fft}")
        w = torch.hann_window(win_length, periodic=True, dtype=torch.float32, device=device)
        if win_length < n_fft:
            left = (n_fft - win_length) // 2
            w = torch.nn.functional.pad(w, (left, n_fft - win_length - left))
        _windows[key] = w.contiguous()
        w = _windows[key]
    return w


def stft_frames(y, n_fft, hop_length, win_length=None, pad_mode="reflect", mag_bins=0, mag_eps=1e-8):
    """y [B,L] -> (spec [B,T,NF,2], mag [B,T,mag_bins] | None).  Zero-copy internal entry."""
    if y.dim() != 2:
        raise AssertionError("stft expects [B, L]")            # feature.py:21 assert
    win_length = n_fft if win_length is None else win_length
    y = y.contiguous()
    return ops.stft_fwd(y, hann_window(n_fft, win_length, y.device), n_fft, hop_length, pad_mode, mag_bins, mag_eps)


def stft(y, n_fft, hop_length, win_length, pad_mode="reflect"):
    """feature.py:10-30: y [B,L] -> complex [B,F,T] (a transposed view of the kernel's output)."""
    spec, _ = stft_frames(y, n_fft, hop_length, win_length, pad_mode)
    return torch.view_as_complex(spec).transpose(1, 2)


def _to_frames(features):
    """complex [B,F,T] or real [B,F,T,2] -> contiguous real [B,T,F,2] (free when it came from ``stft``)."""
    if torch.is_complex(features):
        features = torch.view_as_real(features)                # [B,F,T,2]
    return features.transpose(1, 2).contiguous()


def istft(features, n_fft, hop_length, win_length, length=None, use_mag_phase=False):
    """feature.py:33-61: complex [B,F,T] (or real [B,F,T,2], or (mag, phase)) -> wav [B,L]."""
    if use_mag_phase:
        assert isinstance(features, (tuple, list))            # feature.py:46
        mag, phase = features
        features = torch.stack([mag * torch.cos(phase), mag * torch.sin(phase)], dim=-1)
    frames = _to_frames(features)
    B, T, NF, _ = frames.shape
    if length is None:
        length = hop_length * (T - 1)                          # torch.istft default with center=True
    _, wav = ops.mask_istft_fwd(frames, None, hann_window(n_fft, win_length, frames.device), n_fft, hop_length,
                                length, want_est=False, want_wav=True)
    return wav


class PreProcess:
    """utils/utils.py:365-455.  Shapes follow the reference: ``pre_stft`` -> (stft_inputs [B,2,T,F],
    real, imag, mags, phase each [B,1,T,F]); ``masking`` -> [B,T,F,2]; ``reconstruction`` -> wav."""

    def __init__(self, win_len, win_inc, fft_len, win_type="hanning", post_process_mode="mag_mapping",
                 loss_mode="freq", use_cuda=True):
        if win_type != "hanning":
            raise ValueError("ERROR window type")               # utils.py:385
        self.win_len, self.win_inc, self.fft_len = win_len, win_inc, fft_len
        self.post_process_mode, self.loss_mode = post_process_mode, loss_mode
        self.spec = None

    def pre_stft(self, inputs):
        NF = self.fft_len // 2 + 1
        spec, mag = stft_frames(inputs, self.fft_len, self.win_inc, self.win_len, pad_mode="constant",
                                mag_bins=NF, mag_eps=1e-8)     # utils.py:390-400
        self.spec = spec                                        # [B,T,F,2]
        self.real = spec[..., 0].unsqueeze(1)                   # [B,1,T,F] views
        self.imag = spec[..., 1].unsqueeze(1)
        self.spec_mags = mag.unsqueeze(1)
        stft_inputs = spec.permute(0, 3, 1, 2)                  # [B,2,T,F] view (utils.py:397)
        return stft_inputs, self.real, self.imag, self.spec_mags, self.spec_phase

    @property
    def spec_phase(self):
        # utils.py:401; not consumed anywhere on the hot path, so it is derived lazily
        return torch.atan2(self.imag, self.real)

    def masking(self, mask_real, mask_imag=None):
        """utils.py:417-433.  mag_mapping runs fused on the kernel; the other two modes are
        elementwise layout plumbing kept for API completeness."""
        if self.post_process_mode == "mag_mapping":
            B, T, NF, _ = self.spec.shape
            m = mask_real.contiguous().view(B, T, -1)
            est, _ = ops.mask_istft_fwd(self.spec, m, hann_window(self.fft_len, self.win_len, m.device), self.fft_len,
                                        self.win_inc, 0, want_est=True, want_wav=False)
            return est
        if self.post_process_mode == "complex_mapping":
            out_real, out_imag = mask_real * self.real, mask_imag * self.imag
        elif self.post_process_mode == "mapping":
            out_real, out_imag = mask_real, mask_imag
        else:
            raise NotImplementedError
        return torch.stack([out_real.squeeze(1), out_imag.squeeze(1)], dim=-1).contiguous()

    def reconstruction(self, stft_outputs, sig_len=None):
        """utils.py:443-455: [B,T,F,2] -> wav [B,L]."""
        frames = stft_outputs.contiguous()
        B, T, NF, _ = frames.shape
        if sig_len is None:
            sig_len = self.win_inc * (T - 1)
        _, wav = ops.mask_istft_fwd(frames, None, hann_window(self.fft_len, self.win_len, frames.device),
                                    self.fft_len, self.win_inc, sig_len, want_est=False, want_wav=True)
        return wav

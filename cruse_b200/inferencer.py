"""The inferencer shell around the hot path (train_base/inferencer/base_inferencer.py:120-196).

``Inferencer._load_model`` reads the reference trainer's ``.tar`` checkpoints (key ``model`` + ``epoch``, :121-135),
``multi_channel_mag_to_mag`` is the reference's inference type for magnitude models (:138-161: magnitude in, enhanced magnitude out,
phase of the reference channel, iSTFT) and ``__call__`` the per-file loop with its real-time-factor print and int16 scaling
(:163-196).  With the mask-estimating ``unet_2`` the enhanced magnitude is ``mask * |X|``, so magnitude * cos / sin of the noisy phase
is ``mask * X``: the whole method is one call of the fused path (STFT -> U-Net -> mask*X -> iSTFT), nothing is recomputed in torch.
"""
from __future__ import annotations

import time
from pathlib import Path

import numpy as np
import torch

from . import checkpoint, pipeline


class Inferencer:
    def __init__(self, model, acoustic_config, device=None, enhanced_dir=None):
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise RuntimeError("cruse_b200.inferencer: the hot path runs on sm_100a only (no CPU fallback)")
        self.model = model.to(self.device).eval()
        self.acoustic_config = dict(acoustic_config)                     # sr, n_fft, hop_length, win_length (base_inferencer.py:40-56)
        self.n_fft, self.hop = self.acoustic_config["n_fft"], self.acoustic_config["hop_length"]
        self.enhanced_dir = Path(enhanced_dir) if enhanced_dir is not None else None
        self.rtf = []                                                    # (name, real-time factor) of every file processed

    @staticmethod
    def _load_model(model, checkpoint_path, device):
        """base_inferencer.py:121-135 for an already constructed module: weights from ``checkpoint["model"]``, eval mode;
        returns (model, epoch)."""
        ckpt = checkpoint._load(Path(checkpoint_path).expanduser().absolute().as_posix(), "cpu")
        model.load_state_dict(ckpt["model"])
        print(f"loaded a .tar checkpoint, epoch {ckpt['epoch']}.")
        return model.to(device).eval(), ckpt["epoch"]

    @torch.no_grad()
    def multi_channel_mag_to_mag(self, noisy, inference_args=None):
        """noisy [B, C, L] (or [B, L]) on the device -> enhanced waveform as numpy (batch dim squeezed as in :159); the model sees the
        magnitude of the reference channel 0 (``unet_2`` is single-channel) and the output keeps the noisy phase."""
        if noisy.dim() == 3:
            noisy = noisy[:, 0]                                          # reference channel (:150-151)
        if noisy.dim() != 2:
            raise RuntimeError(f"multi_channel_mag_to_mag: expected [B, C, L] or [B, L], got {tuple(noisy.shape)}")
        wav, _, _, _ = pipeline.enhance(self.model, noisy.contiguous().float(), self.n_fft, self.hop)
        return wav.detach().squeeze(0).cpu().numpy()

    @torch.no_grad()
    def __call__(self, dataloader, inference_type="multi_channel_mag_to_mag", inference_args=None):
        """the per-file loop of base_inferencer.py:163-196: batch size 1, real-time factor printed per file, int16 output scaled to
        0.8 of full scale; files are written only when ``enhanced_dir`` was given.  Returns {name: int16 array}."""
        assert inference_type in dir(self), f"Not implemented Inferencer type: {inference_type}"
        out = {}
        for batch in dataloader:
            noisy, name = batch[0], batch[-1]
            assert len(name) == 1, "The batch size of inference stage must 1."
            name = name[0]
            t1 = time.time()
            enhanced = getattr(self, inference_type)(noisy.to(self.device), inference_args)     # .cpu() inside: synchronises
            t2 = time.time()
            if (abs(enhanced) > 1).any():
                print(f"Warning: enhanced is not in the range [-1, 1], {name}")
            amp = np.iinfo(np.int16).max
            enhanced = np.int16(0.8 * amp * enhanced / np.max(np.abs(enhanced)))
            rtf = (t2 - t1) / (len(enhanced) * 1.0 / self.acoustic_config["sr"])
            print(f"{name}, rtf: {rtf}")
            self.rtf.append((name, rtf))
            if self.enhanced_dir is not None:
                from scipy.io import wavfile                             # (the reference writes with soundfile, absent from the image)
                self.enhanced_dir.mkdir(parents=True, exist_ok=True)
                wavfile.write((self.enhanced_dir / f"{name}.wav").as_posix(), self.acoustic_config["sr"], enhanced)
            out[name] = enhanced
        return out

"""The direct caller of the hot path in training.

``TrainStep``: one optimisation step in the shape of the reference's trainer loop -- ``loss.backward()``, gradient clipping,
``optimizer.step()`` (train_base/trainer/base_trainer.py:378-430; DDP gradient averaging :31; Adam from tools/train_stand.py:68) --
built on the captured train step, so that per batch the host issues one graph replay, one in-place all_reduce on the flat gradient
buffer (N > 1), one norm + scale on that buffer and one fused Adam launch.

``Trainer``: the concrete trainer class the reference's launcher instantiates (``initialize_module(config["trainer"]["path"],
initialize=False)`` then ``trainer_class(dist=..., rank=..., config=..., resume=..., only_validation=..., model=..., loss_function=...,
optimizer=..., train_dataloader=..., validation_dataloader=...).train()``, tools/train_stand.py:76-90).  It keeps ``BaseTrainer``'s
constructor reads and epoch loop (base_trainer.py:26-128, :378-424) and supplies the two methods the reference leaves abstract
(``_train_epoch`` / ``_validation_epoch``, :426-430; its own ``train/trainer_casual.py`` is absent).  Checkpoints are the reference's
files (cruse_b200.checkpoint).  No DistributedDataParallel wrapper: the gradients of a step live in one flat buffer that is averaged
by ONE all_reduce on the process group the launcher initialised (loss_func/distrib.py:100-116 semantics).
"""
from __future__ import annotations

import time
from pathlib import Path

import torch

from . import checkpoint, distrib
from .pipeline import CapturedTrainStep


class TrainStep:
    """step(noisy, clean) -> loss (0-dim device tensor, this batch's loss before the update).

    ``max_grad_norm``: clip_grad_norm_ semantics of the reference (base_trainer.py ``clip_grad_norm_value``) on the global
    L2 norm, computed on the flat buffer (one reduction instead of one per parameter); None = no clipping."""

    def __init__(self, model, B, L, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, max_grad_norm=None, n_fft=512, hop=320,
                 optimizer=None, group=None):
        self.model = model
        self.captured = CapturedTrainStep(model, B, L, n_fft, hop)
        self.params = self.captured.params
        self.group = group
        self.max_grad_norm = max_grad_norm
        self.optimizer = optimizer if optimizer is not None else torch.optim.Adam(self.params, lr=lr, betas=betas,
                                                                                 weight_decay=weight_decay, fused=True)
        self.last_grad_norm = None

    def step(self, noisy, clean):
        loss = self.captured(noisy, clean)                               # param.grad = views of captured.flat_grad
        flat = self.captured.flat_grad
        distrib.sync_grad(self.params, group=self.group, flat=flat)      # no-op on one process
        if self.max_grad_norm is not None:
            norm = torch.linalg.vector_norm(flat)
            self.last_grad_norm = norm
            flat.mul_(torch.clamp(self.max_grad_norm / (norm + 1e-6), max=1.0))     # torch.nn.utils.clip_grad_norm_ formula
        self.optimizer.step()
        return loss


class Trainer:
    """Drop-in for the trainer class of tools/train_stand.py (constructor keywords :79-88, ``train()`` :90)."""

    def __init__(self, dist, rank, config, resume, only_validation, model, loss_function, optimizer, train_dataloader,
                 validation_dataloader):
        if not torch.cuda.is_available():
            raise RuntimeError("cruse_b200.trainer.Trainer: no CUDA device -- the hot path runs on sm_100a only (no CPU fallback)")
        self.dist, self.rank = dist, rank
        self.device = torch.device("cuda", rank)
        torch.cuda.set_device(self.device)
        self.model = model.to(self.device)                               # base_trainer.py:31 (no DDP wrapper, see the module docstring)
        self.optimizer, self.loss_function = optimizer, loss_function
        self.train_dataloader, self.validation_dataloader = train_dataloader, validation_dataloader
        self.use_amp = config["meta"]["use_amp"]                         # :41-42; the kernels compute in fp32 / tf32: the scaler stays disabled
        self.scaler = torch.amp.GradScaler("cuda", enabled=False)
        self.acoustic_config = config["acoustics"]                       # :45-51
        self.n_fft, self.hop = self.acoustic_config["n_fft"], self.acoustic_config["hop_length"]
        if self.acoustic_config.get("win_length", self.n_fft) != self.n_fft:
            raise RuntimeError("Trainer: win_length must equal n_fft (feature.py:22-30 pads shorter windows; not built)")
        self.train_config = config["trainer"]["train"]                   # :70-77
        self.epochs = self.train_config["epochs"]
        self.save_checkpoint_interval = self.train_config["save_checkpoint_interval"]
        self.clip_grad_norm_value = self.train_config["clip_grad_norm_value"]
        assert self.save_checkpoint_interval >= 1, "Check the 'save_checkpoint_interval' parameter in the config. It should be large than one."
        self.validation_config = config["trainer"]["validation"]        # :80-85
        self.validation_interval = self.validation_config["validation_interval"]
        self.save_max_metric_score = self.validation_config["save_max_metric_score"]
        assert self.validation_interval >= 1, "Check the 'validation_interval' parameter in the config. It should be large than one."
        self.start_epoch = 1                                             # :91-96
        self.best_score = -float("inf") if self.save_max_metric_score else float("inf")
        self.save_dir = Path(config["meta"]["save_dir"]).expanduser().absolute() / config["meta"]["experiment_name"]
        self.checkpoints_dir = self.save_dir / "checkpoints"
        self.only_validation = only_validation
        self.history = []                                                # (epoch, mean training loss, validation score | None)
        if resume:
            self.start_epoch, self.best_score = checkpoint.resume_checkpoint(self.checkpoints_dir, self.model, self.optimizer,
                                                                             self.scaler, map_location="cpu")
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            distrib.broadcast_model(self.model)                          # what DDP's constructor does (:31)
        self._step = None

    # ---- the step: captured wo_male train step when the loss is the path's own, eager autograd otherwise -----------------
    def _loss_kind(self):
        return getattr(self.loss_function, "cruse_kind", "time")

    def _train_batch(self, noisy, clean):
        noisy, clean = noisy.to(self.device, non_blocking=True), clean.to(self.device, non_blocking=True)
        if self._loss_kind() == "wo_male":
            if self._step is None or tuple(self._step.captured.noisy.shape) != tuple(noisy.shape):
                self._step = TrainStep(self.model, noisy.shape[0], noisy.shape[1], max_grad_norm=self.clip_grad_norm_value or None,
                                       n_fft=self.n_fft, hop=self.hop, optimizer=self.optimizer)
            return self._step.step(noisy, clean)
        from . import autograd as ag
        from .acoustics import stft_frames
        self.optimizer.zero_grad(set_to_none=True)                       # :  loss on the enhanced WAVEFORM (e.g. si_snr_loss, train_base/loss.py:7-25)
        X, mag = stft_frames(noisy, self.n_fft, self.hop, self.n_fft, "reflect", mag_bins=self.model.in_feat, mag_eps=1e-8)
        mask = ag.unet2_frames_autograd(self.model, mag)
        wav = ag.mask_istft_apply(mask, X, self.n_fft, self.hop, noisy.shape[-1])
        loss = self.loss_function(wav, clean)
        loss.backward()
        params = [p for p in self.model.parameters() if p.grad is not None]
        distrib.sync_grad(params)
        if self.clip_grad_norm_value:
            torch.nn.utils.clip_grad_norm_(params, self.clip_grad_norm_value)      # base_trainer.py clip_grad_norm_value
        self.optimizer.step()
        return loss.detach()

    def _train_epoch(self, epoch):
        total, n = torch.zeros((), device=self.device), 0
        for batch in self.train_dataloader:
            total += self._train_batch(batch[0], batch[1])
            n += 1
        mean = float(total) / max(1, n)
        if self.rank == 0:
            print(f"[epoch {epoch}] train loss {mean:.6f} over {n} batches")
        self._last_train_loss = mean
        return mean

    @torch.no_grad()
    def _validation_epoch(self, epoch):
        """metric score of the epoch: the NEGATIVE mean validation loss of the hot path (forward + wo_male) -- higher is better,
        which is what ``save_max_metric_score`` = true expects (the reference scores STOI / PESQ here, base_trainer.py:330-376;
        those packages are outside the path and absent from the image)."""
        from . import pipeline
        total, n = 0.0, 0
        for batch in self.validation_dataloader:
            noisy, clean = batch[0].to(self.device), batch[1].to(self.device)
            total += float(pipeline.forward_loss(self.model, noisy, clean, self.n_fft, self.hop)[0])
            n += 1
        score = -total / max(1, n)
        if self.rank == 0:
            print(f"[epoch {epoch}] validation score {score:.6f} over {n} clips")
        return score if self.save_max_metric_score else -score

    def _is_best_epoch(self, score):                                     # base_trainer.py:234-245
        better = score > self.best_score if self.save_max_metric_score else score < self.best_score
        if better:
            self.best_score = score
        return better

    def _save(self, epoch, is_best_epoch=False):
        checkpoint.save_checkpoint(self.checkpoints_dir, epoch, self.model, self.optimizer, self.best_score, self.scaler, is_best_epoch)

    def train(self):                                                     # base_trainer.py:378-424
        for epoch in range(self.start_epoch, self.epochs + 1):
            if self.rank == 0:
                print(f"{'=' * 15} {epoch} epoch {'=' * 15}")
                print("[0 seconds] Begin training...")
            if self.only_validation and self.rank == 0:
                self.model.eval()
                if self._is_best_epoch(self._validation_epoch(epoch)):
                    self._save(epoch, is_best_epoch=True)
                continue
            t0 = time.time()
            self.model.train()
            train_loss = self._train_epoch(epoch)
            score = None
            if self.rank == 0 and self.save_checkpoint_interval != 0 and epoch % self.save_checkpoint_interval == 0:
                self._save(epoch)
            if self.rank == 0 and epoch % self.validation_interval == 0:
                print(f"[{int(time.time() - t0)} seconds] Training has finished, validation is in progress...")
                self.model.eval()
                score = self._validation_epoch(epoch)
                if self._is_best_epoch(score):
                    self._save(epoch, is_best_epoch=True)
            self.history.append((epoch, train_loss, score))
            print(f"[{int(time.time() - t0)} seconds] This epoch is finished.")

"""The direct caller of the hot path in training: one optimisation step in the shape of the
reference's trainer loop -- ``loss.backward()``, gradient clipping, ``optimizer.step()``
(train_base/trainer/base_trainer.py:378-430; DDP gradient averaging :31; Adam from tools/train_stand.py:68) -- built on the
captured train step, so that per batch the host issues one graph replay, one in-place all_reduce on the flat gradient buffer
(N > 1), one norm + scale on that buffer and one fused Adam launch.

Only the step is here: epochs, validation, checkpoints and logging stay with the reference's trainer shell.
"""
from __future__ import annotations

import torch

from . import distrib
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

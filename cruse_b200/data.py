"""Synthetic clip datasets in the shape the reference's launcher expects (``initialize_module(config["train_dataset"]["path"],
args=...)`` -> a ``torch.utils.data.Dataset``, tools/train_stand.py:44-63): item i = (noisy [L], clean [L]) with
clean, noise = 0.05 * randn, noisy = clean + noise, reproducible per item.  The validation flavour also returns
a name, as the reference's validation / inference loops unpack it (train_base/inferencer/base_inferencer.py:171-173)."""
from __future__ import annotations

import torch
from torch.utils.data import Dataset


class SyntheticDataset(Dataset):
    def __init__(self, n_items=64, length=64000, seed=20260, with_name=False):
        self.n_items, self.length, self.seed, self.with_name = int(n_items), int(length), int(seed), bool(with_name)

    def __len__(self):
        return self.n_items

    def __getitem__(self, i):
        if not 0 <= i < self.n_items:
            raise IndexError(i)
        g = torch.Generator(device="cpu").manual_seed(self.seed + i)
        clean = 0.05 * torch.randn(self.length, generator=g)
        noisy = clean + 0.05 * torch.randn(self.length, generator=g)
        return (noisy, clean, f"synthetic_{i:05d}") if self.with_name else (noisy, clean)

"""Checkpoint wire format of the reference trainer / inferencer, so that checkpoints flow both
ways between the reference shell and this package:

* ``latest_model.tar`` / ``best_model.tar``: ``torch.save`` of a dict with the keys ``epoch``, ``best_score``, ``optimizer``,
  ``scaler``, ``model`` (train_base/trainer/base_trainer.py:186-232; the model entry is the bare module's ``state_dict()``
  even under DistributedDataParallel, :203-206);
* ``model_<epoch:04d>.pth``: the model ``state_dict()`` alone (:214-217);
* resume: ``start_epoch = epoch + 1``, optimizer / scaler state restored, model loaded strictly (:149-176);
* preload / inference: ``load_state_dict(checkpoint["model"], strict=False)`` from a ``.tar`` (:128-147,
  train_base/inferencer/base_inferencer.py:128-134).

Pure host code: ``unet_2``'s parameter names and shapes equal the REPAIRED reference module (the oracle's
``unet_2``), which is what makes the files interchangeable -- the literal reference constructor builds ``bn1_t`` and no
``conv*_t`` (model/cruse_net.py:138-142), so a file saved from it loads only with ``strict=False``, as the reference does.

Files are read with ``torch.load(weights_only=True)``: the wire format needs tensors, ints, floats and (for ``best_score``) a
numpy scalar, nothing that has to be unpickled as code.  ``unsafe=True`` restores torch's old arbitrary-pickle behaviour for
files written by a trainer that stored other objects; only use it on files you trust.
"""
from __future__ import annotations

from pathlib import Path

import torch


def _load(path, map_location, unsafe=False):
    if unsafe:
        return torch.load(path, map_location=map_location, weights_only=False)
    import numpy as np
    safe = [np.dtype, np.ndarray]
    try:                                        # numpy scalars (best_score = np.float64(...) in the reference's validation)
        from numpy._core.multiarray import _reconstruct, scalar
        safe += [_reconstruct, scalar]
    except Exception:  # noqa: BLE001
        pass
    safe += [type(np.dtype(t)) for t in ("float32", "float64", "int64", "int32")]
    with torch.serialization.safe_globals(safe):
        return torch.load(path, map_location=map_location, weights_only=True)


def _bare(model):
    return model.module if isinstance(model, torch.nn.parallel.DistributedDataParallel) else model


class _NoScaler:
    """stands in for torch.cuda.amp.GradScaler when training runs without AMP: the key is always present in the file"""

    def state_dict(self):
        return {}

    def load_state_dict(self, sd):
        pass


def save_checkpoint(checkpoints_dir, epoch, model, optimizer, best_score, scaler=None, is_best_epoch=False):
    """base_trainer.py:186-232 -> list of the files written."""
    d = Path(checkpoints_dir).expanduser().absolute()
    d.mkdir(parents=True, exist_ok=True)
    state = {
        "epoch": int(epoch),
        "best_score": float(best_score),       # a plain float: loads under weights_only=True everywhere
        "optimizer": optimizer.state_dict(),
        "scaler": (scaler or _NoScaler()).state_dict(),
        "model": _bare(model).state_dict(),
    }
    written = [d / "latest_model.tar", d / f"model_{str(int(epoch)).zfill(4)}.pth"]
    torch.save(state, written[0].as_posix())                     # everything, overwritten every epoch
    torch.save(state["model"], written[1].as_posix())            # the model alone
    if is_best_epoch:
        written.append(d / "best_model.tar")
        torch.save(state, written[-1].as_posix())
    return written


def resume_checkpoint(checkpoints_dir, model, optimizer, scaler=None, map_location="cpu", unsafe=False):
    """base_trainer.py:149-176 -> (start_epoch, best_score)."""
    path = Path(checkpoints_dir).expanduser().absolute() / "latest_model.tar"
    if not path.exists():
        raise FileNotFoundError(f"{path} does not exist, can not load latest checkpoint.")
    ckpt = _load(path.as_posix(), map_location, unsafe)
    for key in ("epoch", "best_score", "optimizer", "scaler", "model"):
        if key not in ckpt:
            raise RuntimeError(f"{path}: not a trainer checkpoint (missing key {key!r})")
    optimizer.load_state_dict(ckpt["optimizer"])
    (scaler or _NoScaler()).load_state_dict(ckpt["scaler"])
    _bare(model).load_state_dict(ckpt["model"])
    return ckpt["epoch"] + 1, ckpt["best_score"]


def preload_model(model_path, model, map_location="cpu", unsafe=False):
    """base_trainer.py:128-147 / base_inferencer.py:128-134: weights from a ``.tar`` (key ``model``) or a bare ``.pth``,
    ``strict=False`` as in the reference; returns the (missing, unexpected) key lists."""
    path = Path(model_path).expanduser().absolute()
    if not path.exists():
        raise FileNotFoundError(f"The file {path.as_posix()} is not exist. please check path.")
    ckpt = _load(path.as_posix(), map_location, unsafe)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict) else ckpt
    res = _bare(model).load_state_dict(sd, strict=False)
    return list(res.missing_keys), list(res.unexpected_keys)

"""Run the reference's launcher UNCHANGED against this package ("unchanged" = file bytes unchanged, environment
shimmed).  ``tools/train_stand.py:23-90 entry()`` of the reference is executed verbatim with

  * an import hook that decodes the reference's GBK-encoded sources (train_base/utils.py, ... have no coding cookie),
  * ``utils.logger.init`` replaced (the reference's writes to ``_file = None``, utils/logger.py:31-40),
  * ``train_base.loss.wo_male_loss`` added (the launcher resolves the loss by name in that module, :73-75),
  * a config dict whose ``model.path`` / ``trainer.path`` / ``*_dataset.path`` name THIS package's classes.

Needs /root/reference (build container).  With a CUDA device the two epochs really train; without one the run must end in
``cruse_b200.trainer.Trainer``'s loud "no CUDA device" error AFTER the launcher has built datasets, loaders, model, optimizer and loss.
   python tools/train_stand_harness.py <save_dir> [loss name]        -> prints HARNESS: ... lines"""
import importlib.machinery
import os
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _GbkLoader(importlib.machinery.SourceFileLoader):
    def source_to_code(self, data, path, *, _optimize=-1):
        if isinstance(data, (bytes, bytearray)):
            try:
                data = bytes(data).decode("utf-8")
            except UnicodeDecodeError:
                data = bytes(data).decode("gbk")
        return super().source_to_code(data, path, _optimize=_optimize)


def _hook(path):
    if not os.path.abspath(path).startswith(REF):
        raise ImportError
    return importlib.machinery.FileFinder(path, (_GbkLoader, [".py"]))


def main():
    save_dir = sys.argv[1]
    loss_name = sys.argv[2] if len(sys.argv) > 2 else "si_snr_loss"
    sys.path_hooks.insert(0, _hook)
    sys.path_importer_cache.clear()
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import utils.logger as ref_logger                      # the reference's module (valid UTF-8)
    ref_logger.init = lambda filename, run_name, slack_url=None: None
    import train_base.loss as ref_loss                     # the reference's module; the launcher looks the loss up here by name
    from cruse_b200.loss import wo_male_loss
    ref_loss.wo_male_loss = wo_male_loss
    import tools.train_stand as launcher                   # the reference's file, unchanged
    assert os.path.abspath(launcher.__file__).startswith(REF)
    from cruse_b200 import trainer as our_trainer
    seen = {}
    orig_init = our_trainer.Trainer.__init__

    def spy(self, **kw):
        seen.update({k: type(v).__module__ + "." + type(v).__name__ for k, v in kw.items()})
        print("HARNESS: trainer kwargs", sorted(kw), flush=True)
        print("HARNESS: model", seen["model"], "optimizer", seen["optimizer"], "train_dataloader", seen["train_dataloader"], flush=True)
        return orig_init(self, **kw)

    our_trainer.Trainer.__init__ = spy
    config = {
        "meta": {"seed": 0, "use_amp": False, "save_dir": save_dir, "experiment_name": "harness"},
        "acoustics": {"sr": 16000, "n_fft": 512, "hop_length": 320, "win_length": 512},
        "train_dataset": {"path": "cruse_b200.data.SyntheticDataset", "args": {"n_items": 8, "length": 6400},
                          "dataloader": {"batch_size": 4, "num_workers": 0}},
        "validation_dataset": {"path": "cruse_b200.data.SyntheticDataset", "args": {"n_items": 2, "length": 6400, "seed": 7, "with_name": True}},
        "model": {"path": "cruse_b200.cruse_net.unet_2", "args": {"in_feat": 256}},
        "optimizer": {"lr": 1e-3, "beta1": 0.9, "beta2": 0.999},
        "loss_function": {"name": loss_name, "args": {}},
        "trainer": {"path": "cruse_b200.trainer.Trainer",
                    "train": {"epochs": 2, "save_checkpoint_interval": 1, "clip_grad_norm_value": 10.0, "alpha": 0},
                    "validation": {"validation_interval": 1, "save_max_metric_score": True}, "visualization": {}},
    }
    try:
        launcher.entry(0, 1, config, False, False)
        print("HARNESS: entry() returned; checkpoints:", sorted(os.listdir(os.path.join(save_dir, "harness", "checkpoints"))), flush=True)
    except RuntimeError as e:
        print("HARNESS: RuntimeError:", e, flush=True)
        raise SystemExit(3)


if __name__ == "__main__":
    main()

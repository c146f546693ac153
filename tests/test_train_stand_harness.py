"""CPU, build container only (needs /root/reference): the reference's launcher tools/train_stand.py, file unchanged, resolves and
constructs THIS package's model / trainer / dataset classes through its own ``initialize_module`` and hands them the keyword arguments
of tools/train_stand.py:79-88 (SURVEY.md App. A.5 harness).  There is no GPU here, so the run must stop in the trainer's loud
"no CUDA device" error -- after the launcher has done everything in front of it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tools"), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("loss_name", ["si_snr_loss", "wo_male_loss"])
def test_reference_launcher_runs_unchanged_against_our_classes(tmp_path, loss_name):
    import torch
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_stand_harness.py"), str(tmp_path), loss_name],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert "HARNESS: trainer kwargs ['config', 'dist', 'loss_function', 'model', 'only_validation', 'optimizer', 'rank', 'resume', " \
           "'train_dataloader', 'validation_dataloader']" in out, out[-3000:]
    assert "HARNESS: model cruse_b200.cruse_net.unet_2 optimizer torch.optim.adam.Adam train_dataloader torch.utils.data.dataloader.DataLoader" in out
    if torch.cuda.is_available():
        assert r.returncode == 0 and "latest_model.tar" in out, out[-3000:]
    else:
        assert r.returncode == 3 and "no CUDA device" in out, out[-3000:]

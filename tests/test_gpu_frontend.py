"""GPU: the step in front of the path on the device against the oracle's restatements, which tests/test_oracle.py pins
to the reference's own source (train_base/model/base_model.py:202-300, dataset/dataset.py:236-260), and against the committed
reference outputs directly."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("norm", ["offline_laplace_norm", "cumulative_laplace_norm", "offline_gaussian_norm", "cumulative_layer_norm"])
def test_feature_norms(cuda, golden_dir, norm):
    from cruse_b200 import frontend
    from oracle import cruse_oracle as o
    g = np.load(os.path.join(golden_dir, "refx_frontend.npz"))
    x = torch.from_numpy(g["x"])                                             # the reference's [B,1,F,T]
    assert rel_err(frontend.feature_norm(x.to(cuda), norm), torch.from_numpy(g[norm])) <= 1e-5
    torch.manual_seed(3)
    mag = torch.rand(5, 1, 256, 501) + 0.01                                  # one 10 s clip geometry
    want = getattr(o, norm)(mag)
    got = frontend.feature_norm(mag[:, 0].transpose(1, 2).contiguous().to(cuda), norm)        # frame-major [B,T,F] in and out
    assert rel_err(got.transpose(1, 2), want[:, 0]) <= 2e-5
    with pytest.raises(NotImplementedError):
        frontend.feature_norm(mag.to(cuda), "forgetting_norm")


def test_snr_mix_with_rir_matches_reference_and_oracle(cuda, golden_dir):
    from cruse_b200 import frontend
    from oracle import cruse_oracle as o
    g = np.load(os.path.join(golden_dir, "refx_frontend.npz"))
    c, n, rir = (torch.from_numpy(g[k]).float() for k in ("mix_clean_in", "mix_noise_in", "mix_rir"))
    noisy, clean = frontend.snr_mix(c[None].to(cuda), n[None].to(cuda), 5.0, rir=rir.to(cuda))
    assert rel_err(clean[0], torch.from_numpy(g["mix_clean"])) <= 1e-4 and rel_err(noisy[0], torch.from_numpy(g["mix_noisy"])) <= 1e-4
    # a batch with per-item SNR / level and a long impulse response, against the oracle
    torch.manual_seed(9)
    B, L, R = 3, 16000, 2500
    cb, nb = torch.randn(B, L), torch.randn(B, L)
    rirs = torch.exp(-torch.arange(R) / 400.0) * torch.randn(B, R)
    snr, lvl = torch.tensor([-5.0, 0.0, 12.0]), torch.tensor([-30.0, -25.0, -20.0])
    got_n, got_c = frontend.snr_mix(cb.to(cuda), nb.to(cuda), snr, lvl, rir=rirs.to(cuda))
    for b in range(B):
        wn, wc = o.snr_mix(cb[b], nb[b], float(snr[b]), float(lvl[b]), rir=rirs[b])
        assert rel_err(got_n[b], wn) <= 1e-4 and rel_err(got_c[b], wc) <= 1e-4
        rms = float((got_n[b] ** 2).mean().sqrt())
        assert abs(20 * np.log10(rms) - float(lvl[b])) <= 1e-3


def test_pcm16_host_buffers_through_the_prefetch_entry(cuda):
    """16-bit PCM host buffers (the wav files' sample format) through CapturedForwardLoss.prefetch: widened on the device as
    soundfile / librosa widen them on the host (x / 32768, dataset/dataset.py:20) -- bit-identical to feeding the float32 values."""
    from cruse_b200 import ops, pipeline
    from cruse_b200.cruse_net import unet_2
    g = torch.Generator().manual_seed(9)
    pcm = torch.randint(-20000, 20000, (2, 16003), generator=g, dtype=torch.int16)
    raw, out = pcm.to(cuda), torch.empty(2, 16003, device=cuda)
    ops.pcm16_to_float(raw, out)
    assert torch.equal(out.cpu(), pcm.float() / 32768.0)
    model = unet_2(in_feat=256).to(cuda).eval()
    B, L = 3, 32000
    n16 = torch.randint(-3000, 3000, (B, L), generator=g, dtype=torch.int16)
    c16 = torch.randint(-2000, 2000, (B, L), generator=g, dtype=torch.int16)
    cap = pipeline.CapturedForwardLoss(model, B, L, 512, 320)
    l_pcm = float(cap.run_prefetched(cap.prefetch(n16.pin_memory(), c16.pin_memory()))[0])
    l_f32 = float(cap.run_prefetched(cap.prefetch((n16.float() / 32768.0).pin_memory(), (c16.float() / 32768.0).pin_memory()))[0])
    assert l_pcm == l_f32

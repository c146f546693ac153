"""Generate tests/golden/*.npz  (run in the BUILD container only; needs /root/reference).

TEST INFRASTRUCTURE.  Two families of fixtures:

ref_*.npz   outputs of the pieces of the UNMODIFIED reference that import and run here
            (SURVEY.md section 8c "What importably runs"): Conv2dNormAct, GroupedGRULayer
            (model/based_model/cust_conv.py:15-62, :250-325), complex_mul
            (train_base/acoustics/mask.py:60-62), si_snr_loss (train_base/loss.py:7-25).
            They pin the oracle's causal strided conv+BN+ReLU stage, grouped GRU with
            state carry, complex mask-apply and SI-SNR against reference code.
oracle_*.npz small end-to-end outputs of oracle/cruse_oracle.py (weights regenerated from
            the seed, only inputs/outputs + a weight checksum are stored) so that the GPU
            box, which has no /root/reference, checks against committed numbers.

Usage:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def sd_np(mod):
    return {k: v.detach().numpy() for k, v in mod.state_dict().items()}


def ref_fragments():
    sys.path.insert(0, REF)
    from model.based_model.cust_conv import Conv2dNormAct, GroupedGRULayer  # noqa: E402
    from train_base.acoustics.mask import complex_mul                       # noqa: E402
    from train_base.loss import si_snr_loss                                 # noqa: E402

    torch.manual_seed(7)
    # causal (2,3) conv, freq stride 2, BN, ReLU  == encoder stage semantics
    # train-mode (batch statistics) output for two stage shapes
    for name, (cin, cout, F) in {"a": (1, 8, 161), "b": (8, 16, 128)}.items():
        m = Conv2dNormAct(cin, cout, (2, 3), fstride=2)
        m[2].weight.data.copy_(1 + 0.1 * torch.randn(cout))
        m[2].bias.data.copy_(0.1 * torch.randn(cout))
        x = torch.randn(2, cin, 10, F)
        params = {"sd." + k: v.copy() for k, v in sd_np(m).items() if "running" not in k and "num_batches" not in k}
        m.train()
        with torch.no_grad():
            y_train = m(x)
        np.savez_compressed(os.path.join(OUT, f"ref_conv2dnormact_train_{name}.npz"), x=x.numpy(),
                            y_train=y_train.numpy(), **params)
    # eval-mode (running statistics) output
    torch.manual_seed(11)
    m = Conv2dNormAct(1, 8, (2, 3), fstride=2).eval()
    m[2].running_mean.copy_(0.1 * torch.randn(8))
    m[2].running_var.copy_(1 + 0.1 * torch.rand(8))
    x = torch.randn(2, 1, 10, 161)
    with torch.no_grad():
        y = m(x)
    np.savez_compressed(os.path.join(OUT, "ref_conv2dnormact_eval.npz"), x=x.numpy(), y=y.numpy(),
                        **{"sd." + k: v for k, v in sd_np(m).items()})

    # grouped GRU with explicit state (streaming API, cust_conv.py:303-325)
    torch.manual_seed(13)
    g = GroupedGRULayer(32, 32, 4)
    x = torch.randn(3, 9, 32)
    h0 = 0.3 * torch.randn(4, 3, 8)
    with torch.no_grad():
        y, h = g(x, h0)
        y0, hz = g(x)
    np.savez_compressed(os.path.join(OUT, "ref_groupedgru.npz"), x=x.numpy(), h0=h0.numpy(), y=y.numpy(), h=h.numpy(),
                        y_zero=y0.numpy(), h_zero=hz.numpy(), **{"sd." + k: v for k, v in sd_np(g).items()})

    torch.manual_seed(17)
    a, b, c, d = (torch.randn(2, 5, 33) for _ in range(4))
    r, i = complex_mul(a, b, c, d)
    s1, s2 = torch.randn(3, 800), torch.randn(3, 800)
    np.savez_compressed(os.path.join(OUT, "ref_misc.npz"), a=a.numpy(), b=b.numpy(), c=c.numpy(), d=d.numpy(),
                        r=r.numpy(), i=i.numpy(), s1=s1.numpy(), s2=s2.numpy(),
                        si_snr=si_snr_loss()(s1, s2).numpy())


def oracle_vectors():
    from oracle import cruse_oracle as o
    for tag, (F, n_fft, hop, L, B) in {"B": (256, 512, 320, 3200, 2), "R": (161, 320, 160, 1600, 2)}.items():
        for act in ("relu", "prelu"):
            m = o.make_model(F, act=act).eval()
            csum = float(sum(p.double().abs().sum() for p in m.state_dict().values()))
            noisy, clean = o.synth_batch(B, L)
            with torch.no_grad():
                loss, wav, est, mask = o.forward_loss(m, noisy, clean, n_fft, hop)
            np.savez_compressed(os.path.join(OUT, f"oracle_fwd_{tag}_{act}.npz"), noisy=noisy.numpy(), clean=clean.numpy(),
                                loss=loss.numpy(), wav=wav.numpy(), est=est.numpy(), mask=mask.numpy(),
                                weight_abs_sum=np.float64(csum), n_fft=n_fft, hop=hop, F=F)
    # wo_male known-answer on a tiny hand-checkable case
    ref = torch.tensor([[[[3.0, 0.0]], [[4.0, 1.0]]]])     # [1,2,1,2]  mags 5, 1
    est = torch.tensor([[[[0.0, 1.0]], [[0.0, 0.0]]]])     # mags 0, 1
    unp = torch.tensor([[[[5.0, 0.0]], [[0.0, 2.0]]]])     # mags 5, 2
    np.savez_compressed(os.path.join(OUT, "oracle_wo_male_kat.npz"), ref=ref.numpy(), est=est.numpy(), unproc=unp.numpy(),
                        loss=o.wo_male(ref, est, unp).numpy())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref_fragments()
    oracle_vectors()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))

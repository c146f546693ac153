"""Run a few steps of the bench workload (for ncu): python tools/prof_step.py [steps] [B] [seconds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import F_BINS, HOP, N_FFT, SR, randomise_bn, synth_batch  # noqa: E402
from cruse_b200 import pipeline  # noqa: E402
from cruse_b200.cruse_net import unet_2  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
dev = torch.device("cuda:0")
torch.manual_seed(1234)
m = unet_2(in_feat=F_BINS)
randomise_bn(m)
m = m.to(dev).eval()
noisy, clean = synth_batch(B, int(secs * SR), 20260)
noisy, clean = noisy.to(dev), clean.to(dev)
with torch.no_grad():
    for _ in range(steps):
        loss = pipeline.forward_loss(m, noisy, clean, N_FFT, HOP)[0]
torch.cuda.synchronize()
print("loss", float(loss))

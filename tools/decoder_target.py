"""Developer tool: the fused decoder alone (32 utterances x T frames), for ncu captures and timing.
   python tools/decoder_target.py [T] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cruse_b200 import ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 63
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B = 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
chans, freqs = [64, 32, 16, 8, 1], [16, 32, 64, 128, 256]
y2 = torch.randn(B, T, 1024, device=dev)
g, b_ = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
skips = [0.5 * torch.randn(B, T, chans[k], freqs[k], device=dev) for k in range(4)]
ws = [torch.randn(chans[k], chans[k + 1], 1, 3, device=dev) / (1.5 * chans[k]) ** 0.5 for k in range(4)]
bs = [0.1 * torch.randn(chans[k + 1], device=dev) for k in range(4)]
scs = [torch.ones(chans[k + 1], device=dev) for k in range(3)]
shs = [torch.zeros(chans[k + 1], device=dev) for k in range(3)]
mask = torch.zeros(B, T, 256, device=dev)
flush = torch.empty(64 << 20, device=dev)
SKC = len(sys.argv) > 3 and sys.argv[3] == "skc"
wsk4, wsk3 = torch.randn(64, 64, 1, 3, device=dev) / 192 ** 0.5, torch.randn(32, 32, 1, 3, device=dev) / 96 ** 0.5
image = ops.decoder_fused_prep(ws, bs, scs, shs, None, "relu", wsk4 if SKC else None, wsk3 if SKC else None)
if SKC:
    skips[0] = skips[0].transpose(0, 1).contiguous()
for _ in range(3):
    ops.decoder_fused_range(y2, g, b_, 1e-5, skips, image, mask, 0, T, skip_convs=SKC)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.decoder_fused_range(y2, g, b_, 1e-5, skips, image, mask, 0, T, skip_convs=SKC)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"decoder_fused B={B} T={T}: median {ts[len(ts)//2]:.1f} us, min {ts[0]:.1f} us  ({B*T} frames, {4*B*T*(5*1024+256)/ts[len(ts)//2]/1e3:.0f} GB/s algorithmic)")

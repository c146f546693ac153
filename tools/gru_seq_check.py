"""GPU check + timing of the two recurrence kernels: python tools/gru_seq_check.py [B] [T]"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cruse_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 501
G, H = 4, 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
grus = [nn.GRU(H, H, 1, batch_first=True) for _ in range(G)]
x = torch.randn(B, T, G * H)
with torch.no_grad():
    ref = torch.cat([grus[g](x[..., g * H:(g + 1) * H].contiguous())[0] for g in range(G)], dim=-1)
P = lambda k: [getattr(g, k).detach().to(dev) for g in grus]
w_ih, w_hh, b_ih, b_hh = P("weight_ih_l0"), P("weight_hh_l0"), P("bias_ih_l0"), P("bias_hh_l0")
xproj = ops.gru_ih_gemm(x.view(B * T, G * H).to(dev), w_ih, b_ih, b_hh, mode="fp32")
from cruse_b200 import lib  # noqa: E402
print("max co-resident clusters (H=256):", lib().cruse_gru_seq_tc_max_clusters(256), " needed:", G * ((B + 15) // 16))
for mode in ("fp32", "tf32"):
    y = ops.gru_seq_fwd(xproj, w_hh, b_hh, B, T, interleave=False, mode=mode)
    torch.cuda.synchronize()
    err = float((y.cpu() - ref).abs().max() / ref.abs().max())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.gru_seq_fwd(xproj, w_hh, b_hh, B, T, interleave=False, mode=mode)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"gru_seq[{mode}] B={B} T={T}: rel err {err:.3e}  {ms:.3f} ms  {1e3 * ms / T:.3f} us/step", flush=True)

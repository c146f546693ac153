"""Developer tool: the GRU input-projection GEMM, A-stationary kernel vs one-tile-per-CTA kernel (cruse_gemm_set_astat), alone on
the GPU with the L2 flushed: time per launch, algorithmic GB/s (x in + xproj out), max error against an fp32 torch matmul.
   python tools/ih_gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cruse_b200 import ops
from cruse_b200._lib import lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
G, H = 4, 256
w = [0.06 * torch.randn(3 * H, H, device=dev) for _ in range(G)]
bi = [0.1 * torch.randn(3 * H, device=dev) for _ in range(G)]
bh = [0.1 * torch.randn(3 * H, device=dev) for _ in range(G)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for M in (63 * 32, 31 * 32, 501 * 32, 201 * 64, 100):
    x = torch.randn(M, G * H, device=dev)
    ref = torch.stack([x[:, g * H:(g + 1) * H].double() @ w[g].double().t() + bi[g].double() +
                       torch.cat([bh[g][:2 * H], torch.zeros(H, device=dev)]).double() for g in range(G)], dim=1)
    for mode in (0, 1):
        lib().cruse_gemm_set_astat(mode)
        out = ops.gru_ih_gemm(x, w, bi, bh, mode="tf32")
        torch.cuda.synchronize()
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.gru_ih_gemm(x, w, bi, bh, mode="tf32"); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        nbytes = 4 * (x.numel() + out.numel())
        print(f"M={M:6d} astat={mode}: median {1e3 * ts[len(ts) // 2]:7.1f} us  min {1e3 * ts[0]:7.1f} us  {nbytes / ts[len(ts) // 2] / 1e6:7.1f} GB/s  rel err {err:.2e}")
lib().cruse_gemm_set_astat(1)

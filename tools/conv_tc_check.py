"""developer check: tensor-core conv stages (conv_tc.cu) against the exact-fp32 CUDA-core kernels, + timing.
   python tools/conv_tc_check.py [B T]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cruse_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(7)


def rnd(*s, scale=1.0):
    return (scale * torch.randn(*s, generator=g)).to(dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run(B, T, timing):
    worst = 0.0
    convs = [(2, 2, 8, 16, 128), (2, 2, 16, 32, 64), (2, 2, 32, 64, 32), (1, 1, 8, 8, 128), (1, 1, 16, 16, 64), (1, 1, 32, 32, 32), (1, 1, 64, 64, 16)]
    for kt, fs, ci, co, fin in convs:
        x = rnd(B, T, ci, fin)
        w = rnd(co, ci, kt, 3, scale=0.2)
        bias = rnd(co, scale=0.1) if kt == 2 else None
        scale = (1 + 0.1 * torch.randn(co, generator=g)).to(dev) if kt == 2 else None
        shift = rnd(co, scale=0.1) if kt == 2 else None
        alpha = (0.25 * torch.rand(co, generator=g)).to(dev) if kt == 2 else None
        act = "prelu" if kt == 2 else "none"
        f = lambda: ops.conv_fwd(x, w, bias, scale, shift, alpha, act, kt, fs)
        ops.set_conv_mode("fp32"); ref = f(); t_ref = timeit(f) if timing else 0
        ops.set_conv_mode("tf32"); out = f(); t_tc = timeit(f) if timing else 0
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        by = (x.numel() + out.numel()) * 4
        print(f"conv{kt}x3 {ci}->{co} F{fin}: rel err {err:.2e}  fp32 {t_ref:.4f} ms  tc {t_tc:.4f} ms  ({by / max(t_tc, 1e-9) / 1e6:.0f} GB/s)", flush=True)
    for ci, co, fin in [(64, 32, 16), (32, 16, 32), (16, 8, 64)]:
        x = rnd(B, T, ci, fin)
        w = rnd(ci, co, 1, 3, scale=0.2)
        bias, shift = rnd(co, scale=0.1), rnd(co, scale=0.1)
        scale = (1 + 0.1 * torch.randn(co, generator=g)).to(dev)
        skip = rnd(B, T, co, 2 * fin)
        f = lambda: ops.convT_fwd(x, w, bias, scale, shift, None, "relu", skip, 2 * fin)
        ops.set_conv_mode("fp32"); ref = f(); t_ref = timeit(f) if timing else 0
        ops.set_conv_mode("tf32"); out = f(); t_tc = timeit(f) if timing else 0
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
        by = (x.numel() + 2 * out.numel()) * 4
        print(f"convT1x3 {ci}->{co} F{fin}: rel err {err:.2e}  fp32 {t_ref:.4f} ms  tc {t_tc:.4f} ms  ({by / max(t_tc, 1e-9) / 1e6:.0f} GB/s)", flush=True)
    return worst


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prof":     # under ncu: the big shapes only, no timing loops
        run(32, 501, False)
        sys.exit(0)
    w1 = run(3, 21, False)
    w2 = run(1, 1, False)
    print("worst rel err (small):", max(w1, w2))
    if len(sys.argv) > 2:
        run(int(sys.argv[1]), int(sys.argv[2]), True)
    else:
        run(32, 501, True)

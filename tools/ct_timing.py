"""per-tile phase stamps of conv_tc CTA 0 (library built with CRUSE_EXTRA_NVCC_FLAGS=-DCRUSE_CT_TIMING)
   python tools/ct_timing.py enc4|skip4|dec4|fused2"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cruse_b200 import ops, lib

which = sys.argv[1] if len(sys.argv) > 1 else "enc4"
dev = torch.device("cuda:0")
B, T = 32, 501
g = torch.Generator().manual_seed(0)
if which == "enc4":
    x = torch.randn(B, T, 32, 32, generator=g).to(dev); w = (0.1 * torch.randn(64, 32, 2, 3, generator=g)).to(dev)
    f = lambda: ops.conv_fwd(x, w, None, None, None, None, "relu", 2, 2)
elif which == "skip4":
    x = torch.randn(B, T, 64, 16, generator=g).to(dev); w = (0.1 * torch.randn(64, 64, 1, 3, generator=g)).to(dev)
    f = lambda: ops.conv_fwd(x, w, None, None, None, None, "none", 1, 1)
elif which == "fused2":
    x = torch.randn(B, T, 8, 128, generator=g).to(dev); w = (0.1 * torch.randn(16, 8, 2, 3, generator=g)).to(dev)
    wsk = (0.1 * torch.randn(8, 8, 1, 3, generator=g)).to(dev)
    ops.set_conv_mode("tf32")
    f = lambda: ops.conv_skip_fwd(x, w, None, None, None, None, "relu", wsk)
else:
    x = torch.randn(B, T, 64, 16, generator=g).to(dev); w = (0.1 * torch.randn(64, 32, 1, 3, generator=g)).to(dev)
    sk = torch.randn(B, T, 32, 32, generator=g).to(dev)
    f = lambda: ops.convT_fwd(x, w, None, None, None, None, "relu", sk, 32)
for _ in range(3):
    f()
torch.cuda.synchronize()
n = 64 * 16
buf = (ctypes.c_longlong * n)()
h = lib()
h.cruse_debug_ct_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert h.cruse_debug_ct_timing(buf, n) == 0
v = [buf[i] for i in range(n)]
t0 = v[63 * 16 + 14]
print("setup cycles:", v[63 * 16 + 13] - t0)
names = ["P0 issued", "P0 slot", "P0 stored", "P1 issued", "P1 slot", "P1 stored", "M wait acc", "M acc free", "M A landed", "M issued", "E wait", "E acc full", "E stored", "E regs", "E stg free", "E staged"]
print("tile | " + " | ".join(names))
for t in range(14):
    row = v[t * 16: t * 16 + 16]
    print(f"{t:4d} | " + " | ".join(f"{(x - t0) if x else 0:9d}" for x in row))

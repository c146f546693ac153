"""Developer tool: per-step phase timing of the tcgen05 recurrence kernel (needs a library built with
CRUSE_EXTRA_NVCC_FLAGS=-DCRUSE_SEQ_TIMING).  python tools/seq_timing.py [B] [T]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cruse_b200 import ops  # noqa: E402
from cruse_b200._lib import LIB_PATH  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 501
G, H = 4, 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
w_hh = [0.06 * torch.randn(3 * H, H, device=dev) for _ in range(G)]
b_hh = [0.06 * torch.randn(3 * H, device=dev) for _ in range(G)]
xproj = torch.randn(B * T, G, 3 * H, device=dev)
for _ in range(3):
    ops.gru_seq_fwd(xproj, w_hh, b_hh, B, T, interleave=False, mode="tf32")
torch.cuda.synchronize()
h = C.CDLL(LIB_PATH)
n = 8 * min(T, 2048)
buf = (C.c_longlong * n)()
assert h.cruse_debug_seq_timing(buf, n) == 0
import numpy as np
a = np.array(buf[:], dtype=np.int64).reshape(-1, 8)[5:T - 5]
names = ["hbar wait (mma thr)", "mma issue+commit", "commit -> acc_full seen (t0)", "tmem ld + sts + syncthreads", "gate math", "st.async x NC", "y store"]
d = [a[:, 1] - a[:, 0], a[:, 2] - a[:, 1], a[:, 3] - a[:, 2], a[:, 4] - a[:, 3], a[:, 7] - a[:, 4], a[:, 5] - a[:, 7], a[:, 6] - a[:, 5]]
for nm, v in zip(names, d):
    print(f"{nm:32s} mean {v.mean():8.1f}  min {v.min():6d}  max {v.max():6d} cycles")
step = a[1:, 0] - a[:-1, 0]
print(f"step period mean {step.mean():.1f} cycles; t0 end-of-step -> next mma-thread start {np.mean(a[1:, 0] - a[:-1, 6]):.1f}")

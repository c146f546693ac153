"""Developer tool: the encoder stages of the head of the inference step alone (cfg-2 shapes), for ncu captures / timing.
   python tools/conv_target.py [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cruse_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, T = 32, 501
dev = torch.device("cuda:0")
torch.manual_seed(0)
ops.set_conv_mode("tf32")
ch, fr = [1, 8, 16, 32, 64], [256, 128, 64, 32, 16]
x = [torch.randn(B, T, ch[k], fr[k], device=dev) for k in range(5)]
w = [torch.randn(ch[k + 1], ch[k], 2, 3, device=dev) * 0.1 for k in range(4)]
b = [torch.randn(ch[k + 1], device=dev) * 0.1 for k in range(4)]
sc = [torch.ones(ch[k + 1], device=dev) for k in range(4)]
sh = [torch.zeros(ch[k + 1], device=dev) for k in range(4)]
wsk = [torch.randn(ch[k], ch[k], 1, 3, device=dev) * 0.1 for k in range(1, 5)]
out = [torch.empty(B, T, ch[k + 1], fr[k + 1], device=dev) for k in range(4)]
osk = [torch.empty(B, T, ch[k], fr[k], device=dev) for k in range(1, 5)]
flush = torch.empty(64 << 20, device=dev)
def run():
    ops.conv_skip_fwd(x[1], w[1], b[1], sc[1], sh[1], None, "relu", wsk[0], out=out[1], out_skip=osk[0])
    ops.conv_skip_fwd(x[2], w[2], b[2], sc[2], sh[2], None, "relu", wsk[1], out=out[2], out_skip=osk[1])
    ops.conv_fwd_range(x[3], w[3], b[3], sc[3], sh[3], None, "relu", 2, 2, B, T, out[3], 0, T)
for _ in range(2):
    run()
torch.cuda.synchronize()
names = ["fused 8->16 + skip1", "fused 16->32 + skip2", "conv 32->64"]
calls = [lambda: ops.conv_skip_fwd(x[1], w[1], b[1], sc[1], sh[1], None, "relu", wsk[0], out=out[1], out_skip=osk[0]),
         lambda: ops.conv_skip_fwd(x[2], w[2], b[2], sc[2], sh[2], None, "relu", wsk[1], out=out[2], out_skip=osk[1]),
         lambda: ops.conv_fwd_range(x[3], w[3], b[3], sc[3], sh[3], None, "relu", 2, 2, B, T, out[3], 0, T)]
for nm, f in zip(names, calls):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    ts.sort()
    print(f"{nm:24s} median {ts[len(ts)//2]:.1f} us")

"""Developer tool: progress of the two flag-synchronised recurrences inside ONE replay of the captured inference step
(cfg-2: 32 x 10 s).  Cluster 0 of each layer stamps %globaltimer every 8th step and around every chunk wait
(cruse_debug_seq_trace); the table shows, per chunk, when each layer waited, for how long, and its us/step in between --
i.e. where layer 2 lags layer 1 and where the steps are stretched by what runs beside them.
   python tools/wavefront_trace.py [out.md]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cruse_b200 import pipeline
from cruse_b200._lib import lib
from cruse_b200.cruse_net import unet_2

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = unet_2(in_feat=256)
bench.randomise_bn(model)
model = model.to(dev).eval()
B, L = 32, 160000
noisy, clean = bench.synth_batch(B, L, 20260)
buf = torch.zeros(2 * 160, dtype=torch.int64, device=dev)
lib().cruse_debug_seq_trace(buf.data_ptr())           # the capture below bakes the trace pointers into the two recurrence launches
cap = pipeline.CapturedForwardLoss(model, B, L, 512, 320, warmup=1)
lib().cruse_debug_seq_trace(None)
cap.noisy.copy_(noisy.to(dev)); cap.clean.copy_(clean.to(dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    flush.zero_(); cap.replay()
torch.cuda.synchronize()
buf.zero_(); flush.zero_()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); cap.replay(); b.record()
torch.cuda.synchronize()
t = buf.cpu().view(2, 160).double()
T = 1 + L // 320
bounds = model.gru.plan(B, T, dev)["bounds"]
t0 = float(t[:, :128][t[:, :128] > 0].min())
out = [f"step {a.elapsed_time(b) * 1e3:.1f} us (graph replay, L2 flushed); times below in us relative to layer 1's first stamped step; chunk bounds {bounds}", "",
       "| layer | chunk | frames | wait begin | wait end | waited | first step | last step | us/step inside |", "|---|---:|---|---:|---:|---:|---:|---:|---:|"]
for lay in range(2):
    steps = t[lay, :128]
    for k in range(len(bounds) - 1):
        wb, we = float(t[lay, 128 + 2 * k]), float(t[lay, 128 + 2 * k + 1])
        i0, i1 = (bounds[k] + 7) // 8, (bounds[k + 1] - 1) // 8
        s0, s1 = float(steps[i0]), float(steps[i1])
        rate = (s1 - s0) / 1e3 / max(1, 8 * (i1 - i0))
        out.append(f"| {lay + 1} | {k} | [{bounds[k]},{bounds[k + 1]}) | {(wb - t0) / 1e3:.1f} | {(we - t0) / 1e3:.1f} | {(we - wb) / 1e3:.1f} | {(s0 - t0) / 1e3:.1f} | {(s1 - t0) / 1e3:.1f} | {rate:.3f} |")
last = [float(t[lay, (T - 1) // 8]) for lay in range(2)]
waits = [sum(float(t[lay, 128 + 2 * k + 1] - t[lay, 128 + 2 * k]) for k in range(1, len(bounds) - 1)) / 1e3 for lay in range(2)]
print(f"SUMMARY env={ {k: v for k, v in os.environ.items() if k.startswith('CRUSE_')} } step {a.elapsed_time(b) * 1e3:.1f} us  L1 end {(last[0] - t0) / 1e3:.1f}  L2 start {(float(t[1, 0]) - t0) / 1e3:.1f}  L2 end {(last[1] - t0) / 1e3:.1f}  lag {(last[1] - last[0]) / 1e3:.1f}  L2 waits after chunk 0: {waits[1]:.1f}")
out.append("")
out.append(f"last stamped step: layer 1 at {(last[0] - t0) / 1e3:.1f} us, layer 2 at {(last[1] - t0) / 1e3:.1f} us (lag {(last[1] - last[0]) / 1e3:.1f} us)")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(txt + "\n")

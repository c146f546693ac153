"""kernel timeline of one inference step (CUPTI through torch.profiler): start / duration / stream of every kernel,
so that overlap between the GRU wavefront streams and the side work is visible.
   python tools/trace_step.py [out.md] [--graph | --train]"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from cruse_b200 import pipeline
from cruse_b200.cruse_net import unet_2

use_graph = "--graph" in sys.argv
out = next((a for a in sys.argv[1:] if not a.startswith("--")), None)
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = unet_2(in_feat=256)
bench.randomise_bn(model)
train = "--train" in sys.argv
model = model.to(dev).train(train)
B, L = (64, 64000) if train else (32, 160000)
params = list(model.parameters())
noisy, clean = bench.synth_batch(B, L, 20260)
noisy, clean = noisy.to(dev), clean.to(dev)
cap = None
if use_graph:
    cap = pipeline.CapturedTrainStep(model, B, L, 512, 320) if train else pipeline.CapturedForwardLoss(model, B, L, 512, 320)
if cap is not None:
    cap.noisy.copy_(noisy); cap.clean.copy_(clean)


def step():
    if cap is not None:
        return cap.replay()
    if train:
        for p in params:
            p.grad = None
        loss = pipeline.train_forward_loss(model, noisy, clean, 512, 320)
        loss.backward()
        return loss.detach()
    with torch.no_grad():
        return pipeline.forward_loss(model, noisy, clean, 512, 320)[0]


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
streams = {}
lines = ["| start us | dur us | end us | stream | kernel |", "|---:|---:|---:|---:|---|"]
for e in ev:
    s = streams.setdefault(e["args"].get("stream"), len(streams))
    name = e["name"].replace("cruse::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    lines.append(f"| {e['ts'] - t0:8.1f} | {e['dur']:7.1f} | {e['ts'] - t0 + e['dur']:8.1f} | {s} | {name[:70]} |")
end = max(e["ts"] + e["dur"] for e in ev) - t0
lines.append(f"\n{len(ev)} kernels, span {end:.1f} us, sum of durations {sum(e['dur'] for e in ev):.1f} us ({'graph replay' if use_graph else 'eager'})")
txt = "\n".join(lines)
print(txt)
if out:
    open(out, "w").write(txt + "\n")

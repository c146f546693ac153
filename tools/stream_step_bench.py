"""BASELINE cfg-5 / SURVEY 8(d): streaming step, 2048 concurrent utterances x 1 frame, persistent state, one GPU.
One step = mag frame [B,1,256] -> mask [B,1,256] through cruse_b200.streaming.step (encoder history frames + both GRU
states carried in StreamState).  Prints one JSON line (frames/s, us per step) and, with --table, the per-launch table.
   python tools/stream_step_bench.py [--utts 2048] [--steps 50] [--table out.md] [--no-graph]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cruse_b200 import ops, streaming
from cruse_b200.cruse_net import unet_2

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=2048)
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--warmup", type=int, default=10)
ap.add_argument("--table", default=None)
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = unet_2(in_feat=256)
bench.randomise_bn(model)
model = model.to(dev).eval()
B = args.utts
g = torch.Generator(device="cpu").manual_seed(20260)
frames = torch.rand(8, B, 1, 256, generator=g).to(dev)          # a few distinct input frames, cycled
state = streaming.StreamState()
for i in range(3):                                              # fill the state (first call allocates it)
    streaming.step(model, frames[i % 8], state)
torch.cuda.synchronize()

# static-buffer variant for graph capture: the state tensors are updated in place after every step
static_in = frames[0].clone()
hist = [h.clone() for h in state.hist]
gru = [h.clone() for h in state.gru]


def step_static():
    st = streaming.StreamState()
    st.hist, st.gru = list(hist), tuple(gru)
    out = streaming.step(model, static_in, st)
    for dst, src in zip(hist, st.hist):
        dst.copy_(src)
    for dst, src in zip(gru, st.gru):
        dst.copy_(src)
    return out


graph = None
if not args.no_graph:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step_static()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = step_static()


def step(i):
    static_in.copy_(frames[i % 8], non_blocking=True)
    if graph is not None:
        graph.replay()
        return static_out
    return step_static()


for i in range(args.warmup):
    step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(args.steps):
    out = step(i)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
prof = ops.Profile(timing=True)
ops.set_profile(prof)
step_static()
ops.set_profile(None)
rows = prof.rows()
line = {"metric": "frames/sec (16 kHz, 20 ms hop) CRUSE streaming step", "value": B / (ms / 1e3), "unit": "frames/s",
        "us_per_step": 1e3 * ms, "config": {"workload": f"cfg5: {B} concurrent utterances x 1 frame, persistent state (encoder history + 2 GRU states)",
                                            "launch": "CUDA graph replay" if graph is not None else "eager"},
        "steps": args.steps, "gpu_launches": len(rows), "finite": bool(torch.isfinite(out).all()),
        "real_time_factor": (ms / 1e3) / 0.02}
print(json.dumps(line))
if args.table:
    with open(args.table, "w") as f:
        f.write("| call | tag | ms | alg GB/s |\n|---|---|---:|---:|\n")
        for n, tag, by, fl, t in rows:
            f.write(f"| {n} | {tag} | {t:.4f} | {by / (t * 1e6) if t > 0 else 0:.1f} |\n")
        f.write(f"\nsum {sum(r[4] for r in rows):.3f} ms; step {ms:.3f} ms\n")

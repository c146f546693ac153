"""run-to-run determinism of the recurrence kernel and of the training forward/backward with poisoned allocator memory"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cruse_b200 import ops, pipeline
from cruse_b200.cruse_net import unet_2
from oracle import cruse_oracle as o

dev = torch.device("cuda:0")


def poison():
    junk = [torch.full((64 << 20,), float("nan"), device=dev) for _ in range(4)]
    junk2 = [torch.full((1 << 20,), 1e30, device=dev) for _ in range(64)]
    del junk, junk2


torch.manual_seed(0)
G, H = 4, 256
for B, T in ((3, 21), (2, 16), (16, 40), (20, 33)):
    w_hh = [(0.06 * torch.randn(3 * H, H)).to(dev) for _ in range(G)]
    b_hh = [(0.06 * torch.randn(3 * H)).to(dev) for _ in range(G)]
    xproj = torch.randn(B * T, G, 3 * H).to(dev)
    outs = []
    for rep in range(6):
        poison()
        y, gates = ops.gru_seq_fwd(xproj, w_hh, b_hh, B, T, interleave=(rep % 2 == 0), mode="tf32", want_gates=True)
        torch.cuda.synchronize()
        outs.append((y.clone(), gates.clone(), rep % 2))
    same_a = all(torch.equal(outs[0][0], x[0]) and torch.equal(outs[0][1], x[1]) for x in outs if x[2] == 0)
    same_b = all(torch.equal(outs[1][0], x[0]) and torch.equal(outs[1][1], x[1]) for x in outs if x[2] == 1)
    print(f"gru_seq B={B} T={T}: deterministic interleave={same_a} cat={same_b}  finite={bool(torch.isfinite(outs[0][0]).all())}")

ref = o.make_model(256, act="relu", eval_stats=False)
ours = unet_2(in_feat=256, act="relu")
ours.load_state_dict(ref.state_dict())
ours = ours.to(dev).train()
noisy, clean = o.synth_batch(3, 6400)
res = []
for rep in range(5):
    poison()
    for p in ours.parameters():
        p.grad = None
    loss = pipeline.train_forward_loss(ours, noisy.to(dev), clean.to(dev), 512, 320)
    loss.backward()
    torch.cuda.synchronize()
    res.append((float(loss), {n: p.grad.clone() for n, p in ours.named_parameters() if p.grad is not None}))
print("losses:", [f"{r[0]:.9f}" for r in res])
bitwise = sum(all(torch.equal(res[0][1][n], r[1][n]) for r in res[1:]) for n in res[0][1])
print(f"gradient tensors bitwise identical over {len(res)} runs: {bitwise} of {len(res[0][1])}")
for n in res[0][1]:
    d = max(float((res[0][1][n] - r[1][n]).abs().max() / res[0][1][n].abs().max().clamp_min(1e-30)) for r in res[1:])
    if d > 1e-5:
        print(f"  grad {n}: max run-to-run rel diff {d:.2e}")
print("done")

"""Developer tool: a short, ncu-friendly run of the kernels of the inference step at the bench shapes (cfg-2: 32 x 10 s), each C-ABI
entry launched alone on the current stream (no wavefront, no graph) so that `ncu --set full -k regex:...` captures them.
   python tools/ncu_target.py gru|side"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cruse_b200 import ops, pipeline
from cruse_b200.cruse_net import unet_2

what = sys.argv[1] if len(sys.argv) > 1 else "side"
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = unet_2(in_feat=256)
bench.randomise_bn(model)
model = model.to(dev).eval()
B, L, T, G, H = 32, 160000, 501, 4, 256
if what == "gru":
    x = torch.randn(T * B, G * H, device=dev)
    grus = model.gru.gru_list1
    w_ih, b_ih, b_hh, w_hh = ([g.weight_ih_l0 for g in grus], [g.bias_ih_l0 for g in grus], [g.bias_hh_l0 for g in grus],
                              [g.weight_hh_l0 for g in grus])
    for _ in range(2):
        xp = ops.gru_ih_gemm(x[: 63 * B], w_ih, b_ih, b_hh, mode="tf32")            # one wavefront chunk (63 frames x 32 utterances)
    xp = ops.gru_ih_gemm(x, w_ih, b_ih, b_hh, mode="tf32")
    for _ in range(2):
        ops.gru_seq_fwd(xp, w_hh, b_hh, B, T, interleave=False, mode="tf32")        # the recurrence, all 501 steps, alone on the GPU
else:
    noisy, clean = bench.synth_batch(B, L, 20260)
    old = ops.GRU_WAVEFRONT
    ops.GRU_WAVEFRONT = False                                                         # layers back to back: every stage is a whole-tensor launch
    with torch.no_grad():
        for _ in range(2):
            pipeline.forward_loss(model, noisy.to(dev), clean.to(dev), 512, 320)
    ops.GRU_WAVEFRONT = old
torch.cuda.synchronize()

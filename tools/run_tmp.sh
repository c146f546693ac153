timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 -k "wavefront_modes" 2>&1 | grep -E "^E|assert|passed|failed" | head
timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from cruse_b200 import ops, pipeline
from cruse_b200.cruse_net import unet_2
from oracle import cruse_oracle as o
cuda=torch.device('cuda:0')
ref=o.make_model(256); ours=unet_2(in_feat=256); ours.load_state_dict(ref.state_dict()); ours=ours.to(cuda).eval()
g = torch.Generator().manual_seed(5)
noisy, clean = 0.1 * torch.randn(5, 96000, generator=g), 0.05 * torch.randn(5, 96000, generator=g)
out={}
for mode in ("flags","relaunch"):
    ops.GRU_WAVEFRONT_MODE=mode
    with torch.no_grad(): out[mode]=pipeline.forward_loss(ours, noisy.to(cuda), clean.to(cuda), 512, 320)
    torch.cuda.synchronize()
d=(out["flags"][3]-out["relaunch"][3]).abs()
print("max diff", float(d.max()), "first differing frame", int((d.amax(dim=(0,2))>0).nonzero()[0]) if d.max()>0 else None, "loss", float(out["flags"][0]), float(out["relaunch"][0]))
PY

"""Developer tool (CPU, uses the oracle): how sensitive are the oracle's OWN gradients to tf32-sized perturbations?
A relative 2^-11 noise on the GRU weights alone moves every upstream gradient by ~2.7e-2 rel-L2 (|log-error| sign
flips + small-batch BatchNorm), which is the floor any tf32 training path can reach against the fp32 oracle.
Output committed as profiles/grad_conditioning_r1.log (3 x 0.4 s) and profiles/grad_conditioning_r2.log (also 16 x 4 s:
the amplification is linear in the perturbation -- fp32-vs-fp64 1e-6..3e-4, 2^-11 noise 1e-2..3e-2 -- and does not vanish with size)."""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cruse_oracle as o
F,n_fft,hop=256,512,320
B,L=(int(sys.argv[1]),int(sys.argv[2])) if len(sys.argv)>2 else (3,6400)   # e.g. 16 64000
print(f'--- B={B} L={L} (T={1+L//hop})')
def run(dtype, perturb=0.0):
    ref=o.make_model(F, act='relu', eval_stats=False).to(dtype); ref.train()
    noisy,clean=o.synth_batch(B,L)
    if perturb:
        # emulate tf32 rounding of GRU weights: relative 2^-11 noise on W_hh and W_ih
        g=torch.Generator().manual_seed(1)
        for n,p in ref.named_parameters():
            if 'gru_list' in n and 'weight' in n:
                p.data.mul_(1+perturb*(2*torch.rand(p.shape,generator=g,dtype=torch.float64).to(dtype)-1))
    loss=o.forward_loss(ref,noisy.to(dtype),clean.to(dtype),n_fft,hop)[0]
    loss.backward()
    return {n:p.grad.double() for n,p in ref.named_parameters() if p.grad is not None}, float(loss)
g64,l64=run(torch.float64)
g32,l32=run(torch.float32)
gp,lp=run(torch.float64, 2**-11)
print('loss',l64,l32,lp)
for n in ['conv1.weight','conv4.weight','gru.ln2.weight','gru.gru_list1.0.weight_hh_l0','conv4_t.weight','conv3_t.weight','conv2_t.weight','conv1_t.weight','bn4.weight']:
    a,b,c=g64[n].flatten(),g32[n].flatten(),gp[n].flatten()
    print(f"{n:34s} fp32-vs-fp64 relL2 {float((a-b).norm()/a.norm()):.2e}   tf32-like weight noise (2^-11) relL2 {float((a-c).norm()/a.norm()):.2e}")

"""developer check of cruse_gemm_tc's operand-major combinations on the GPU (prints errors instead of asserting)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cruse_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (M, N, K, splitk) in [(128, 256, 32, 1), (128, 256, 64, 1), (768, 256, 1000, 3), (96, 32, 45, 1)]:
    for a_mn, b_mn, shift in [(False, False, 0), (True, False, 0), (False, True, 0), (True, True, 0), (True, True, 1)]:
        A = torch.randn(1, M, K)
        Bm = torch.randn(1, N, K)
        want = torch.einsum("gmk,gnk->gmn", A[:, :, shift:].double(), Bm[:, :, :K - shift].double())[0]
        Kp = K + (-K) % 4
        a_dev = (A[0].t().contiguous() if a_mn else torch.nn.functional.pad(A[0], (0, Kp - K))).to(dev).contiguous()
        b_dev = (Bm[0].t().contiguous() if b_mn else torch.nn.functional.pad(Bm[0], (0, Kp - K))).to(dev).contiguous()
        part = torch.full((splitk, M * N), float("nan"), device=dev)
        ops.gemm_tc([a_dev], [b_dev], [part], M, N, K, a_dev.shape[-1], b_dev.shape[-1], N, a_mn=a_mn, b_mn=b_mn, b_kshift=shift,
                    splitk=splitk, c_plane=M * N)
        torch.cuda.synchronize()
        got = part.sum(0).view(M, N).double().cpu()
        err = float((got - want).abs().max() / want.abs().max())
        print(f"M{M} N{N} K{K} sk{splitk} a_mn={int(a_mn)} b_mn={int(b_mn)} shift={shift}: rel err {err:.3e}  got[0,:4]={got[0,:4].tolist()} want[0,:4]={want[0,:4].tolist()}",
              "nan" if torch.isnan(got).any() else "")

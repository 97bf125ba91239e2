"""One GEMM shape, a few launches (for ncu).  Usage: python -m tools.gemm_one N K taps res(0/1) out(bf16|f32) act [M]"""
import sys
import torch
from vrdone_b200.cuda_ops import CudaOps
ops = CudaOps()
N, K, taps, res, odt, act = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6])
M = int(sys.argv[7]) if len(sys.argv) > 7 else 294912
odt = torch.bfloat16 if odt == "bf16" else torch.float32
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, taps * K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
r1 = torch.randn(M, N, device="cuda") if res else None
out = torch.empty(M, N, dtype=odt, device="cuda")
for _ in range(3):
    ops.gemm(a, w, out, bias=bias, taps=taps, act=act, res1=r1)
torch.cuda.synchronize()
print("ok")

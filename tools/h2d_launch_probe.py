"""Where does the slowdown of kernels under a concurrent host->device copy come from?  Times (a) 400 tiny kernels, (b) 40 large
GEMMs, (c) 40 large HBM-bound LayerNorms on the compute stream, alone and while a side stream copies 1.1 GB of pinned host memory."""
import torch
from vrdone_b200.cuda_ops import CudaOps

ops = CudaOps()
side = torch.cuda.Stream()
src = torch.empty(160 << 20, dtype=torch.uint8, pin_memory=True)
dst = [torch.empty(160 << 20, dtype=torch.uint8, device="cuda") for _ in range(2)]
M = 262144
a = torch.randn(M, 512, device="cuda").to(torch.bfloat16)
w = (torch.randn(512, 512, device="cuda") * 0.04).to(torch.bfloat16)
bias = torch.randn(512, device="cuda")
out = torch.empty(M, 512, dtype=torch.bfloat16, device="cuda")
x32 = torch.randn(M, 512, device="cuda")
g, b = torch.ones(512, device="cuda"), torch.zeros(512, device="cuda")
tiny = torch.zeros(1024, device="cuda")


def work(kind):
    if kind == "tiny":
        for _ in range(400):
            tiny.add_(1.0)
    elif kind == "gemm":
        for _ in range(40):
            ops.gemm(a, w, out, bias=bias)
    else:
        for _ in range(40):
            ops.layernorm(x32, g, b, out, relu=False, lay=None, streams=1)


def run(kind, copies):
    work(kind)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if copies:
        with torch.cuda.stream(side):
            for j in range(copies):
                dst[j % 2].copy_(src, non_blocking=True)
    work(kind)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


ops.bind_stream()
for kind in ("tiny", "gemm", "ln"):
    print(f"{kind:5s} alone {run(kind, 0):7.2f} ms   with 1.1 GB H2D {run(kind, 7):7.2f} ms   alone {run(kind, 0):7.2f} ms")

"""Timing experiments on the K = 512 GEMM shapes (VRD_GEMM_DBG / VRD_GEMM_CG / VRD_GEMM_WIDE are read once per process)."""
import sys
import torch
from vrdone_b200.cuda_ops import CudaOps
ops = CudaOps()
M = 294912
for name, N, K, odt, act in (("qkv 512->512 bf16", 512, 512, torch.bfloat16, 0), ("mlp0 512->2048 gelu bf16", 2048, 512, torch.bfloat16, 2),
                             ("fuse1 512->512 f32", 512, 512, torch.float32, 0), ("k2048->512 bf16", 512, 2048, torch.bfloat16, 0)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, dtype=odt, device="cuda")
    for _ in range(3):
        ops.gemm(a, w, out, bias=bias, act=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, w, out, bias=bias, act=act)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{name:28s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s")

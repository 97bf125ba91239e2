"""GEMM / kernel microbenchmarks on the B200 (CUDA events, inputs larger than L2).  Usage: python -m tools.gemm_bench [M]"""
import sys

import torch

from vrdone_b200.cuda_ops import CudaOps

ops = CudaOps()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 98304
shapes = [  # (name, N, K, taps, res, out dtype, act)
    ("embd0 3x1024->512 f32", 512, 1024, 3, 0, torch.float32, 0),
    ("embd1 3x512->512 f32", 512, 512, 3, 0, torch.float32, 0),
    ("embd0 3x1024->512 +ln bf16", 512, 1024, 3, 0, torch.bfloat16, -1),
    ("embd1 3x512->512 +ln bf16", 512, 512, 3, 0, torch.bfloat16, -1),
    ("fuse0 1024->512 gelu bf16", 512, 1024, 1, 0, torch.bfloat16, 2),
    ("qkv 512->512 bf16", 512, 512, 1, 0, torch.bfloat16, 0),
    ("proj 512->512 +res f32", 512, 512, 1, 1, torch.float32, 0),
    ("proj 512->512 +res +res2 f32", 512, 512, 1, 2, torch.float32, 0),
    ("mlp0 512->2048 gelu bf16", 2048, 512, 1, 0, torch.bfloat16, 2),
    ("mlp3 2048->512 +res f32", 512, 2048, 1, 1, torch.float32, 0),
    ("lat 512->256 f32", 256, 512, 1, 0, torch.float32, 0),
    ("fuse1 512->512 f32", 512, 512, 1, 0, torch.float32, 0),
]
print(f"M = {M}")
for name, N, K, taps, res, odt, act in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, taps * K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    r1 = torch.randn(M, N, device="cuda") if res else None
    r2 = torch.randn(M, N, device="cuda") if res == 2 else None
    out = torch.empty(M, N, dtype=odt, device="cuda")
    if act < 0:      # LayerNorm + ReLU epilogue (the tile spans the 512-channel row)
        gm, be = torch.rand(N, device="cuda") + 0.5, torch.randn(N, device="cuda")
        run = lambda: ops.gemm_ln(a, w, out, (gm, be), bias=bias, taps=taps, relu=True)
    else:
        run = lambda: ops.gemm(a, w, out, bias=bias, taps=taps, act=act, res1=r1, res2=r2)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * M * N * K * taps
    by = M * K * 2 + M * N * out.element_size() + M * N * 4 * res
    print(f"{name:28s} {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  {by/ms/1e6:8.1f} GB/s (compulsory)")
    del a, w, out, r1, r2

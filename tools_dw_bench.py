"""Timing experiments on dwconv_ln_tile (VRD_DW_DEBUG bit flags skip parts of the kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from vrdone_b200.cuda_ops import CudaOps
from vrdone_b200.layout import PackLayout
ops = CudaOps()
rng = np.random.default_rng(0)
lens = rng.integers(2, 300, 1300).tolist()
lay = PackLayout(lens, [512] * len(lens), 4, "cuda").levels[0]
R, C = lay.R, 512
x = torch.randn(2 * R, C, device="cuda")
g = lambda: torch.randn(C, device="cuda")
for nb, pre_flags in ((3, (1, 1, 1)), (3, (1, 1, 0)), (2, (0, 0)), (1, (1,))):
    branches = [(torch.randn(3, C, device="cuda"), bool(p), g(), g(), torch.empty(2 * R, C, dtype=torch.bfloat16, device="cuda")) for p in pre_flags]
    pre = (g(), g()) if any(pre_flags) else None
    byts = 2 * R * (C * 4 + nb * C * 2)
    for dbg in (0, 1, 2, 4, 8, 15):
        os.environ["VRD_DW_DEBUG"] = str(dbg)
        for _ in range(3):
            ops.dwconv_ln(x, lay, lay, 1, pre, branches, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.dwconv_ln(x, lay, lay, 1, pre, branches, 2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"NB={nb} pre={pre_flags} rows={2*R} dbg={dbg:2d}: {ms*1e3:7.1f} us  {byts/ms/1e6:7.1f} GB/s")

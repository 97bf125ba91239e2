#!/bin/bash
mkdir -p gpurun_out
echo "== attention tests ctas=3"; VRD_FA_CTAS=3 timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -4
echo "== attention tests ctas=3 safe"; VRD_FA_CTAS=3 VRD_FA_SAFE=1 timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x -k "attention" 2>&1 | tail -4
echo "== attn bench ctas=2"; timeout -k 10 300 python -m tools.ab_attn 2>&1 | tail -1
echo "== attn bench ctas=3"; VRD_FA_CTAS=3 timeout -k 10 300 python -m tools.ab_attn 2>&1 | tail -1
echo "== attn bench ctas=3 safe"; VRD_FA_CTAS=3 VRD_FA_SAFE=1 timeout -k 10 300 python -m tools.ab_attn 2>&1 | tail -1

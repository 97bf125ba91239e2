#!/bin/bash
echo "== gemm tests"; timeout -k 10 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -x 2>&1 | tail -3
echo "== gemm bench (wide K>=1024)"; timeout -k 10 300 python -m tools.gemm_bench 294912 2>&1 | grep -E "fuse0|qkv|mlp0"
echo "== gemm bench spec=4 (no wide)"; VRD_GEMM_SPEC=4 timeout -k 10 300 python -m tools.gemm_bench 294912 2>&1 | grep -E "fuse0|qkv|mlp0"
echo "== forward tests"; timeout -k 10 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --tb=short -x 2>&1 | tail -3

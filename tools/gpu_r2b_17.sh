#!/bin/bash
echo "== gemm tests"; timeout -k 10 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -x 2>&1 | tail -3
echo "== gemm tests spec=3"; VRD_GEMM_SPEC=3 timeout -k 10 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -x 2>&1 | tail -3
echo "== gemm bench"; timeout -k 10 300 python -m tools.gemm_bench 294912 2>&1 | tail -14
echo "== gemm bench spec=3 (512-column tiles for plain bf16 512->512)"; VRD_GEMM_SPEC=3 timeout -k 10 300 python -m tools.gemm_bench 294912 2>&1 | grep -E "qkv|mlp0"

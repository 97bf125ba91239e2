#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
echo "== gemm tests"; timeout -k 10 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -x 2>&1 | tail -8
echo "== forward tests"; timeout -k 10 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --tb=short -x 2>&1 | tail -6
for i in 1 2; do
echo "== ab embed_ln=1"; timeout -k 10 600 python -m tools.ab_switch 4 4 one 2>&1 | tail -1 | cut -c1-260
echo "== ab embed_ln=0"; VRD_EMBED_LN=0 timeout -k 10 600 python -m tools.ab_switch 4 4 one 2>&1 | tail -1 | cut -c1-260
done

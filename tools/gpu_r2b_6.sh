#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
echo "== gemm tests"; timeout -k 10 600 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_kernels.py -m gpu -q --tb=short -x 2>&1 | tail -4
echo "== ab"; timeout -k 10 600 python -m tools.ab_switch 4 4 dw > $G/ab4.json 2> $G/ab4.err; echo "rc=$?"; cat $G/ab4.json; tail -3 $G/ab4.err
echo "== forward tests, dw=5"; VRD_DW_CFG=5 timeout -k 10 600 python -m pytest tests/test_gpu_forward.py -m gpu -q --tb=short -x 2>&1 | tail -4

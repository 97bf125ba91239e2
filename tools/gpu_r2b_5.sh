#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
echo "== gemm tests"; timeout -k 10 600 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_kernels.py -m gpu -q --tb=short -x 2>&1 | tail -8
echo "== all tests"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -x > $G/t_all.log 2>&1; echo "rc=$?"; tail -6 $G/t_all.log
echo "== ab"; timeout -k 10 600 python -m tools.ab_switch 4 3 short > $G/ab3.json 2> $G/ab3.err; echo "rc=$?"; cat $G/ab3.json; tail -3 $G/ab3.err
echo "== gemm bench"; timeout -k 10 300 python -m tools.gemm_bench 294912 2>&1 | tee $G/gemm_bench_spec.log | tail -30
echo "== gemm bench spec=2"; VRD_GEMM_SPEC=2 timeout -k 10 300 python -m tools.gemm_bench 294912 2>&1 | tee $G/gemm_bench_spec2.log | tail -30

#!/bin/bash
# First-contact GPU run: each stage in its own process so that a CUDA fault in one does not poison the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
echo "== kernels" ; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_kernels.log
echo "== debug fp32"; timeout 600 python tools_gpu_debug.py vidvrd fp32 > gpurun_out/debug_fp32.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/debug_fp32.log
echo "== tcgen05 gemm"; timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short > gpurun_out/t_tcgen05.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_tcgen05.log
echo "== debug bf16"; timeout 600 python tools_gpu_debug.py vidvrd bf16 > gpurun_out/debug_bf16.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/debug_bf16.log
echo "== forward"; timeout 1500 python -m pytest tests/test_gpu_forward.py -m gpu -q --tb=short > gpurun_out/t_forward.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_forward.log
echo "== bench small"; timeout 900 python bench.py --steps 2 --warmup 1 --tracklets 12 --videos 1 --cpu-pairs 8 > gpurun_out/bench_small.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench_small.log

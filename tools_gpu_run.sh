#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --tb=short -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_all.log
echo "== bench full"; timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_full.log

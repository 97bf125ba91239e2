#!/bin/bash
mkdir -p gpurun_out
echo "== kernels+forward"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -m gpu -q -x --tb=short > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_all.log
echo "== gemm bench"; timeout 600 python tools_gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/gemm_bench.log
echo "== bench full"; timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_full.log
echo "== ncu launch list"
CMD="python bench.py --steps 1 --warmup 1 --videos 1 --tracklets 24 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu1.log
echo "== ncu full gemm"
CMD2="python tools_gemm_bench.py 32768"
$CMD2 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 45 -c 3 -o gpurun_out/prof_gemm $CMD2 > gpurun_out/ncu2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu2.log

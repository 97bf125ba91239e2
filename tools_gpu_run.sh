#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_all.log
echo "== bench full"; timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_full.log
echo "== ncu launch list"
CMD="python bench.py --steps 1 --warmup 1 --videos 1 --tracklets 24 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "rc=$?"; tail -c 600 gpurun_out/ncu1.log
echo "== ncu full, forward kernels"
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|dwconv_ln|flash_attn|window_attn|pack_pairs|layernorm" -s 150 -c 60 -o gpurun_out/prof_fwd $CMD > gpurun_out/ncu3.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/ncu3.log

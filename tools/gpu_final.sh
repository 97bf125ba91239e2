#!/bin/bash
# Round-end evidence run on one B200: tests, smoke, bench for every BASELINE.json config, reference arm, ncu launch list + GEMM DRAM traffic.
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_all.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "rc=$?"; tail -c 600 gpurun_out/bench_reference.log
for c in vidor_local vidor_x; do echo "== bench $c"; timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; echo "rc=$?"; done
echo "== bench vidvrd"; timeout 900 python bench.py --config vidvrd --tracklets 6 --frames 150 --cpu-pairs 30 > gpurun_out/bench_vidvrd.log 2>&1; echo "rc=$?"
echo "== bench fp32"; timeout 900 python bench.py --precision fp32 --tracklets 16 --steps 4 --no-cpu-baseline > gpurun_out/bench_vidor_fp32.log 2>&1; echo "rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --videos 1 --no-cpu-baseline"
echo "== ncu launch list"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "rc=$?"
echo "== ncu gemm traffic"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 -s 236 -c 118 --csv --log-file gpurun_out/gemm_traffic.csv $CMD > gpurun_out/ncu2.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -5

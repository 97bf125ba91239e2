#!/bin/bash
# Round-2 final evidence (trimmed to the GPU minutes left): tests, smoke, the bf16 configs, ncu launch list and GEMM DRAM traffic.
# The fp32 / vidvrd rows and the ncu detail pages stay from tools/gpu_final_r2b.sh (their kernels did not change afterwards).
mkdir -p gpurun_out
G=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --tb=short > $G/f_tests.log 2>&1; echo "rc=$?"; tail -3 $G/f_tests.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench default"; timeout 1500 python bench.py --steps 5 --warmup 3 > $G/f_bench_default.json 2> $G/f_bench_default.err; echo "rc=$?"
for c in vidor_local vidor_x; do
  echo "== bench $c bf16"; timeout 900 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --sweep-videos 0 > $G/f_bench_$c.json 2> $G/f_bench_$c.err; echo "rc=$?"
done
echo "== bench r1 workload (40 x 1200)"; timeout 900 python bench.py --tracklets 40 --frames 1200 --videos 2 --steps 10 --warmup 3 --no-cpu-baseline --sweep-videos 0 > $G/f_bench_r1workload.json 2> $G/f_bench_r1workload.err; echo "rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
echo "== ncu launch list"
$CMD > $G/f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 14500 -c 7500 --csv --log-file $G/f_launches.csv $CMD > $G/f_ncu1.log 2>&1; echo "rc=$?"
echo "== gemm microbench"; timeout -k 10 300 python -m tools.gemm_bench 294912 > $G/f_gemm_bench.log 2>&1; tail -14 $G/f_gemm_bench.log

#!/bin/bash
# usage: tools/gpu_run.sh [tests] [bench] [launches] [ncu]
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --videos 1 --tracklets 24 --no-cpu-baseline"
for what in "$@"; do
case $what in
gemm) echo "== gemm tests + microbench"; timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=short -x 2>&1 | tail -15
  echo "-- forced cta pairs"; VRD_GEMM_CG=2 timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -m gpu -q --tb=line -x 2>&1 | tail -5
  echo "-- cg1"; VRD_GEMM_CG=1 timeout 300 python -m tools.gemm_bench 294912 2>&1 | tee gpurun_out/gemm_bench_cg1.log
  echo "-- auto"; timeout 300 python -m tools.gemm_bench 294912 2>&1 | tee gpurun_out/gemm_bench.log;;
tests) echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_all.log;;
bench) echo "== bench full"; timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_full.log;;
launches) echo "== ncu launch list"
  $CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "rc=$?"; tail -c 300 gpurun_out/ncu1.log;;
ncu) echo "== ncu full"
  $CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-dwconv_ln_tile}" -s ${NCU_S:-10} -c ${NCU_C:-4} -o gpurun_out/prof_sel $CMD > gpurun_out/ncu3.log 2>&1; echo "rc=$?"; tail -c 200 gpurun_out/ncu3.log
  ls -la gpurun_out;;
esac
done

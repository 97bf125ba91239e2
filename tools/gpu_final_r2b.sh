#!/bin/bash
# Round-2 (second session) evidence run on one B200: tests, smoke, bench for every BASELINE.json config, reference arm, ncu launch list, GEMM DRAM
# traffic and one full ncu capture of the three heaviest kernels.  Outputs land in gpurun_out/ (tools/make_profiles_r2.py turns
# them into profiles/r2_*).
mkdir -p gpurun_out
G=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q --tb=short > $G/f_tests.log 2>&1; echo "rc=$?"; tail -3 $G/f_tests.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench default"; timeout 1500 python bench.py --steps 5 --warmup 3 > $G/f_bench_default.json 2> $G/f_bench_default.err; echo "rc=$?"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > $G/f_bench_reference.json 2> $G/f_bench_reference.err; echo "rc=$?"
for c in vidor_local vidor_x; do
  echo "== bench $c bf16"; timeout 900 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --sweep-videos 0 > $G/f_bench_$c.json 2> $G/f_bench_$c.err; echo "rc=$?"
done
echo "== bench vidor_local fp32"; timeout 900 python bench.py --config vidor_local --precision fp32 --steps 2 --warmup 1 --no-cpu-baseline --sweep-videos 0 > $G/f_bench_vidor_local_fp32.json 2> $G/f_bench_vidor_local_fp32.err; echo "rc=$?"
echo "== bench vidor fp32"; timeout 900 python bench.py --precision fp32 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $G/f_bench_vidor_fp32.json 2> $G/f_bench_vidor_fp32.err; echo "rc=$?"
echo "== bench vidvrd"; timeout 900 python bench.py --config vidvrd --tracklets 6 --frames 150 --videos 4 --steps 20 --warmup 3 --cpu-pairs 30 --sweep-videos 0 > $G/f_bench_vidvrd.json 2> $G/f_bench_vidvrd.err; echo "rc=$?"
echo "== bench r1 workload (40 x 1200)"; timeout 900 python bench.py --tracklets 40 --frames 1200 --videos 2 --steps 10 --warmup 3 --no-cpu-baseline --sweep-videos 0 > $G/f_bench_r1workload.json 2> $G/f_bench_r1workload.err; echo "rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
echo "== ncu launch list"
$CMD > $G/f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 14500 -c 7500 --csv --log-file $G/f_launches.csv $CMD > $G/f_ncu1.log 2>&1; echo "rc=$?"
echo "== ncu gemm traffic"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 -s 2652 -c 1326 --csv --log-file $G/f_gemm_traffic.csv $CMD > $G/f_ncu2.log 2>&1; echo "rc=$?"
echo "== ncu full: gemm / attention / dwconv"
PV="python -m tools.prof_video 5 1"
$PV > $G/f_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|flash_attn_tc|dwconv_ln_tile" -s 40 -c 12 -o $G/f_prof_top $PV > $G/f_ncu3.log 2>&1; echo "rc=$?"
ncu -i $G/f_prof_top.ncu-rep --page details > $G/f_prof_top_details.txt 2>/dev/null
ls -la $G | tail -5
echo "== ncu full: attention on the long-pair video"
PV3="python -m tools.prof_video 3 1"
ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 2 -c 1 -o $G/r2_prof_d $PV3 > $G/f_ncu4.log 2>&1; echo "rc=$?"
echo "== gemm microbench"; timeout -k 10 300 python -m tools.gemm_bench 294912 > $G/f_gemm_bench.log 2>&1; tail -12 $G/f_gemm_bench.log
echo "== attention microbench"; timeout -k 10 300 python -m tools.ab_attn > $G/f_attn_bench.json 2>&1; tail -1 $G/f_attn_bench.json

#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
PV="python -m tools.prof_video 0 1"
export VRD_PDL=0
$PV > $G/p_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -s 14 -c 14 -o $G/prof_gemm $PV > $G/p_ncu_gemm.log 2>&1; echo "rc=$?"
ncu -i $G/prof_gemm.ncu-rep --page details > $G/prof_gemm_details.txt 2>/dev/null
ls -la $G | tail -4

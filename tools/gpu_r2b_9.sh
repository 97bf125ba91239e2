#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
PV="python -m tools.prof_video 0 1"
$PV > $G/p_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"flash_attn_tc" -s 2 -c 2 -o $G/prof_fa $PV > $G/p_ncu_fa.log 2>&1; echo "rc=$?"
ncu -i $G/prof_fa.ncu-rep --page details > $G/prof_fa_details.txt 2>/dev/null
ls -la $G | tail -3

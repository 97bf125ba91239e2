#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
PV="python -m tools.prof_video 5 1"
export VRD_PDL=0
VRD_DW_CFG=4 $PV > $G/p_plain.log 2>&1 && VRD_DW_CFG=4 ncu --set full --clock-control none --import-source on -k regex:"dwconv_ln_qr" -s 2 -c 3 -o $G/prof_qr $PV > $G/p_ncu_qr.log 2>&1; echo "rc=$?"
ncu -i $G/prof_qr.ncu-rep --page details > $G/prof_qr_details.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"dwconv_ln_tile" -s 2 -c 3 -o $G/prof_tile $PV > $G/p_ncu_tile.log 2>&1; echo "rc=$?"
ncu -i $G/prof_tile.ncu-rep --page details > $G/prof_tile_details.txt 2>/dev/null
ls -la $G

#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
echo "== attn bench new"; timeout -k 10 300 python -m tools.ab_attn 2>&1 | tail -1
echo "== attn bench base"; VRD_LIB_PATH=$PWD/vrdone_b200/libvrdone_base.so timeout -k 10 300 python -m tools.ab_attn 2>&1 | tail -1
done

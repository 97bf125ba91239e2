#!/bin/bash
mkdir -p gpurun_out
echo "== no-trigger build: pdl 0 / 1"; VRD_LIB_PATH=$PWD/vrdone_b200/libvrdone_notrig.so timeout -k 10 600 python -m tools.ab_switch 4 4 pdl 2>&1 | tail -1 | cut -c1-700
echo "== default build: pdl 0 / 1"; timeout -k 10 600 python -m tools.ab_switch 4 4 pdl 2>&1 | tail -1 | cut -c1-700

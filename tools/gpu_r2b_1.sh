#!/bin/bash
# round 2, second session, call 1: PDL + quarter-row dwconv_ln validation and A/B
mkdir -p gpurun_out
G=gpurun_out
echo "== dwconv tests, qr kernel"; VRD_DW_CFG=4 timeout -k 10 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k dwconv 2>&1 | tail -8
echo "== forward tests, qr kernel"; VRD_DW_CFG=4 timeout -k 10 600 python -m pytest tests/test_gpu_forward.py -m gpu -q --tb=short -x 2>&1 | tail -8
echo "== all tests (pdl on)"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -x > $G/t_all.log 2>&1; echo "rc=$?"; tail -6 $G/t_all.log
echo "== ab"; timeout -k 10 600 python -m tools.ab_switch 4 3 > $G/ab1.json 2> $G/ab1.err; echo "rc=$?"; cat $G/ab1.json; tail -3 $G/ab1.err

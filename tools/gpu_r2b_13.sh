#!/bin/bash
mkdir -p gpurun_out
echo "== late trigger in gemm: pdl 0 / 1"; timeout -k 10 600 python -m tools.ab_switch 4 5 pdl 2>&1 | tail -1 | cut -c1-700
echo "== all tests (pdl default on)"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -3

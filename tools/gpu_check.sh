#!/bin/bash
# quick end-of-session check: GPU tests, smoke, a short default bench
mkdir -p gpurun_out
G=gpurun_out
echo "== tests"; timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short > $G/c_tests.log 2>&1; echo "rc=$?"; tail -3 $G/c_tests.log
echo "== smoke"; timeout -k 10 600 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $G/c_bench.json 2> $G/c_bench.err; echo "rc=$?"; grep -E "value:|e2e:" $G/c_bench.err | cut -c1-160

#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
echo "== bench default"; timeout -k 10 1500 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $G/b_default.json 2> $G/b_default.err; echo "rc=$?"
tail -12 $G/b_default.err | cut -c1-600

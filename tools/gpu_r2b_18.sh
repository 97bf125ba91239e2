#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
echo "== bench seeds"; timeout -k 10 1500 python bench.py --steps 5 --warmup 3 --extra-seeds 1,2 --no-cpu-baseline --no-extras > $G/b_seeds.json 2> $G/b_seeds.err; echo "rc=$?"
grep "seeds" $G/b_seeds.err | cut -c1-900

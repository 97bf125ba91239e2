#!/bin/bash
# multi-GPU run of bench.py (weak-scaling legs + the config-5 strong-scaling sweep); usage: gpurun --gpus N -- 'bash tools/gpu_r2b_8.sh N'
N=${1:-2}
mkdir -p gpurun_out
G=gpurun_out
echo "== bench $N gpus"; timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > $G/b_n$N.json 2> $G/b_n$N.err; echo "rc=$?"
grep "\[bench\]" $G/b_n$N.err | cut -c1-300 | tail -6
python - <<PY
import json
d=json.load(open('gpurun_out/b_n$N.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['e2e'].get('tracklet_api_value'))
print(json.dumps(d.get('sweep')))
PY

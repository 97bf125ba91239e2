#!/bin/bash
mkdir -p gpurun_out
G=gpurun_out
echo "== bench 2 gpus"; timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > $G/b_n2.json 2> $G/b_n2.err; echo "rc=$?"
grep "\[bench\]" $G/b_n2.err | cut -c1-400 | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_n2.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['e2e'].get('tracklet_api_value'))
print(json.dumps(d.get('sweep')))
PY

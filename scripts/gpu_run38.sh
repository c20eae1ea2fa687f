#!/bin/bash
cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/run38_n8.json 2> gpurun_out/run38_n8.err; echo "n8 rc=$?"
tail -c 400 gpurun_out/run38_n8.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/run38_n8.json').read().strip().splitlines()[-1])
    print(j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'])
except Exception as e:
    print('ERR', e)
PY

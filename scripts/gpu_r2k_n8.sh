#!/bin/bash
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 --no-parity --no-cpu-baseline --no-other-configs > $O/r2k_bench_n8.json 2> $O/r2k_bench_n8.err; echo "bench rc=$?"; tail -c 500 $O/r2k_bench_n8.err
python -c "
import json
ls=[l for l in open('$O/r2k_bench_n8.json') if l.startswith('{')]
if not ls: raise SystemExit('no bench line')
d=json.loads(ls[-1])
print('N=8 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']); print('train', d['train']); print('strip', d['strip'])"

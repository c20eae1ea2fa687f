#!/bin/bash
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-configs > $O/r2h_bench_n2.json 2> $O/r2h_bench_n2.err; echo "bench rc=$?"; tail -c 600 $O/r2h_bench_n2.err
python -c "
import json
ls=[l for l in open('$O/r2h_bench_n2.json') if l.startswith('{')]
if not ls: raise SystemExit('no bench line')
d=json.loads(ls[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step']); print('train', d['train']); print('strip', d['strip']); print('parity', d['parity'], d['fp32_grade'])"

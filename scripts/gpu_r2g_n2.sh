#!/bin/bash
# 2-GPU call: strip partition with peer-store halos vs all-gather vs unpartitioned (correctness), then the bench at N=2
# (batch-sharded value, train with the gradient all-reduce, strip strong scaling).
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_strip_gpu.py > $O/r2g_strip.log 2>&1; echo "strip rc=$?"; grep "rank" $O/r2g_strip.log | head; tail -5 $O/r2g_strip.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $O/r2g_bench_n2.json 2> $O/r2g_bench_n2.err; echo "bench rc=$?"; tail -c 600 $O/r2g_bench_n2.err
python -c "
import json
ls=[l for l in open('$O/r2g_bench_n2.json') if l.startswith('{')]
if not ls: raise SystemExit('no bench line')
d=json.loads(ls[-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step']); print('train', d['train']); print('strip', d['strip'])"

#!/bin/bash
# Final single-GPU validation of the round: full GPU test suite, smoke, reference arm, full bench.
O=gpurun_out
T=${1:-r2final}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/${T}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_ref.json 2> $O/${T}_ref.err; echo "ref rc=$?"; cut -c1-300 $O/${T}_ref.json
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -c 300 $O/${T}_bench.err
python -c "
import json
d=json.load(open('$O/${T}_bench.json'))
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
print('train', d['train']['ms_per_step'], 'f16x2', d['fp32_grade']['ms_per_step'], d['fp32_grade']['parity'], 'parity', d['parity']['max_rel_err'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','ms_per_forward','ms_per_forward_cuda_graph','parity','value')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'], 'gpu_ref', d['gpu_reference_port']['value'])"

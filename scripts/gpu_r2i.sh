#!/bin/bash
# GPU call r2i: A/B of the A-stage count of k_apply_tc, full GPU test suite, smoke, full bench.
O=gpurun_out
Q="--no-cpu-baseline --no-parity --no-train --no-other-configs"
rm -f $O/r2i_ab.log
for V in "base" "NNCONV_APPLY_PASSES=4" "base" "NNCONV_APPLY_PASSES=4"; do
  if [ "$V" = base ]; then E=""; else E="$V"; fi
  env $E timeout 300 python bench.py --steps 5 --warmup 3 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('AB %-30s ms/step %.2f apply %.2f l1 %.2f hidden %.2f clk %s' % ('$V', d['ms_per_step'], k['apply_fused']['ms'], k['edge_layer1']['ms'], k['hidden_gemm']['ms'], d['clocks']['sm_mhz']))" >> $O/r2i_ab.log 2>&1
done
cat $O/r2i_ab.log
timeout 1500 python -m pytest tests -m gpu -q > $O/r2i_tests.log 2>&1; tail -4 $O/r2i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2i_bench.json 2> $O/r2i_bench.err; echo bench rc=$?; tail -c 400 $O/r2i_bench.err
python -c "
import json
d=json.load(open('$O/r2i_bench.json'))
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print('train', d['train']['ms_per_step'], 'f16x2', d['fp32_grade']['ms_per_step'], d['fp32_grade']['parity'], 'parity', d['parity']['max_rel_err'])
for k,v in d['configs'].items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','ms_per_forward','ms_per_forward_cuda_graph','parity','value')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'], 'gpu_ref', d['gpu_reference_port']['value'])"

#!/bin/bash
# GPU call r2f: tests after the f16x2 weight scaling / SIMT fixes / formulation-B test fix, L2 A/B, train probe.
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward_tc.py tests/test_gpu_fullsize.py tests/test_gpu_graphs.py -m gpu -q > $O/r2f_tests.log 2>&1; tail -4 $O/r2f_tests.log
Q="--no-cpu-baseline --no-parity --no-train --no-other-configs"
rm -f $O/r2f_ab.log
for V in "base" "NNCONV_L2_PERSIST=48" "NNCONV_L2_PERSIST=48 NNCONV_GEMM_B_POLICY=0" "NNCONV_L2_PERSIST=48 NNCONV_L2_RESET=1" "NNCONV_L2_PERSIST=64" "NNCONV_GEMM_B_POLICY=0" "base"; do
  if [ "$V" = base ]; then E=""; else E="$V"; fi
  env $E timeout 300 python bench.py --steps 5 --warmup 3 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('AB %-50s ms/step %.2f apply %.2f l1 %.2f hidden %.2f clk %s' % ('$V', d['ms_per_step'], k['apply_fused']['ms'], k['edge_layer1']['ms'], k['hidden_gemm']['ms'], d['clocks']['sm_mhz']))" >> $O/r2f_ab.log 2>&1
done
cat $O/r2f_ab.log
python scripts/train_probe.py darcy241 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > $O/r2f_bench.json 2> $O/r2f_bench.err; echo bench rc=$?; tail -c 400 $O/r2f_bench.err
python -c "
import json
d=json.load(open('$O/r2f_bench.json'))
print('ms', d['ms_per_step'], 'train', d['train']['ms_per_step'], 'f16x2', d['fp32_grade']['ms_per_step'], d['fp32_grade']['parity'], 'parity', d['parity']['max_rel_err'])"

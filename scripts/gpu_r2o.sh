#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "alternative_schedules or config2" > $O/r2o_tests.log 2>&1; tail -3 $O/r2o_tests.log
NNCONV_SCATTER_MODE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "not config3" > $O/r2o_tests_sc.log 2>&1; tail -3 $O/r2o_tests_sc.log
Q="--no-cpu-baseline --no-train --no-other-configs"
rm -f $O/r2o_ab.log
for V in "base" "NNCONV_SCATTER_MODE=1" "NNCONV_SCATTER_MODE=1 NNCONV_APPLY_PASSES=4" "base" "NNCONV_SCATTER_MODE=1"; do
  if [ "$V" = base ]; then E=""; else E="$V"; fi
  env $E timeout 400 python bench.py --steps 5 --warmup 3 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('AB %-45s ms/step %.2f apply %.2f l1 %.2f hidden %.2f clk %s parity %.3e' % ('$V', d['ms_per_step'], k['apply_fused']['ms'], k['edge_layer1']['ms'], k['hidden_gemm']['ms'], d['clocks']['sm_mhz'], d['parity']['max_rel_err']))" >> $O/r2o_ab.log 2>&1
done
cat $O/r2o_ab.log

#!/bin/bash
O=gpurun_out
Q="--no-cpu-baseline --no-parity --no-train --no-other-configs"
rm -f $O/r2q_ab.log
for V in "base" "NNCONV_OVERFLOW_CHECK=0" "NNCONV_TMAP_PROMO=1" "base" "NNCONV_OVERFLOW_CHECK=0" "NNCONV_TMAP_PROMO=1"; do
  if [ "$V" = base ]; then E=""; else E="$V"; fi
  env $E NNCONV_B200_OVERFLOW_CHECK=$( [ "$V" = "NNCONV_OVERFLOW_CHECK=0" ] && echo 0 || echo 1 ) timeout 300 python bench.py --steps 5 --warmup 3 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('AB %-30s ms/step %.2f apply %.2f l1 %.2f hidden %.2f clk %s' % ('$V', d['ms_per_step'], k['apply_fused']['ms'], k['edge_layer1']['ms'], k['hidden_gemm']['ms'], d['clocks']['sm_mhz']))" >> $O/r2q_ab.log 2>&1
done
cat $O/r2q_ab.log
python scripts/train_probe.py darcy241 | tail -1

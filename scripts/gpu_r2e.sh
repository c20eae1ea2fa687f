#!/bin/bash
# GPU call r2e: tests of the new paths, ncu launch lists (forward bench + one training step), ncu --set full of the
# dominant kernels, L2-residency A/B of the application kernel, full bench.
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward_tc.py -m gpu -q -x > $O/r2e_tests.log 2>&1; tail -3 $O/r2e_tests.log
Q="--no-cpu-baseline --no-parity --no-train --no-other-configs"
for V in "base" "NNCONV_L2_PERSIST=1" "NNCONV_Y_STORE_POLICY=1" "NNCONV_L2_PERSIST=1 NNCONV_Y_STORE_POLICY=1" "NNCONV_RING=4"; do
  if [ "$V" = base ]; then E=""; else E="$V"; fi
  env $E timeout 300 python bench.py --steps 5 --warmup 3 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('AB %-50s ms/step %.2f apply %.2f l1 %.2f hidden %.2f clk %s' % ('$V', d['ms_per_step'], k['apply_fused']['ms'], k['edge_layer1']['ms'], k['hidden_gemm']['ms'], d['clocks']['sm_mhz']))" >> $O/r2e_ab.log 2>&1
done
cat $O/r2e_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2e_launches.csv python bench.py --steps 2 --warmup 1 $Q > $O/r2e_ncu_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2e_train_launches.csv python scripts/train_probe.py darcy241 > $O/r2e_train_probe.log 2>&1; tail -2 $O/r2e_train_probe.log
for K in k_apply_tc k_dy k_dh k_gemm_tn; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c 1 -f -o $O/prof_r2e_$K python scripts/train_probe.py darcy241 > $O/r2e_ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
python scripts/train_probe.py darcy241 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2e_bench.json 2> $O/r2e_bench.err; echo bench rc=$?; tail -c 600 $O/r2e_bench.err

#!/bin/bash
# r2v: the round's final check first (tests, smoke, reference arm, bench), then the MGKN V-cycle A/B of the new knobs
O=gpurun_out
bash scripts/final_check.sh r2v
for V in "NNCONV_RING_DEEP=1" "NNCONV_RING_DEEP=0" "NNCONV_RING_DEEP=1 NNCONV_B200_FUSED_STEPS=0"; do
  env $V timeout 200 python scripts/mgkn_bench.py 2>&1 | grep "this library\|parity" | sed "s/^/[$V] /"
done

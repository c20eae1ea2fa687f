#!/bin/bash
# r2v: vectorised per-edge kernels, fused V-cycle steps, deep Y ring -- tests + MGKN / Burgers A/B, then the round's final check
O=gpurun_out
for V in "NNCONV_RING_DEEP=1" "NNCONV_RING_DEEP=0" "NNCONV_RING_DEEP=1 NNCONV_B200_Y_BYTES=100663296" "NNCONV_RING_DEEP=1 NNCONV_B200_FUSED_STEPS=0" "NNCONV_RING_DEEP=1 NNCONV_B200_EDGE_KERNELS=off"; do
  env $V timeout 200 python scripts/mgkn_bench.py 2>&1 | grep "this library\|parity" | sed "s/^/[$V] /"
done
timeout 200 python scripts/burgers1d_bench.py 2>&1 | tail -4
bash scripts/final_check.sh r2v

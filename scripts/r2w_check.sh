#!/bin/bash
# r2w: defaults after r2v (ring_deep off): the two test files that touch the changed paths, MGKN V-cycle numbers with an
# A/B of the per-edge-kernel size threshold, and the launch list of one replay-shaped (eager) forward
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -q 2>&1 | tail -2
for V in "NNCONV_B200_EDGE_KERNELS_MAX_EDGES=8192" "NNCONV_B200_EDGE_KERNELS_MAX_EDGES=16384"; do
  env $V timeout 200 python scripts/mgkn_bench.py 2>&1 | grep "this library\|parity" | sed "s/^/[$V] /"
done
MGKN_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2w_mgkn_launches.csv python scripts/mgkn_bench.py > $O/r2w_ncu_mgkn.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# one --set full capture of the first-layer (K=64) and the hidden (K=1024) tcgen05 GEMM launches of the bench
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 4 -c 2 -o gpurun_out/prof_gemm_r1g python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
cd /root/repo
echo "== TMA store"; timeout 300 python scripts/gemm_probe.py 2>&1 | tail -14
echo "== direct store"; NNCONV_GEMM_DIRECT_STORE=1 timeout 300 python scripts/gemm_probe.py 2>&1 | tail -11
PROBE_M=65536 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 8 -c 2 -o gpurun_out/prof_gemm_r1g python scripts/gemm_probe.py > gpurun_out/run29_ncu.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_tc -s 2 -c 1 -o gpurun_out/prof_apply_r1g python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/run33_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep

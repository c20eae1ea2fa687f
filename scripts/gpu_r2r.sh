#!/bin/bash
# final captures: ncu launch list of the bench command, ncu --set full of k_apply_tc (final defaults), then the full check
O=gpurun_out
Q="--no-cpu-baseline --no-parity --no-train --no-other-configs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2r_launches.csv python bench.py --steps 2 --warmup 1 $Q > $O/r2r_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_tc -s 2 -c 1 -f -o $O/prof_r2r_k_apply_tc python bench.py --steps 1 --warmup 1 $Q > $O/r2r_ncu_apply.log 2>&1; echo "ncu apply rc=$?"
scripts/final_check.sh r2r

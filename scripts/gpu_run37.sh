#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/run37_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/run37_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/run37_bench.json 2> gpurun_out/run37_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/run37_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/run37_ref.json 2> gpurun_out/run37_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --workload darcy85 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/run37_ncu.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# round-end style validation on one B200: GPU tests, smoke, default bench, bf16 bench, MGKN probe
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --precision bf16 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/final_bench_bf16.json 2> gpurun_out/final_bench_bf16.err; echo "bf16 rc=$?"
python - <<'PY'
import json
for n in ('final_bench','final_bench_bf16'):
    j=json.loads(open('gpurun_out/%s.json'%n).read().strip().splitlines()[-1])
    print(n, j['dtype'], j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['gpu_launches'], j['clocks'])
PY
timeout 300 python scripts/mgkn_bench.py 2>&1 | tail -5

#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/run35_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/run35_tests.log
timeout 300 python scripts/mgkn_bench.py 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/run35_1.json 2> gpurun_out/run35_1.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/run35_1.json').read().strip().splitlines()[-1])
print(j['value'], j['ms_per_step'], {k:round(v['ms'],2) for k,v in j['kernel_ms_per_step'].items()}, j['roofline']['frac'], j['clocks'])
PY

#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -x -q > gpurun_out/run34_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/run34_tests.log

for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/run34_$i.json 2> gpurun_out/run34_$i.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for n in ('1','2'):
    try:
        j=json.loads(open('gpurun_out/run34_%s.json'%n).read().strip().splitlines()[-1])
        print(n, j['value'], j['ms_per_step'], {k:round(v['ms'],2) for k,v in j['kernel_ms_per_step'].items()}, j['roofline']['frac'], j['clocks'])
    except Exception as e:
        print(n,'ERR',e)
PY

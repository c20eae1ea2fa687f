#!/bin/bash
cd /root/repo
for m in 0 1 2; do
NNCONV_DEBUG_SCATTER=$m timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/run30_$m.json 2> gpurun_out/run30_$m.err; echo "bench rc=$?"
done
python - <<'PY'
import json
for n in ('0','1','2'):
    try:
        j=json.loads(open('gpurun_out/run30_%s.json'%n).read().strip().splitlines()[-1])
        print(n, j['value'], j['ms_per_step'], {k:round(v['ms'],2) for k,v in j['kernel_ms_per_step'].items()}, j['clocks'])
    except Exception as e:
        print(n,'ERR',e)
PY

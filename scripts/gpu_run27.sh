#!/bin/bash
# TMA-store epilogue: correctness first, then A/B timing
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/run27_tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/run27_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/run27_bench_tma.json 2> gpurun_out/run27_bench_tma.err; echo "bench rc=$?"
NNCONV_GEMM_DIRECT_STORE=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/run27_bench_direct.json 2> gpurun_out/run27_bench_direct.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ('tma','direct'):
    try:
        j=json.loads(open('gpurun_out/run27_bench_%s.json'%n).read().strip().splitlines()[-1])
        print(n, j['value'], j['ms_per_step'], j.get('kernel_ms'), j['e2e']['value'], j.get('gpu_reference_port'), j['clocks'])
    except Exception as e:
        print(n,'ERR',e)
PY

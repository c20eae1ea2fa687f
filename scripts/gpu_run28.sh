mkdir -p gpurun_out
echo "=== tests ring"; NNCONV_MLP12=ring timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -q 2>&1 | tail -4
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks']['sm_mhz'], [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']])"; }
echo "=== split"; run
echo "=== ring"; NNCONV_MLP12=ring run
echo "=== ring darcy85"; NNCONV_MLP12=ring NNCONV_BENCH_WORKLOAD=darcy85 run

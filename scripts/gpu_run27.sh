mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks']['sm_mhz'], [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']])"; }
echo "=== default (ybn128, 6 stages)"; run
echo "=== ybn64 (7 stages)"; NNCONV_Y_BLOCKN=64 run
echo "=== ybn128, 5 stages"; NNCONV_APPLY_STAGES=5 run
echo "=== ybn128, 4 stages"; NNCONV_APPLY_STAGES=4 run
echo "=== tests ybn64"; NNCONV_Y_BLOCKN=64 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3

mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), round(d['e2e']['value']/1e6,1), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks']['sm_mhz'], [(r['kernel'][:10], round(r['frac'],3)) for r in d['roofline_kernels']])"; }
echo "=== default"; run
echo "=== persist 64MB"; NNCONV_L2_PERSIST=1 run
echo "=== persist ring2 nb128 (32MB)"; NNCONV_L2_PERSIST=1 NNCONV_RING=2 NNCONV_B200_Y_BYTES=33554432 run
echo "=== persist + evict_last stores"; NNCONV_L2_PERSIST=1 NNCONV_Y_STORE_POLICY=1 run
echo "=== default again"; run

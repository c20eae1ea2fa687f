mkdir -p gpurun_out
run() { timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,1), round(d['ms_per_step'],2), {k:round(v['ms'],1) for k,v in d['kernel_ms_per_step'].items() if v['ms']>0}, d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
echo "=== policy normal"; NNCONV_Y_STORE_POLICY=0 run
echo "=== policy evict_last"; NNCONV_Y_STORE_POLICY=1 run
echo "=== policy normal, 64MB ring2"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=67108864 NNCONV_Y_STORE_POLICY=0 run
echo "=== policy evict_last, 64MB ring2"; NNCONV_RING=2 NNCONV_B200_Y_BYTES=67108864 NNCONV_Y_STORE_POLICY=1 run
echo "=== policy normal again"; NNCONV_Y_STORE_POLICY=0 run
